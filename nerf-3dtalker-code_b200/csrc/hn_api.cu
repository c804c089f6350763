// hn_api.cu — C-ABI plumbing: version, errors, buffer sizes, and the standalone ray sampler.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "hn_api.h"
#include "hn_sample.cuh"

namespace hn {

static thread_local char g_err[256] = "";

int set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return HN_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
int check_geometry(int B, int n_rays, int n_samples, const char* who) {
    char buf[200];
    if (B <= 0 || n_rays <= 0) { snprintf(buf, sizeof buf, "%s: empty problem (B=%d, n_rays=%d)", who, B, n_rays); return set_error(HN_E_BADARG, buf); }
    if (n_samples != 32 && n_samples != 64 && n_samples != 128) { snprintf(buf, sizeof buf, "%s: n_samples must be 32, 64 or 128 (got %d)", who, n_samples); return set_error(HN_E_UNSUPPORTED, buf); }
    if (((int64_t)n_rays * n_samples) % HN_TILE != 0) { snprintf(buf, sizeof buf, "%s: n_rays*n_samples must be a multiple of %d", who, HN_TILE); return set_error(HN_E_UNSUPPORTED, buf); }
    return HN_OK;
}

bool use_cta_pairs(int n_tiles) {
    static const int env = [] { const char* e = getenv("HN_CTA_PAIRS"); return e ? atoi(e) : 0; }();
    return env != 0 && n_tiles >= 2 && (n_tiles % 2) == 0;
}

// One thread per sample; NetWorks/utils.py:147-161,64-89.
__global__ void sample_rays_kernel(hn_camera_t cam, float* pts, float* zvals, float* z_dists, float* ray_d, float* ray_l) {
    const int64_t M = (int64_t)cam.B * cam.n_rays * cam.n_samples;
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int s = (int)(m % cam.n_samples);
    const int64_t ray_idx = m / cam.n_samples;
    const int r = (int)(ray_idx % cam.n_rays), b = (int)(ray_idx / cam.n_rays);
    const Ray ray = make_ray(cam, b, r);
    const Sample q = make_sample(cam, ray, b, r, s);
    if (pts) { pts[m * 3 + 0] = q.px; pts[m * 3 + 1] = q.py; pts[m * 3 + 2] = q.pz; }
    if (zvals) zvals[m] = q.zval;
    if (z_dists) z_dists[m] = q.zdist;
    if (s == 0) {
        if (ray_d) { ray_d[ray_idx * 3 + 0] = ray.dx; ray_d[ray_idx * 3 + 1] = ray.dy; ray_d[ray_idx * 3 + 2] = ray.dz; }
        if (ray_l) ray_l[ray_idx] = ray.l;
    }
}

}  // namespace hn

extern "C" int hn_abi_version(void) { return HN_ABI_VERSION; }
extern "C" const char* hn_last_error(void) { return hn::g_err; }

extern "C" size_t hn_act_bytes(int64_t M) { return (size_t)HN_ACT_BLOCKS * (size_t)(M / HN_TILE) * 16384; }
extern "C" size_t hn_grads_bytes(int64_t M) { return (size_t)HN_GRAD_BLOCKS * (size_t)(M / HN_TILE) * 16384; }
extern "C" size_t hn_mask_bytes(int64_t M) { return (size_t)M * HN_MASK_WORDS * 4; }
extern "C" size_t hn_dfeat_image_bytes(int64_t M) { return (size_t)(HN_FEAT / 64) * (size_t)(M / HN_TILE) * 16384; }

extern "C" int hn_sample_rays(const hn_camera_t* cam, float* pts, float* zvals, float* z_dists, float* ray_d, float* ray_l, void* stream) {
    using namespace hn;
    if (!cam || !cam->xy || !cam->Rmats || !cam->Tvecs || !cam->inv_inmats) return set_error(HN_E_BADARG, "hn_sample_rays: null camera pointer");
    if (cam->B <= 0 || cam->n_rays <= 0 || cam->n_samples <= 0) return set_error(HN_E_BADARG, "hn_sample_rays: empty problem");
    const int64_t M = total_samples(cam->B, cam->n_rays, cam->n_samples);
    const int threads = 256;
    sample_rays_kernel<<<(unsigned)((M + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(*cam, pts, zvals, z_dists, ray_d, ray_l);
    return check_launch("hn_sample_rays");
}
