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

// Backward of the ray set-up (NetWorks/utils.py:147-158): per-ray gradients w.r.t. the origin o = T, v = d * l and l
// (produced by the data-gradient kernels) -> dL/dR, dL/dT, dL/dK^-1 of the batch item.
//   c = K^-1 [x, y, 1],  d0 = R c,  d = d0 / |d0|,  l = -1 / d_z,  v = d l
// One thread per ray, block-level reduction, one atomic add per block and output element.
__global__ void __launch_bounds__(256) camera_bwd_kernel(hn_camera_t cam, const float* g_o, const float* g_v, const float* g_l,
                                                         float* dR, float* dT, float* dK) {
    __shared__ float red[8][21];
    const int b = blockIdx.y;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float acc[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) acc[i] = 0.f;
    if (r < cam.n_rays) {
        const float x = __ldg(cam.xy + ((size_t)b * 2 + 0) * cam.n_rays + r), y = __ldg(cam.xy + ((size_t)b * 2 + 1) * cam.n_rays + r);
        const float* K = cam.inv_inmats + b * 9;
        const float* R = cam.Rmats + b * 9;
        const float xyz[3] = {x, y, 1.f};
        float c[3], d0[3], d[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) c[i] = __ldg(K + 3 * i) * x + __ldg(K + 3 * i + 1) * y + __ldg(K + 3 * i + 2);
#pragma unroll
        for (int i = 0; i < 3; ++i) d0[i] = __ldg(R + 3 * i) * c[0] + __ldg(R + 3 * i + 1) * c[1] + __ldg(R + 3 * i + 2) * c[2];
        const float inv_n = rsqrtf(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]);
#pragma unroll
        for (int i = 0; i < 3; ++i) d[i] = d0[i] * inv_n;
        const float l = -1.0f / d[2];
        const size_t ray = (size_t)b * cam.n_rays + r;
        const float go[3] = {g_o[ray * 3], g_o[ray * 3 + 1], g_o[ray * 3 + 2]};
        const float gv[3] = {g_v[ray * 3], g_v[ray * 3 + 1], g_v[ray * 3 + 2]};
        const float gl = g_l[ray];
        float gd[3] = {l * gv[0], l * gv[1], l * gv[2]};
        gd[2] += l * l * (gl + gv[0] * d[0] + gv[1] * d[1] + gv[2] * d[2]);          // through l = -1 / d_z  (dl/dd_z = l^2)
        const float dg = d[0] * gd[0] + d[1] * gd[1] + d[2] * gd[2];
        float gd0[3], gc[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) gd0[i] = (gd[i] - d[i] * dg) * inv_n;                // through the normalisation
#pragma unroll
        for (int j = 0; j < 3; ++j) gc[j] = __ldg(R + j) * gd0[0] + __ldg(R + 3 + j) * gd0[1] + __ldg(R + 6 + j) * gd0[2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) { acc[3 * i + j] = gd0[i] * c[j]; acc[12 + 3 * i + j] = gc[i] * xyz[j]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) acc[9 + i] = go[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 21; ++i) {
        float v = acc[i];
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 21) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        const int i = threadIdx.x;
        if (i < 9) { if (dR) atomicAdd(dR + b * 9 + i, v); }
        else if (i < 12) { if (dT) atomicAdd(dT + b * 3 + i - 9, v); }
        else if (dK) atomicAdd(dK + b * 9 + i - 12, v);
    }
}

}  // namespace hn

extern "C" int hn_camera_bwd(const hn_camera_t* cam, const float* g_ray_o, const float* g_ray_v, const float* g_ray_l,
                             float* dR, float* dT, float* dKinv, void* stream) {
    using namespace hn;
    if (!cam || !cam->xy || !cam->Rmats || !cam->Tvecs || !cam->inv_inmats || !g_ray_o || !g_ray_v || !g_ray_l)
        return set_error(HN_E_BADARG, "hn_camera_bwd: null pointer");
    if (cam->B <= 0 || cam->n_rays <= 0) return set_error(HN_E_BADARG, "hn_camera_bwd: empty problem");
    camera_bwd_kernel<<<dim3((cam->n_rays + 255) / 256, cam->B), 256, 0, (cudaStream_t)stream>>>(*cam, g_ray_o, g_ray_v, g_ray_l, dR, dT, dKinv);
    return check_launch("hn_camera_bwd");
}

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
