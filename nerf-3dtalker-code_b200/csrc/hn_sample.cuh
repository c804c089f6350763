// hn_sample.cuh — ray generation and stratified sampling, shared by the standalone sampler and by the
// fused MLP kernels (which evaluate it while producing the first GEMM's operand).
// Follows NetWorks/utils.py:147-161 (rays) and :64-89,118-142 (samples) op for op: every product/sum the
// reference performs as a separate rounded fp32 torch op is a separate rounded op here (no FMA
// contraction), because the positional encoding amplifies input rounding by up to 2^9.
#pragma once
#include "../../include/headnerf_b200.h"

namespace hn {

struct Ray {
    float ox, oy, oz;   // origin  = Tvec
    float dx, dy, dz;   // unit direction
    float l;            // -1/dz : camera-space z -> ray length
    float vx, vy, vz;   // d * l  (what multiplies zvals)
};

__device__ __forceinline__ Ray make_ray(const hn_camera_t& cam, int b, int r) {
    const float x = __ldg(cam.xy + ((size_t)b * 2 + 0) * cam.n_rays + r);
    const float y = __ldg(cam.xy + ((size_t)b * 2 + 1) * cam.n_rays + r);
    const float* K = cam.inv_inmats + b * 9;
    const float* R = cam.Rmats + b * 9;
    float c[3], d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        c[i] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(K + 3 * i), x), __fmul_rn(__ldg(K + 3 * i + 1), y)), __ldg(K + 3 * i + 2));
#pragma unroll
    for (int i = 0; i < 3; ++i)
        d[i] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(R + 3 * i), c[0]), __fmul_rn(__ldg(R + 3 * i + 1), c[1])), __fmul_rn(__ldg(R + 3 * i + 2), c[2]));
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    Ray ray;
    ray.dx = __fdiv_rn(d[0], nrm); ray.dy = __fdiv_rn(d[1], nrm); ray.dz = __fdiv_rn(d[2], nrm);
    ray.l = __fdiv_rn(-1.0f, ray.dz);
    ray.ox = __ldg(cam.Tvecs + b * 3 + 0); ray.oy = __ldg(cam.Tvecs + b * 3 + 1); ray.oz = __ldg(cam.Tvecs + b * 3 + 2);
    ray.vx = __fmul_rn(ray.dx, ray.l); ray.vy = __fmul_rn(ray.dy, ray.l); ray.vz = __fmul_rn(ray.dz, ray.l);
    return ray;
}

// un-jittered edge j of the stratification (utils.py:139): rela_z1*(1-t) + rela_z2*t, t = j/n_samples
__device__ __forceinline__ float plain_edge(float oz, float z1, float z2, int j, int n_samples) {
    const float t = (float)j / (float)n_samples;          // exact: n_samples is a power of two
    return __fadd_rn(__fmul_rn(__fsub_rn(oz, z1), __fsub_rn(1.0f, t)), __fmul_rn(__fsub_rn(oz, z2), t));
}

// edge j after optional stratified jitter (utils.py:73-78); u = t_rand[b, r, j]
__device__ __forceinline__ float sample_edge(const hn_camera_t& cam, float oz, int b, int r, int j) {
    const int ns = cam.n_samples;
    const float e = plain_edge(oz, cam.world_z1, cam.world_z2, j, ns);
    if (cam.t_rand == nullptr) return e;
    const float lower = (j == 0) ? e : __fmul_rn(0.5f, __fadd_rn(e, plain_edge(oz, cam.world_z1, cam.world_z2, j - 1, ns)));
    const float upper = (j == ns) ? e : __fmul_rn(0.5f, __fadd_rn(plain_edge(oz, cam.world_z1, cam.world_z2, j + 1, ns), e));
    const float u = __ldg(cam.t_rand + ((size_t)b * cam.n_rays + r) * (ns + 1) + j);
    return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u));
}

struct Sample { float px, py, pz, zval, zdist; };

__device__ __forceinline__ Sample make_sample(const hn_camera_t& cam, const Ray& ray, int b, int r, int s) {
    const float e0 = sample_edge(cam, ray.oz, b, r, s);
    const float e1 = sample_edge(cam, ray.oz, b, r, s + 1);
    Sample q;
    q.zval = e0;
    q.zdist = __fmul_rn(__fsub_rn(e1, e0), ray.l);
    q.px = __fadd_rn(ray.ox, __fmul_rn(ray.vx, e0));
    q.py = __fadd_rn(ray.oy, __fmul_rn(ray.vy, e0));
    q.pz = __fadd_rn(ray.oz, __fmul_rn(ray.vz, e0));
    return q;
}

}  // namespace hn
