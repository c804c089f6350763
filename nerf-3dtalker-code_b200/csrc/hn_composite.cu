// hn_composite.cu — volume-rendering alpha compositing, forward and backward (CalcRayColor,
// NetWorks/utils.py:273-309).  HBM-bound: one warp per ray, sample-major [M,C] features read with
// 128-bit coalesced loads, transmittance as a multiplicative warp scan (shuffles), backward as the
// matching reverse (suffix) scan.  Lane l owns samples l, l+32, l+64, ... of its ray.
//
// Algorithmic bytes (fp32 features, C channels, N_s samples/ray):
//   fwd : read  N_s*(C+2)*4           write (C+1)*4                per ray
//   bwd : read  N_s*(C+2)*4 + (C+1)*4 write N_s*(C*{4|2}+{4|8})    per ray
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "hn_api.h"
#include "hn_mlp_common.cuh"

namespace hn {

constexpr int kCompThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

// streaming 128-bit load: features are touched once, keep them out of L1
__device__ __forceinline__ float4 ld_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// Per-ray weights.  alpha = 1-exp(-sigma*delta); x = 1-alpha+1e-10; T = exclusive cumprod(x); w = alpha*T.
template <int SPL>
__device__ __forceinline__ void ray_weights(const float* __restrict__ sigma, const float* __restrict__ delta,
                                            size_t m0, int lane, float (&alpha)[SPL], float (&x)[SPL],
                                            float (&T)[SPL], float (&w)[SPL], float (&sg)[SPL], float (&dl)[SPL]) {
    float carry = 1.0f;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        sg[i] = __ldg(sigma + m0 + i * 32 + lane);
        dl[i] = __ldg(delta + m0 + i * 32 + lane);
        alpha[i] = 1.0f - expf(-sg[i] * dl[i]);
        x[i] = (1.0f - alpha[i]) + 1e-10f;
        float incl = x[i];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            float t = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl *= t;
        }
        float excl = __shfl_up_sync(kFull, incl, 1);
        if (lane == 0) excl = 1.0f;
        T[i] = carry * excl;
        w[i] = alpha[i] * T[i];
        carry *= __shfl_sync(kFull, incl, 31);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}

template <int NS, int CV>   // CV = C / 128
__global__ void __launch_bounds__(kCompThreads) composite_fwd_kernel(hn_composite_fwd_t a) {
    constexpr int SPL = NS / 32, C = CV * 128;
    const int lane = threadIdx.x & 31;
    const int ray = blockIdx.x * (kCompThreads / 32) + (threadIdx.x >> 5);
    if (ray >= a.n_rays_total) return;
    const size_t m0 = (size_t)ray * NS;
    float alpha[SPL], x[SPL], T[SPL], w[SPL], sg[SPL], dl[SPL];
    ray_weights<SPL>(a.sigma, a.delta, m0, lane, alpha, x, T, w, sg, dl);

    float4 acc[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* frow = a.feat + m0 * C + lane * 4;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
#pragma unroll 8
        for (int ls = 0; ls < 32; ++ls) {
            const float ws = __shfl_sync(kFull, w[i], ls);
            const float* p = frow + (size_t)(i * 32 + ls) * C;
#pragma unroll
            for (int k = 0; k < CV; ++k) {
                const float4 f = ld_stream(p + k * 128);
                acc[k].x = fmaf(ws, f.x, acc[k].x); acc[k].y = fmaf(ws, f.y, acc[k].y);
                acc[k].z = fmaf(ws, f.z, acc[k].z); acc[k].w = fmaf(ws, f.w, acc[k].w);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < CV; ++k) *reinterpret_cast<float4*>(a.F + (size_t)ray * C + k * 128 + lane * 4) = acc[k];

    float wsum = 0.f, dsum = 0.f;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        wsum += w[i];
        if (a.zvals) dsum = fmaf(w[i], __ldg(a.zvals + m0 + i * 32 + lane), dsum);
        if (a.weights) a.weights[m0 + i * 32 + lane] = w[i];
    }
    wsum = warp_sum(wsum);
    if (a.depth && a.zvals) { dsum = warp_sum(dsum); if (lane == 0) a.depth[ray] = dsum; }
    if (lane == 0) a.bg_alpha[ray] = 1.0f - wsum;
}

// 8 partial values per lane (8 rows) -> the four lanes 4j..4j+3 all receive the full sum of row j  (4+2+1+2 = 9 shuffles).
// A batch of 8 (instead of 32) rows keeps the register count low enough for 3-4 resident CTAs per SM.
__device__ __forceinline__ float transpose_reduce8(float (&part)[8], int lane) {
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {              // lane bits 4,3,2 select the row
        const int lbit = off << 2;
        const bool up = (lane & lbit) != 0;
#pragma unroll
        for (int k = 0; k < off; ++k) {
            const float send = up ? part[k] : part[k + off];
            const float keep = up ? part[k + off] : part[k];
            part[k] = keep + __shfl_xor_sync(kFull, send, lbit);
        }
    }
    float v = part[0];
    v += __shfl_xor_sync(kFull, v, 2);
    v += __shfl_xor_sync(kFull, v, 1);
    return v;
}

// resident CTAs per SM: measured 2 / 3 / 4 / 5 / 6 / 8 -> 0.170 / 0.192 / 0.158 / 0.201 / 0.210 / 0.177 ms (Reso64 batch 2): four
// (64 registers, a few spilled words) hides the most latency
#ifndef HN_COMP_BWD_CTAS
#define HN_COMP_BWD_CTAS 4
#endif
template <int NS, int CV>
__global__ void __launch_bounds__(kCompThreads, HN_COMP_BWD_CTAS) composite_bwd_kernel(hn_composite_bwd_t a, int n_tiles) {
    constexpr int SPL = NS / 32, C = CV * 128;
    const int lane = threadIdx.x & 31;
    const int ray = blockIdx.x * (kCompThreads / 32) + (threadIdx.x >> 5);
    if (ray >= a.n_rays_total) return;
    const size_t m0 = (size_t)ray * NS;
    float alpha[SPL], x[SPL], T[SPL], w[SPL], sg[SPL], dl[SPL];
    ray_weights<SPL>(a.sigma, a.delta, m0, lane, alpha, x, T, w, sg, dl);

    float4 g[CV];
#pragma unroll
    for (int k = 0; k < CV; ++k) g[k] = *reinterpret_cast<const float4*>(a.gF + (size_t)ray * C + k * 128 + lane * 4);
    const float gbg = __ldg(a.g_bg + ray);
    const float gdp = a.g_depth ? __ldg(a.g_depth + ray) : 0.f;
    const float scale = a.dfeat_image ? __ldg(a.grad_scale) : 1.0f;

    // q_s = dL/dw_s = gF . f_s - g_bg (+ g_depth z_s);  dfeat_s = w_s gF
    float q[SPL];
    const float* frow = a.feat + m0 * C + lane * 4;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
        float qi = 0.f;
#pragma unroll
        for (int b8 = 0; b8 < 4; ++b8) {
            float part[8];
#pragma unroll
            for (int l8 = 0; l8 < 8; ++l8) {
                const int ls = b8 * 8 + l8;
                const size_t s = i * 32 + ls;
                const float ws = __shfl_sync(kFull, w[i], ls);
                float d = 0.f;
#pragma unroll
                for (int k = 0; k < CV; ++k) {
                    const float4 f = ld_stream(frow + s * C + k * 128);
                    d = fmaf(g[k].x, f.x, d); d = fmaf(g[k].y, f.y, d); d = fmaf(g[k].z, f.z, d); d = fmaf(g[k].w, f.w, d);
                    if (a.dfeat) {
                        float4 o = make_float4(ws * g[k].x, ws * g[k].y, ws * g[k].z, ws * g[k].w);
                        *reinterpret_cast<float4*>(a.dfeat + (m0 + s) * C + k * 128 + lane * 4) = o;
                    }
                    if (a.dfeat_image) {
                        const float wsc = ws * scale;
                        const size_t m = m0 + s;
                        const int kb = (lane >> 4) + 2 * k, col = (lane & 15) * 4;
                        uint8_t* dst = (uint8_t*)a.dfeat_image + ((size_t)kb * n_tiles + (m >> 7)) * kBlockBytes +
                                       image_offset((uint32_t)(m & 127), col);
                        uint2 pk;
                        pk.x = pack_sat(wsc * g[k].x, wsc * g[k].y);          // one saturating conversion per pair (+-65504)
                        pk.y = pack_sat(wsc * g[k].z, wsc * g[k].w);
                        *reinterpret_cast<uint2*>(dst) = pk;
                    }
                }
                part[l8] = d;
            }
            const float r = transpose_reduce8(part, lane);           // lanes 4j..4j+3 hold row b8*8 + j
            const float mine = __shfl_sync(kFull, r, 4 * (lane & 7)); // row b8*8 + (lane % 8)
            if ((lane >> 3) == b8) qi = mine;
        }
        q[i] = qi - gbg;
        if (a.g_depth && a.zvals) q[i] = fmaf(gdp, __ldg(a.zvals + m0 + i * 32 + lane), q[i]);
    }

    // d alpha_s = T_s q_s - (sum_{j>s} q_j w_j) / x_s : suffix sums, last block first
    float carry = 0.f;
#pragma unroll
    for (int i = SPL - 1; i >= 0; --i) {
        const float v = q[i] * w[i];
        float incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            float t = __shfl_down_sync(kFull, incl, d);
            if (lane + d < 32) incl += t;
        }
        const float suffix = (incl - v) + carry;          // strictly-after sum
        carry += __shfl_sync(kFull, incl, 0);
        const float dalpha = T[i] * q[i] - suffix / x[i];
        const float e = 1.0f - alpha[i];                  // exp(-sigma*delta)
        a.dsigma[m0 + i * 32 + lane] = dalpha * dl[i] * e;
        if (a.ddelta) a.ddelta[m0 + i * 32 + lane] = dalpha * sg[i] * e;
    }
}

template <int NS>
static int launch_fwd(const hn_composite_fwd_t& a, cudaStream_t st) {
    const int blocks = (a.n_rays_total + kCompThreads / 32 - 1) / (kCompThreads / 32);
    if (a.C == 256) composite_fwd_kernel<NS, 2><<<blocks, kCompThreads, 0, st>>>(a);
    else if (a.C == 128) composite_fwd_kernel<NS, 1><<<blocks, kCompThreads, 0, st>>>(a);
    else return set_error(HN_E_UNSUPPORTED, "hn_composite_fwd: C must be 128 or 256");
    return check_launch("hn_composite_fwd");
}
template <int NS>
static int launch_bwd(const hn_composite_bwd_t& a, cudaStream_t st) {
    const int blocks = (a.n_rays_total + kCompThreads / 32 - 1) / (kCompThreads / 32);
    const int n_tiles = (int)(((int64_t)a.n_rays_total * NS) / 128);
    if (a.C == 256) composite_bwd_kernel<NS, 2><<<blocks, kCompThreads, 0, st>>>(a, n_tiles);
    else if (a.C == 128) composite_bwd_kernel<NS, 1><<<blocks, kCompThreads, 0, st>>>(a, n_tiles);
    else return set_error(HN_E_UNSUPPORTED, "hn_composite_bwd: C must be 128 or 256");
    return check_launch("hn_composite_bwd");
}

}  // namespace hn

extern "C" int hn_composite_fwd(const hn_composite_fwd_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->feat || !a->sigma || !a->delta || !a->F || !a->bg_alpha || a->n_rays_total <= 0)
        return set_error(HN_E_BADARG, "hn_composite_fwd: null pointer or empty problem");
    cudaStream_t st = (cudaStream_t)stream;
    switch (a->n_samples) {
        case 32: return launch_fwd<32>(*a, st);
        case 64: return launch_fwd<64>(*a, st);
        case 128: return launch_fwd<128>(*a, st);
    }
    return set_error(HN_E_UNSUPPORTED, "hn_composite_fwd: n_samples must be 32, 64 or 128");
}

extern "C" int hn_composite_bwd(const hn_composite_bwd_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->feat || !a->sigma || !a->delta || !a->gF || !a->g_bg || !a->dsigma || a->n_rays_total <= 0)
        return set_error(HN_E_BADARG, "hn_composite_bwd: null pointer or empty problem");
    if (a->dfeat_image && (!a->grad_scale || ((int64_t)a->n_rays_total * a->n_samples) % 128 != 0))
        return set_error(HN_E_BADARG, "hn_composite_bwd: image output needs grad_scale and M % 128 == 0");
    cudaStream_t st = (cudaStream_t)stream;
    switch (a->n_samples) {
        case 32: return launch_bwd<32>(*a, st);
        case 64: return launch_bwd<64>(*a, st);
        case 128: return launch_bwd<128>(*a, st);
    }
    return set_error(HN_E_UNSUPPORTED, "hn_composite_bwd: n_samples must be 32, 64 or 128");
}
