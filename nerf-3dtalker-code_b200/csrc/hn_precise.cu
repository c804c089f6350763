// hn_precise.cu — the "high" precision mode of the fg_CD_predictor path for sm_100a: every GEMM of the chain runs on
// the tensor cores with SPLIT operands, x = hi + lo and w = hi + lo (two half-precision numbers each, ~22 significant
// bits), as three tcgen05 products  hi*hi + lo*hi + hi*lo  accumulated in fp32 tensor memory.  The fused single-pass
// kernels (hn_mlp_fwd.cu / hn_mlp_bwd.cu) keep 11 significant bits per operand, which bounds the feature error at
// ~5e-4 x activation scale and the (ill-conditioned, 2^9-gain) camera gradients at cosine ~0.998; this
// mode removes both limits at ~3x the tensor work plus fp32 activations in HBM.
//
// Layer by layer, activations stay fp32 row-major [M, width] in HBM:
//   pe_kernel        ray -> stratified sample -> positional encoding rows [M,64]      (NetWorks/utils.py:43-51,147-161)
//   dense3x_kernel   Y = epi(sum_seg X_seg * W_seg^T): persistent CTA per SM, 128-sample tiles.  Eight worker warps load
//                    the fp32 A block, split it into hi/lo operand images in shared memory (double buffered), and later
//                    drain the accumulator (TMEM -> swizzled staging -> coalesced fp32 rows with bias / ReLU / ReLU mask /
//                    density rank-1 term applied); one thread bulk-copies pre-split weight units (hi|lo, 32 KiB) through
//                    a four-stage ring; one elected thread issues the 12 MMAs of every (K block, 128-column chunk).
//   density_kernel   sigma = relu(w_d . h7 + b_d) in fp32                                (NetWorks/models.py:78,83)
//   pe_bwd_kernel    dL/dPE -> dL/dpts -> per-ray sums for the camera chain
//   colsum_kernel    per-item column sums of the pre-activation gradients = bias / latent-code gradients (A4)
//   to_image_kernel  fp32 rows -> half-precision operand images, so that the weight-gradient kernel (hn_mlp_wgrad.cu)
//                    can consume this mode's activations and gradients unchanged
// Reference semantics: NetWorks/models.py:62-87 and its autograd.
#include <mutex>
#include <vector>
#include "hn_api.h"
#include "hn_mlp_common.cuh"
#include "hn_mlp_sched.h"
#include "hn_sample.cuh"

namespace hn {

constexpr int kPAStages = 2, kPWStages = 4;
constexpr uint32_t kPStageBytes = 2 * kUnitBytes;            // hi image | lo image
constexpr uint32_t kPOffA = 0;
constexpr uint32_t kPOffW = kPAStages * kPStageBytes;
constexpr uint32_t kPSmem = kPOffW + kPWStages * kPStageBytes + 1024;   // + slack for the 1 KiB alignment of the window
constexpr int kPWorkerWarps = 8;
constexpr int kPWorkers = kPWorkerWarps * 32;
constexpr int kPThreads = (2 + kPWorkerWarps) * 32;

// float offsets (per sample) of the fp32 activation / gradient workspaces: buffer = base + M * offset, row stride = width
constexpr int kActPE = 0, kActH0 = 64, kActR0 = 64 + 8 * 384, kActX = kActR0 + 384, kActFloats = kActX + 192;   // 3712
constexpr int kGzZ0 = 0, kGzR0 = 8 * 384, kGzR1 = kGzR0 + 384, kGzPE = kGzR1 + 192, kGzFloats = kGzPE + 64;     // 3712

struct DenseSeg { const float* x; int ld; int nkb; };
struct DenseArgs {
    DenseSeg seg[2];
    int n_seg;
    const uint8_t* w;          // weight units of this op: for every K block, for every chunk: hi unit, lo unit
    int n_chunks;
    int chunk_n[3];
    float* y;
    int ldy;
    const float* bias;         // per-item effective bias row (+ offset of this layer), or NULL
    int bias_stride;
    int relu;
    const float* mask;         // ReLU mask source (saved post-ReLU activation, > 0 passes), or NULL
    int ld_mask;
    const float* sigma;        // density head term: y += [sigma > 0] * dsigma * scale * w_density[col], or w_density NULL
    const float* dsigma;
    const float* w_density;
    const float* scale;
    const float* a_scale;      // device scalar multiplying the A operand (loss scale on dL/dfeat), or NULL
    int n_tiles, tiles_per_item;
    int* status;
};

struct PShared {
    uint64_t a_full[kPAStages], a_empty[kPAStages], w_full[kPWStages], w_empty[kPWStages], acc_full, acc_empty;
    uint32_t tmem_base;
    volatile int abort;
};

// two floats -> (hi, lo) packed half pairs with hi + lo ~ the inputs to ~22 bits
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_sat(a, b);
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    lo = pack_sat(a - f.x, b - f.y);
}

__global__ void __launch_bounds__(kPThreads, 1) dense3x_kernel(const DenseArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ PShared sh;
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < kPAStages; ++i) { mbar_init(smem_u32(&sh.a_full[i]), kPWorkerWarps); mbar_init(smem_u32(&sh.a_empty[i]), 1); }
        for (int i = 0; i < kPWStages; ++i) { mbar_init(smem_u32(&sh.w_full[i]), 1); mbar_init(smem_u32(&sh.w_empty[i]), 1); }
        mbar_init(smem_u32(&sh.acc_full), 1);
        mbar_init(smem_u32(&sh.acc_empty), kPWorkerWarps);
        sh.abort = 0;
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(&sh.tmem_base));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = sh.tmem_base;
    const int total_kb = a.seg[0].nkb + (a.n_seg > 1 ? a.seg[1].nkb : 0);
    const int n_chunks = a.n_chunks;

    if (warp == 0) {
        // ======================= weight producer =======================
        if (lane == 0) {
            uint32_t wc = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles && !sh.abort; tile += gridDim.x) {
                for (int u = 0; u < total_kb * n_chunks; ++u, ++wc) {
                    const uint32_t st = wc % kPWStages, par = (wc / kPWStages) & 1;
                    if (!wait_spin(&sh.w_empty[st], par ^ 1, &sh.abort, a.status, 701)) break;
                    const uint32_t fb = smem_u32(&sh.w_full[st]);
                    mbar_arrive_expect_tx(fb, kPStageBytes);
                    bulk_g2s(smem + kPOffW + st * kPStageBytes, a.w + (size_t)u * kPStageBytes, kPStageBytes, fb);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer (warp-uniform walk, one elected lane issues) =======================
        uint32_t ac = 0, wc = 0, tile_i = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles && !sh.abort; tile += gridDim.x, ++tile_i) {
            bool ok = true;
            if (tile_i > 0) ok = wait_spin(&sh.acc_empty, (tile_i - 1) & 1, &sh.abort, a.status, 710);
            for (int kb = 0; ok && kb < total_kb; ++kb, ++ac) {
                const uint32_t sa = ac % kPAStages;
                ok = wait_spin(&sh.a_full[sa], (ac / kPAStages) & 1, &sh.abort, a.status, 711);
                const uint32_t a_hi = desc_lo(smem + kPOffA + sa * kPStageBytes, 16), a_lo = desc_lo(smem + kPOffA + sa * kPStageBytes + kUnitBytes, 16);
                for (int c = 0; ok && c < n_chunks; ++c, ++wc) {
                    const uint32_t sw = wc % kPWStages;
                    ok = wait_spin(&sh.w_full[sw], (wc / kPWStages) & 1, &sh.abort, a.status, 712);
                    if (!ok) break;
                    tc_fence_after_sync();
                    const uint32_t b_hi = desc_lo(smem + kPOffW + sw * kPStageBytes, 16), b_lo = desc_lo(smem + kPOffW + sw * kPStageBytes + kUnitBytes, 16);
                    const uint32_t idesc = umma_idesc(128, (uint32_t)a.chunk_n[c], kF16, kF16, 0, 0);
                    const uint32_t d = tmem_base + (uint32_t)c * 128;
                    const uint32_t first = (kb == 0) ? 0u : 1u;
                    const bool last_chunk = (c == n_chunks - 1), last_kb = (kb == total_kb - 1);
                    if (elect_one()) {
                        // small products first, the dominant hi*hi last
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_lohi(d, a_lo + ks * 2, b_hi + ks * 2, idesc, ks == 0 ? first : 1u);
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_lohi(d, a_hi + ks * 2, b_lo + ks * 2, idesc, 1u);
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_lohi(d, a_hi + ks * 2, b_hi + ks * 2, idesc, 1u);
                        umma_commit(smem_u32(&sh.w_empty[sw]));
                        if (last_chunk) {
                            umma_commit(smem_u32(&sh.a_empty[sa]));
                            if (last_kb) umma_commit(smem_u32(&sh.acc_full));
                        }
                    }
                    __syncwarp();
                }
            }
            if (!ok) break;
        }
    } else {
        // ======================= workers: A-operand producers, then the epilogue =======================
        const int ww = warp - 2, wt = tid - 64;
        const int quarter = warp & 3, g = ww >> 2;                 // TMEM lane quarter is fixed by the hardware warp id
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const float a_mul = a.a_scale ? __ldg(a.a_scale) : 1.0f;
        const float gscale = a.scale ? __ldg(a.scale) : 1.0f;
        int ncols = 0;
        for (int c = 0; c < n_chunks; ++c) ncols += a.chunk_n[c];
        const int n_pieces = ncols / 32;
        uint32_t ac = 0, tile_i = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles && !sh.abort; tile += gridDim.x, ++tile_i) {
            bool ok = true;
            for (int kb = 0; ok && kb < total_kb; ++kb, ++ac) {
                const DenseSeg& sg = (kb < a.seg[0].nkb) ? a.seg[0] : a.seg[1];
                const int kl = (kb < a.seg[0].nkb) ? kb : kb - a.seg[0].nkb;
                const float* xb = sg.x + (size_t)tile * HN_TILE * sg.ld + kl * 64;
                float4 u[4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j) {                      // global loads first: they do not depend on the stage
                    const int i = wt + kPWorkers * j, r = i >> 3, c8 = i & 7;
                    const float4* p = reinterpret_cast<const float4*>(xb + (size_t)r * sg.ld + c8 * 8);
                    u[j][0] = __ldg(p); u[j][1] = __ldg(p + 1);
                }
                const uint32_t sa = ac % kPAStages;
                ok = wait_spin(&sh.a_empty[sa], ((ac / kPAStages) & 1) ^ 1, &sh.abort, a.status, 720);
                if (!ok) break;                                     // (leaves the K loop only)
                const uint32_t base = smem + kPOffA + sa * kPStageBytes;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = wt + kPWorkers * j, r = i >> 3, c8 = i & 7;
                    const float v[8] = {u[j][0].x * a_mul, u[j][0].y * a_mul, u[j][0].z * a_mul, u[j][0].w * a_mul,
                                        u[j][1].x * a_mul, u[j][1].y * a_mul, u[j][1].z * a_mul, u[j][1].w * a_mul};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int h = 0; h < 4; ++h) split2(v[2 * h], v[2 * h + 1], hi[h], lo[h]);
                    const uint32_t off = image_offset((uint32_t)r, (uint32_t)c8 * 8);
                    st_shared_v4(base + off, hi[0], hi[1], hi[2], hi[3]);
                    st_shared_v4(base + kUnitBytes + off, lo[0], lo[1], lo[2], lo[3]);
                }
                fence_async_smem();
                warp_arrive(smem_u32(&sh.a_full[sa]), lane);
            }
            // ---- epilogue: every MMA of the tile has retired, so the A stages double as staging space
            if (ok) ok = wait_spin(&sh.acc_full, tile_i & 1, &sh.abort, a.status, 730);
            tc_fence_after_sync();
            const int b = tile / a.tiles_per_item;
            const uint32_t stg = smem + kPOffA + (uint32_t)ww * 4096;
            for (int p = g; ok && p < n_pieces; p += 2) {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + 32 * p, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    st_shared_v4(stg + lane * 128 + ((i ^ (lane & 7)) << 4), v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int rr = 4 * j + (lane >> 3), ch = lane & 7;
                    const uint4 raw = ld_shared_v4(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
                    float4 o = make_float4(__uint_as_float(raw.x), __uint_as_float(raw.y), __uint_as_float(raw.z), __uint_as_float(raw.w));
                    const size_t m = (size_t)tile * HN_TILE + quarter * 32 + rr;
                    const int col = 32 * p + 4 * ch;
                    if (a.bias) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + (size_t)b * a.bias_stride + col));
                        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                    }
                    if (a.w_density) {
                        const float s = (__ldg(a.sigma + m) > 0.f) ? __ldg(a.dsigma + m) * gscale : 0.f;
                        const float4 wd = __ldg(reinterpret_cast<const float4*>(a.w_density + col));
                        o.x = fmaf(s, wd.x, o.x); o.y = fmaf(s, wd.y, o.y); o.z = fmaf(s, wd.z, o.z); o.w = fmaf(s, wd.w, o.w);
                    }
                    if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (a.mask) {
                        const float4 h = __ldg(reinterpret_cast<const float4*>(a.mask + m * a.ld_mask + col));
                        o.x = h.x > 0.f ? o.x : 0.f; o.y = h.y > 0.f ? o.y : 0.f; o.z = h.z > 0.f ? o.z : 0.f; o.w = h.w > 0.f ? o.w : 0.f;
                    }
                    *reinterpret_cast<float4*>(a.y + m * a.ldy + col) = o;
                }
                __syncwarp();
            }
            tc_fence_before_sync();
            warp_arrive(smem_u32(&sh.acc_empty), lane);
            named_sync(1, kPWorkers);                               // staging space becomes A stages again (never skipped: a
            if (!ok) break;                                         // faulting warp must not strand the others at the barrier)
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_free<512>(tmem_base);
}

// ----------------------------------------------------------------------------------------------- small SIMT kernels
// positional encoding rows (fp32), z_dists, zvals: one thread per sample
__global__ void __launch_bounds__(128) pe_kernel(hn_camera_t cam, float* pe, float* delta, float* zvals) {
    const int64_t M = (int64_t)cam.B * cam.n_rays * cam.n_samples;
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int s = (int)(m % cam.n_samples);
    const int64_t ray_idx = m / cam.n_samples;
    const int r = (int)(ray_idx % cam.n_rays), b = (int)(ray_idx / cam.n_rays);
    const Ray ray = make_ray(cam, b, r);
    const Sample q = make_sample(cam, ray, b, r, s);
    delta[m] = q.zdist;
    if (zvals) zvals[m] = q.zval;
    const float p[3] = {q.px, q.py, q.pz};
    float v[64];
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[63] = 0.f;
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) sincosf(p[d] * (float)(1 << k), &v[3 + 6 * k + d], &v[3 + 6 * k + 3 + d]);
    float4* dst = reinterpret_cast<float4*>(pe + m * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// sigma = relu(w_d . h7 + b_d): one warp per sample
__global__ void __launch_bounds__(256) density_kernel(const float* h7, const float* w_density, const float* bias, int bias_stride,
                                                      int samples_per_item, int64_t M, float* sigma) {
    const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const float4* row = reinterpret_cast<const float4*>(h7 + m * HN_HIDDEN);
    const float4* w = reinterpret_cast<const float4*>(w_density);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float4 x = __ldg(row + lane + 32 * i), ww = __ldg(w + lane + 32 * i);
        acc = fmaf(x.x, ww.x, acc); acc = fmaf(x.y, ww.y, acc); acc = fmaf(x.z, ww.z, acc); acc = fmaf(x.w, ww.w, acc);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) sigma[m] = fmaxf(acc + __ldg(bias + (size_t)(m / samples_per_item) * bias_stride + HN_BIAS_OFF_DENSITY), 0.f);
}

// dL/dPE [M,64] (loss-scaled) -> dL/dpts -> per-ray sums; one thread per sample, a warp never straddles rays
__global__ void __launch_bounds__(128) pe_bwd_kernel(hn_camera_t cam, const float* dpe, const float* ddelta, const float* scale,
                                                     float* g_ray_o, float* g_ray_v, float* g_ray_l) {
    const int64_t M = (int64_t)cam.B * cam.n_rays * cam.n_samples;
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int lane = threadIdx.x & 31;
    const float inv_scale = 1.0f / __ldg(scale);
    const int ns = cam.n_samples;
    const int64_t ray_idx = m / ns;
    const int s = (int)(m % ns), r = (int)(ray_idx % cam.n_rays), b = (int)(ray_idx / cam.n_rays);
    const Ray ray = make_ray(cam, b, r);
    const float e0 = sample_edge(cam, ray.oz, b, r, s), e1 = sample_edge(cam, ray.oz, b, r, s + 1);
    const float p[3] = {__fadd_rn(ray.ox, __fmul_rn(ray.vx, e0)), __fadd_rn(ray.oy, __fmul_rn(ray.vy, e0)),
                        __fadd_rn(ray.oz, __fmul_rn(ray.vz, e0))};
    float gq[64];
    const float4* src = reinterpret_cast<const float4*>(dpe + m * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float4 t = __ldg(src + i); gq[4 * i] = t.x; gq[4 * i + 1] = t.y; gq[4 * i + 2] = t.z; gq[4 * i + 3] = t.w; }
    float dp[3] = {gq[0], gq[1], gq[2]};
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float f = (float)(1 << k);
            float sn, cs;
            sincosf(p[d] * f, &sn, &cs);
            dp[d] = fmaf(f, gq[3 + 6 * k + d] * cs - gq[3 + 6 * k + 3 + d] * sn, dp[d]);
        }
#pragma unroll
    for (int d = 0; d < 3; ++d) dp[d] *= inv_scale;
    const float dpv = dp[0] * ray.vx + dp[1] * ray.vy + dp[2] * ray.vz;                 // through z_s = o_z - const
    float red[7] = {dp[0], dp[1], dp[2] + dpv, e0 * dp[0], e0 * dp[1], e0 * dp[2], ddelta ? (e1 - e0) * __ldg(ddelta + m) : 0.f};
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) red[i] += __shfl_xor_sync(0xffffffffu, red[i], d);
    if (lane == 0) {
        atomicAdd(g_ray_o + ray_idx * 3 + 0, red[0]); atomicAdd(g_ray_o + ray_idx * 3 + 1, red[1]); atomicAdd(g_ray_o + ray_idx * 3 + 2, red[2]);
        atomicAdd(g_ray_v + ray_idx * 3 + 0, red[3]); atomicAdd(g_ray_v + ray_idx * 3 + 1, red[4]); atomicAdd(g_ray_v + ray_idx * 3 + 2, red[5]);
        atomicAdd(g_ray_l + ray_idx, red[6]);
    }
}

// dbias[item, off + c] += mult * sum over the item's rows of src[row, c]; grid (tiles of 128 rows, column blocks of 128)
__global__ void __launch_bounds__(128) colsum_kernel(const float* src, int ld, int ncols, int rows_per_item, const float* scale, int unscale,
                                                     float* dbias, int bias_stride, int off) {
    const int c = blockIdx.y * 128 + threadIdx.x;
    if (c >= ncols) return;
    const int64_t row0 = (int64_t)blockIdx.x * HN_TILE;
    float acc = 0.f;
#pragma unroll 8
    for (int i = 0; i < HN_TILE; ++i) acc += __ldg(src + (row0 + i) * ld + c);
    const float mult = unscale ? 1.0f / __ldg(scale) : 1.0f;
    atomicAdd(dbias + (size_t)(row0 / rows_per_item) * bias_stride + off + c, acc * mult);
}
// density head: d(bias_density) = sum over samples of [sigma > 0] dsigma
__global__ void __launch_bounds__(128) dens_colsum_kernel(const float* sigma, const float* dsigma, int rows_per_item, float* dbias, int bias_stride) {
    const int64_t row0 = (int64_t)blockIdx.x * HN_TILE;
    const int64_t m = row0 + threadIdx.x;
    float v = (__ldg(sigma + m) > 0.f) ? __ldg(dsigma + m) : 0.f;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(dbias + (size_t)(row0 / rows_per_item) * bias_stride + HN_BIAS_OFF_DENSITY, v);
}

// fp32 rows -> half-precision operand image blocks (slot + kb) for the weight-gradient kernel; grid (n_tiles, n_blocks)
__global__ void __launch_bounds__(256) to_image_kernel(const float* src, int ld, uint8_t* image, int slot, int n_tiles, const float* scale) {
    const int tile = blockIdx.x, kb = blockIdx.y;
    const float mul = scale ? __ldg(scale) : 1.0f;
    uint8_t* dst = image + ((size_t)(slot + kb) * n_tiles + tile) * kUnitBytes;
    const float* xb = src + (size_t)tile * HN_TILE * ld + kb * 64;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = threadIdx.x + 256 * j, r = i >> 3, c8 = i & 7;
        const float4* p = reinterpret_cast<const float4*>(xb + (size_t)r * ld + c8 * 8);
        const float4 u0 = __ldg(p), u1 = __ldg(p + 1);
        *reinterpret_cast<uint4*>(dst + image_offset((uint32_t)r, (uint32_t)c8 * 8)) =
            make_uint4(pack_sat(u0.x * mul, u0.y * mul), pack_sat(u0.z * mul, u0.w * mul), pack_sat(u1.x * mul, u1.y * mul), pack_sat(u1.z * mul, u1.w * mul));
    }
}
// density head as a one-channel pseudo layer: block rows = [dsr * scale, 0, ..., 0]
__global__ void __launch_bounds__(128) dens_image_kernel(const float* sigma, const float* dsigma, const float* scale, uint8_t* image, int slot, int n_tiles) {
    const int tile = blockIdx.x, r = threadIdx.x;
    const size_t m = (size_t)tile * HN_TILE + r;
    const float dsr = (__ldg(sigma + m) > 0.f) ? __ldg(dsigma + m) * __ldg(scale) : 0.f;
    uint8_t* row = image + ((size_t)slot * n_tiles + tile) * kUnitBytes + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4)) = make_uint4(c == 0 ? pack_sat(dsr, 0.f) : 0u, 0u, 0u, 0u);
}

// ----------------------------------------------------------------------------------------------- weight units (hi | lo)
struct PackPArgs { const float* w[12]; int ld[12]; int l5_hidden_col; };

// one CTA per (unit, part): part 0 = hi image, part 1 = lo image (residual of the half-precision rounding)
__global__ void __launch_bounds__(256) pack_precise_kernel(PackPArgs a, const PackOp* ops, uint8_t* packed) {
    const int u = blockIdx.x >> 1, part = blockIdx.x & 1;
    const PackOp op = ops[u];
    const float* W = a.w[op.w_idx];
    const int ld = a.ld[op.w_idx];
    const int col0 = op.col0 + (op.l5_hidden ? a.l5_hidden_col : 0);
    uint8_t* dst = packed + ((size_t)u * 2 + part) * kUnitBytes;
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int r = i >> 3, c8 = (i & 7) * 8;
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c8 + 2 * h + e;
                float x = 0.f;
                if (r < op.valid_r && c < op.valid_c)
                    x = op.transposed ? __ldg(W + (size_t)(op.row0 + c) * ld + col0 + r) : __ldg(W + (size_t)(op.row0 + r) * ld + col0 + c);
                v[e] = x;
            }
            uint32_t hi, lo;
            split2(v[0], v[1], hi, lo);
            pk[h] = part ? lo : hi;
        }
        *reinterpret_cast<uint4*>(dst + image_offset((uint32_t)r, (uint32_t)c8) ) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// ----------------------------------------------------------------------------------------------- host-side op list
namespace {

enum Buf { B_PE = 0, B_H0 = 1, /* .. B_H7 = 8 */ B_R0 = 9, B_X = 10, B_FEAT = 11, B_DFEAT = 12, B_Z0 = 13, /* .. B_Z7 = 20 */ B_ZR0 = 21, B_ZR1 = 22, B_DPE = 23 };

struct SegDef { int src; int nkb; int w_idx; int wcol0; bool l5h; int kvalid; };
struct OpDef {
    int n_seg; SegDef seg[2];
    bool transposed;
    int n;              // MMA N total (multiple of 64)
    int nvalid;         // valid weight rows along N
    int dst;
    int bias_off;       // -1 = none
    bool relu;
    int mask_src;       // -1 = none
    bool density, a_scale;
    int unit_base;      // first (hi|lo) unit pair of this op in the packed stream
};

struct PreciseSchedule {
    std::vector<OpDef> fwd, bwd, bwd_pe;     // bwd_pe: the dL/dPE op (camera gradients only)
    std::vector<PackOp> pack;                // one per unit pair
};

int chunk_of(int n, int c) { return std::min(128, n - 128 * c); }
int chunks_of(int n) { return (n + 127) / 128; }

void add_op(PreciseSchedule& S, std::vector<OpDef>& list, OpDef op) {
    op.unit_base = (int)S.pack.size();
    for (int s = 0; s < op.n_seg; ++s) {
        const SegDef& sg = op.seg[s];
        for (int kb = 0; kb < sg.nkb; ++kb)
            for (int c = 0; c < chunks_of(op.n); ++c) {
                PackOp p{};
                p.w_idx = (int8_t)sg.w_idx;
                p.l5_hidden = (int8_t)sg.l5h;
                if (!op.transposed) {            // unit(r,c) = W[(row0 + r) * ld + col0 + c]: rows = outputs, cols = inputs (K)
                    p.transposed = 0;
                    p.row0 = (int16_t)(128 * c);
                    p.valid_r = (int16_t)std::min(chunk_of(op.n, c), op.nvalid - 128 * c);
                    p.col0 = (int16_t)(sg.wcol0 + 64 * kb);
                    p.valid_c = (int16_t)std::min(64, sg.kvalid - 64 * kb);
                } else {                         // unit(r,c) = W[(row0 + c) * ld + col0 + r]: rows = inputs (N), cols = outputs (K)
                    p.transposed = 1;
                    p.row0 = (int16_t)(64 * kb);
                    p.valid_c = (int16_t)std::min(64, sg.kvalid - 64 * kb);
                    p.col0 = (int16_t)(sg.wcol0 + 128 * c);
                    p.valid_r = (int16_t)std::min(chunk_of(op.n, c), op.nvalid - 128 * c);
                }
                S.pack.push_back(p);
            }
    }
    list.push_back(op);
}

const PreciseSchedule& precise_schedule() {
    static PreciseSchedule S;
    static std::once_flag once;
    std::call_once(once, [] {
        auto fop = [](int nseg, SegDef s0, SegDef s1, int n, int dst, int bias_off, bool relu) {
            OpDef o{}; o.n_seg = nseg; o.seg[0] = s0; o.seg[1] = s1; o.transposed = false; o.n = n; o.nvalid = n; o.dst = dst;
            o.bias_off = bias_off; o.relu = relu; o.mask_src = -1; o.density = false; o.a_scale = false; return o;
        };
        const SegDef none{};
        // forward chain (NetWorks/models.py:62-87)
        add_op(S, S.fwd, fop(1, SegDef{B_PE, 1, W_L0, 0, false, HN_PE}, none, 384, B_H0, 0, true));
        for (int i = 1; i < 8; ++i) {
            if (i == 5) add_op(S, S.fwd, fop(2, SegDef{B_PE, 1, W_L5, 0, false, HN_PE}, SegDef{B_H0 + 4, 6, W_L5, 0, true, 384}, 384, B_H0 + 5, 384 * 5, true));
            else add_op(S, S.fwd, fop(1, SegDef{B_H0 + i - 1, 6, i, 0, false, 384}, none, 384, B_H0 + i, 384 * i, true));
        }
        add_op(S, S.fwd, fop(1, SegDef{B_H0 + 7, 6, W_R0, 0, false, 384}, none, 384, B_R0, HN_BIAS_OFF_R0, false));
        add_op(S, S.fwd, fop(1, SegDef{B_R0, 6, W_R1, 0, false, 384}, none, 192, B_X, HN_BIAS_OFF_R1, true));
        add_op(S, S.fwd, fop(1, SegDef{B_X, 3, W_R2, 0, false, 192}, none, 256, B_FEAT, HN_BIAS_OFF_R2, false));
        // data-gradient chain: dX = mask * (dZ * W (+ density term)); seg.kvalid = output channels of the layer (contraction)
        auto bop = [](SegDef s0, int n, int nvalid, int dst, int mask_src, bool density, bool a_scale) {
            OpDef o{}; o.n_seg = 1; o.seg[0] = s0; o.transposed = true; o.n = n; o.nvalid = nvalid; o.dst = dst; o.bias_off = -1;
            o.relu = false; o.mask_src = mask_src; o.density = density; o.a_scale = a_scale; return o;
        };
        add_op(S, S.bwd, bop(SegDef{B_DFEAT, 4, W_R2, 0, false, 256}, 192, 192, B_ZR1, B_X, false, true));
        add_op(S, S.bwd, bop(SegDef{B_ZR1, 3, W_R1, 0, false, 192}, 384, 384, B_ZR0, -1, false, false));
        add_op(S, S.bwd, bop(SegDef{B_ZR0, 6, W_R0, 0, false, 384}, 384, 384, B_Z0 + 7, B_H0 + 7, true, false));
        for (int i = 7; i >= 1; --i)
            add_op(S, S.bwd, bop(SegDef{B_Z0 + i, 6, i, 0, i == 5, 384}, 384, 384, B_Z0 + i - 1, B_H0 + i - 1, false, false));
        OpDef pe = bop(SegDef{B_Z0 + 5, 6, W_L5, 0, false, 384}, 64, HN_PE, B_DPE, -1, false, false);
        pe.n_seg = 2; pe.seg[1] = SegDef{B_Z0, 6, W_L0, 0, false, 384};
        add_op(S, S.bwd_pe, pe);
    });
    return S;
}

std::mutex g_pp_mu;
PackOp* g_pack_dev[64] = {};

struct BufTable { const float* p[24]; int ld[24]; };

int launch_dense(const OpDef& op, const BufTable& bt, const uint8_t* packed, const float* bias, const float* sigma, const float* dsigma,
                 const float* w_density, const float* scale, int n_tiles, int tiles_per_item, int* status, int n_sm, cudaStream_t st) {
    DenseArgs a{};
    a.n_seg = op.n_seg;
    for (int s = 0; s < op.n_seg; ++s) a.seg[s] = DenseSeg{bt.p[op.seg[s].src], bt.ld[op.seg[s].src], op.seg[s].nkb};
    a.w = packed + (size_t)op.unit_base * kPStageBytes;
    a.n_chunks = chunks_of(op.n);
    for (int c = 0; c < a.n_chunks; ++c) a.chunk_n[c] = chunk_of(op.n, c);
    a.y = const_cast<float*>(bt.p[op.dst]);
    a.ldy = bt.ld[op.dst];
    a.bias = (op.bias_off >= 0) ? bias + op.bias_off : nullptr;
    a.bias_stride = HN_BIAS_STRIDE;
    a.relu = op.relu ? 1 : 0;
    a.mask = (op.mask_src >= 0) ? bt.p[op.mask_src] : nullptr;
    a.ld_mask = (op.mask_src >= 0) ? bt.ld[op.mask_src] : 0;
    if (op.density) { a.sigma = sigma; a.dsigma = dsigma; a.w_density = w_density; }
    a.scale = scale;
    a.a_scale = op.a_scale ? scale : nullptr;
    a.n_tiles = n_tiles;
    a.tiles_per_item = tiles_per_item;
    a.status = status;
    dense3x_kernel<<<n_tiles < n_sm ? n_tiles : n_sm, kPThreads, kPSmem, st>>>(a);
    return check_launch("hn precise dense layer");
}

void fill_act_bufs(BufTable& bt, const float* acts, int64_t M) {
    bt.p[B_PE] = acts + M * kActPE; bt.ld[B_PE] = 64;
    for (int i = 0; i < 8; ++i) { bt.p[B_H0 + i] = acts + M * (kActH0 + 384 * i); bt.ld[B_H0 + i] = 384; }
    bt.p[B_R0] = acts + M * kActR0; bt.ld[B_R0] = 384;
    bt.p[B_X] = acts + M * kActX; bt.ld[B_X] = 192;
}
void fill_gz_bufs(BufTable& bt, const float* gz, int64_t M) {
    for (int i = 0; i < 8; ++i) { bt.p[B_Z0 + i] = gz + M * (kGzZ0 + 384 * i); bt.ld[B_Z0 + i] = 384; }
    bt.p[B_ZR0] = gz + M * kGzR0; bt.ld[B_ZR0] = 384;
    bt.p[B_ZR1] = gz + M * kGzR1; bt.ld[B_ZR1] = 192;
    bt.p[B_DPE] = gz + M * kGzPE; bt.ld[B_DPE] = 64;
}

int prepare_device(int* n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    static bool ready[64] = {};
    {
        std::lock_guard<std::mutex> lk(g_pp_mu);
        if (dev < 64 && !ready[dev]) {
            cudaError_t e = cudaFuncSetAttribute(dense3x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmem);
            if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
            ready[dev] = true;
        }
    }
    *n_sm = 148;
    cudaDeviceGetAttribute(n_sm, cudaDevAttrMultiProcessorCount, dev);
    return HN_OK;
}

}  // namespace
}  // namespace hn

extern "C" size_t hn_precise_packed_bytes(void) { return hn::precise_schedule().pack.size() * (size_t)hn::kPStageBytes; }
extern "C" size_t hn_precise_workspace_floats(int64_t M) { return (size_t)hn::kActFloats * (size_t)M; }

extern "C" int hn_pack_weights_precise(const hn_weights_t* w, void* packed, void* stream) {
    using namespace hn;
    if (!w || !packed) return set_error(HN_E_BADARG, "hn_pack_weights_precise: null pointer");
    for (int i = 0; i < 12; ++i)
        if (!w->w[i] || w->ld[i] <= 0) return set_error(HN_E_BADARG, "hn_pack_weights_precise: null weight pointer or bad leading dimension");
    if (w->ld[W_L0] < HN_PE + 1 || w->l5_hidden_col < HN_PE || w->l5_hidden_col + HN_HIDDEN > w->ld[W_L5] ||
        w->ld[W_R1] < HN_HIDDEN || w->ld[W_R2] != HN_RGB1)
        return set_error(HN_E_UNSUPPORTED, "hn_pack_weights_precise: layer shapes do not match fg_CD_predictor (hidden 384, feat 256)");
    const PreciseSchedule& S = precise_schedule();
    int dev = 0;
    cudaGetDevice(&dev);
    PackOp* ops_dev = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_pp_mu);
        if (dev >= 64) return set_error(HN_E_UNSUPPORTED, "hn_pack_weights_precise: device index >= 64");
        if (!g_pack_dev[dev]) {
            // the only allocation of the library: a < 8 KiB immutable table per device, made once
            cudaError_t e = cudaMalloc(&g_pack_dev[dev], S.pack.size() * sizeof(PackOp));
            if (e == cudaSuccess) e = cudaMemcpy(g_pack_dev[dev], S.pack.data(), S.pack.size() * sizeof(PackOp), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        }
        ops_dev = g_pack_dev[dev];
    }
    PackPArgs a;
    for (int i = 0; i < 12; ++i) { a.w[i] = w->w[i]; a.ld[i] = w->ld[i]; }
    a.l5_hidden_col = w->l5_hidden_col;
    pack_precise_kernel<<<(unsigned)(2 * S.pack.size()), 256, 0, (cudaStream_t)stream>>>(a, ops_dev, (uint8_t*)packed);
    return check_launch("hn_pack_weights_precise");
}

extern "C" int hn_mlp_fwd_precise(const hn_mlp_fwd_precise_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->cam.xy || !a->cam.Rmats || !a->cam.Tvecs || !a->cam.inv_inmats || !a->bias || !a->w_density ||
        !a->packed_hl || !a->feat || !a->sigma || !a->delta || !a->acts || !a->status)
        return set_error(HN_E_BADARG, "hn_mlp_fwd_precise: null pointer");
    if ((((uintptr_t)a->w_density | (uintptr_t)a->bias | (uintptr_t)a->feat | (uintptr_t)a->acts) & 15) != 0)
        return set_error(HN_E_BADARG, "hn_mlp_fwd_precise: w_density, bias, feat and acts must be 16-byte aligned (vector loads)");
    if (int rc = check_geometry(a->cam.B, a->cam.n_rays, a->cam.n_samples, "hn_mlp_fwd_precise")) return rc;
    int n_sm = 148;
    if (int rc = prepare_device(&n_sm)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M = total_samples(a->cam.B, a->cam.n_rays, a->cam.n_samples);
    const int n_tiles = (int)(M / HN_TILE);
    const int tiles_per_item = (int)(((int64_t)a->cam.n_rays * a->cam.n_samples) / HN_TILE);
    BufTable bt{};
    fill_act_bufs(bt, a->acts, M);
    bt.p[B_FEAT] = a->feat; bt.ld[B_FEAT] = HN_FEAT;
    pe_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(a->cam, a->acts + M * kActPE, a->delta, a->zvals);
    if (int rc = check_launch("hn_mlp_fwd_precise (positional encoding)")) return rc;
    for (const OpDef& op : precise_schedule().fwd)
        if (int rc = launch_dense(op, bt, (const uint8_t*)a->packed_hl, a->bias, nullptr, nullptr, nullptr, nullptr, n_tiles, tiles_per_item, a->status, n_sm, st)) return rc;
    density_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, st>>>(bt.p[B_H0 + 7], a->w_density, a->bias, HN_BIAS_STRIDE,
                                                                     a->cam.n_rays * a->cam.n_samples, M, a->sigma);
    return check_launch("hn_mlp_fwd_precise (density head)");
}

extern "C" int hn_mlp_bwd_data_precise(const hn_mlp_bwd_data_precise_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->cam.xy || !a->cam.Rmats || !a->cam.Tvecs || !a->cam.inv_inmats || !a->packed_hl || !a->w_density ||
        !a->dfeat || !a->dsigma || !a->sigma || !a->grad_scale || !a->acts || !a->gz || !a->status)
        return set_error(HN_E_BADARG, "hn_mlp_bwd_data_precise: null pointer");
    if ((((uintptr_t)a->w_density | (uintptr_t)a->dfeat | (uintptr_t)a->acts | (uintptr_t)a->gz) & 15) != 0)
        return set_error(HN_E_BADARG, "hn_mlp_bwd_data_precise: w_density, dfeat, acts and gz must be 16-byte aligned (vector loads)");
    if (int rc = check_geometry(a->cam.B, a->cam.n_rays, a->cam.n_samples, "hn_mlp_bwd_data_precise")) return rc;
    const bool with_pe = (a->g_ray_o != nullptr);
    if (with_pe && (!a->g_ray_v || !a->g_ray_l))
        return set_error(HN_E_BADARG, "hn_mlp_bwd_data_precise: g_ray_o, g_ray_v and g_ray_l must be given together");
    const bool images = (a->act_image != nullptr);
    if (images && (!a->grads_image || !a->dfeat_image))
        return set_error(HN_E_BADARG, "hn_mlp_bwd_data_precise: act_image, grads_image and dfeat_image must be given together");
    int n_sm = 148;
    if (int rc = prepare_device(&n_sm)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M = total_samples(a->cam.B, a->cam.n_rays, a->cam.n_samples);
    const int n_tiles = (int)(M / HN_TILE);
    const int rows_per_item = a->cam.n_rays * a->cam.n_samples;
    const int tiles_per_item = rows_per_item / HN_TILE;
    BufTable bt{};
    fill_act_bufs(bt, a->acts, M);
    fill_gz_bufs(bt, a->gz, M);
    bt.p[B_DFEAT] = a->dfeat; bt.ld[B_DFEAT] = HN_FEAT;
    const PreciseSchedule& S = precise_schedule();
    for (const OpDef& op : S.bwd)
        if (int rc = launch_dense(op, bt, (const uint8_t*)a->packed_hl, nullptr, a->sigma, a->dsigma, a->w_density, a->grad_scale, n_tiles, tiles_per_item, a->status, n_sm, st)) return rc;
    if (with_pe) {
        if (int rc = launch_dense(S.bwd_pe[0], bt, (const uint8_t*)a->packed_hl, nullptr, nullptr, nullptr, nullptr, a->grad_scale, n_tiles, tiles_per_item, a->status, n_sm, st)) return rc;
        pe_bwd_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(a->cam, bt.p[B_DPE], a->ddelta, a->grad_scale, a->g_ray_o, a->g_ray_v, a->g_ray_l);
        if (int rc = check_launch("hn_mlp_bwd_data_precise (positional-encoding gradient)")) return rc;
    }
    if (a->dbias) {
        // bias / latent-code gradients: per-item column sums of the pre-activation gradients
        const dim3 g384((unsigned)n_tiles, 3), g192((unsigned)n_tiles, 2), g256((unsigned)n_tiles, 2);
        for (int i = 0; i < 8; ++i)
            colsum_kernel<<<g384, 128, 0, st>>>(bt.p[B_Z0 + i], 384, 384, rows_per_item, a->grad_scale, 1, a->dbias, HN_BIAS_STRIDE, 384 * i);
        colsum_kernel<<<g384, 128, 0, st>>>(bt.p[B_ZR0], 384, 384, rows_per_item, a->grad_scale, 1, a->dbias, HN_BIAS_STRIDE, HN_BIAS_OFF_R0);
        colsum_kernel<<<g192, 128, 0, st>>>(bt.p[B_ZR1], 192, 192, rows_per_item, a->grad_scale, 1, a->dbias, HN_BIAS_STRIDE, HN_BIAS_OFF_R1);
        colsum_kernel<<<g256, 128, 0, st>>>(a->dfeat, HN_FEAT, HN_FEAT, rows_per_item, a->grad_scale, 0, a->dbias, HN_BIAS_STRIDE, HN_BIAS_OFF_R2);
        dens_colsum_kernel<<<(unsigned)n_tiles, 128, 0, st>>>(a->sigma, a->dsigma, rows_per_item, a->dbias, HN_BIAS_STRIDE);
        if (int rc = check_launch("hn_mlp_bwd_data_precise (bias gradients)")) return rc;
    }
    if (images) {
        uint8_t* ai = (uint8_t*)a->act_image;
        uint8_t* gi = (uint8_t*)a->grads_image;
        to_image_kernel<<<dim3(n_tiles, 1), 256, 0, st>>>(bt.p[B_PE], 64, ai, HN_SLOT_PE, n_tiles, nullptr);
        for (int i = 0; i < 8; ++i) {
            to_image_kernel<<<dim3(n_tiles, 6), 256, 0, st>>>(bt.p[B_H0 + i], 384, ai, HN_SLOT_H0 + 6 * i, n_tiles, nullptr);
            to_image_kernel<<<dim3(n_tiles, 6), 256, 0, st>>>(bt.p[B_Z0 + i], 384, gi, HN_GSLOT_Z0 + 6 * i, n_tiles, nullptr);
        }
        to_image_kernel<<<dim3(n_tiles, 6), 256, 0, st>>>(bt.p[B_R0], 384, ai, HN_SLOT_R0, n_tiles, nullptr);
        to_image_kernel<<<dim3(n_tiles, 3), 256, 0, st>>>(bt.p[B_X], 192, ai, HN_SLOT_X, n_tiles, nullptr);
        to_image_kernel<<<dim3(n_tiles, 6), 256, 0, st>>>(bt.p[B_ZR0], 384, gi, HN_GSLOT_R0, n_tiles, nullptr);
        to_image_kernel<<<dim3(n_tiles, 3), 256, 0, st>>>(bt.p[B_ZR1], 192, gi, HN_GSLOT_R1, n_tiles, nullptr);
        dens_image_kernel<<<n_tiles, 128, 0, st>>>(a->sigma, a->dsigma, a->grad_scale, gi, HN_GSLOT_DENS, n_tiles);
        to_image_kernel<<<dim3(n_tiles, 4), 256, 0, st>>>(a->dfeat, HN_FEAT, (uint8_t*)a->dfeat_image, 0, n_tiles, a->grad_scale);
        if (int rc = check_launch("hn_mlp_bwd_data_precise (operand images)")) return rc;
    }
    return HN_OK;
}
