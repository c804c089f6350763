// hn_mlp_wgrad.cu — weight gradients of fg_CD_predictor: dW_l = sum over samples of dZ_l^T x X_l, where dZ_l are
// the (loss-scaled, fp16) pre-activation gradients saved by hn_mlp_bwd_data and X_l the layer inputs saved by
// hn_mlp_fwd — both already stored as tensor-core operand images, consumed here as MN-major operands
// (contraction over the 128 sample rows of a block), so nothing is transposed or re-laid-out.
//
// Work item = (layer, 128-channel chunk of dZ, batch item, sample split).  A CTA accumulates the item's
// [128 x K_in] slice of dW in TMEM over all its tiles, plus a 16-column "ones" product that yields the
// bias gradient (= per-item column sums of dZ, from which the host derives the latent-code and folded
// weight-column gradients), then flushes once with atomic adds.  The density head rides along as a
// one-channel pseudo layer (its gradient block is written by hn_mlp_bwd_data).
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>
#include "hn_api.h"
#include "hn_mlp_sched.h"
#include "hn_tc.cuh"

#ifndef HN_WEXP
#define HN_WEXP 0
#endif

namespace hn {

constexpr int kWStages = 3;
constexpr int kHalfBytes = kUnitBytes / 2;                  // 64 sample rows of a block
constexpr int kWMaxX = 7;
constexpr uint32_t kWStageBytes = (2 + kWMaxX) * kHalfBytes;  // 72 KiB
constexpr uint32_t kWOffOnes = kWStages * kWStageBytes;
constexpr uint32_t kWgradSmem = kWOffOnes + 2048 + 1024;
constexpr int kWThreads = 192;                              // warp 0 producer, warp 1 MMA (+TMEM alloc), warps 2..5 flush
constexpr uint32_t kBiasCol = 448;
#ifndef HN_WPREFETCH
#define HN_WPREFETCH 0
#endif
constexpr int kWPrefetch = HN_WPREFETCH;                    // L2 prefetch distance in 64-sample stages.  0 = off: measured SLOWER with it (2.24 ms at 4 stages, 2.39 at 8,
                                                            // 2.63 at 32, against 2.11 without) - the memory system is already saturated
constexpr int kWProd = 5;                                   // producer lanes (see the producer role)

struct WItem {
    int16_t w_idx;            // destination weight (index into dw[]), -1: none
    int16_t row0;             // first dW row of this chunk
    int16_t rows;             // valid rows (128, 64 or 1)
    int16_t n_x;              // number of X blocks (0: bias-only item)
    int32_t g_blk;            // first block of the dZ chunk (in grads, or in dfeat_image when g_dfeat)
    int16_t g_dfeat;
    int16_t bias_off;         // offset in the bias row, -1: none
    int32_t x_blk[kWMaxX];    // act block ids
    int16_t x_col[kWMaxX];    // dW column of the block's first column
    int16_t x_valid[kWMaxX];  // valid columns (64, or 63 for the PE block)
    int32_t b;                // batch item
    int32_t tile0, tile1;     // tile range [tile0, tile1)
};

struct WShared {
    uint64_t full[kWStages], empty[kWStages], acc_full, acc_empty;
    uint32_t tmem_base;
    volatile int abort;
};

__device__ __forceinline__ bool wwait(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* status, int code) {
    const uint32_t b = smem_u32(bar);
    if (mbar_try_wait(b, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(b, parity)) {
        if (*abort_flag) return false;
        if (clock64() - t0 > 2000000000ll) { *abort_flag = 1; atomicCAS(status, 0, code); return false; }
    }
    return true;
}

struct WArgs {
    const uint8_t* act; const uint8_t* grads; const uint8_t* dfeat_image;
    const float* grad_scale;
    float* dw[12]; int ld[12];
    float* dbias;
    const WItem* items; int n_items; int n_tiles;
    int* status;
};

// CL = 1: one CTA per work item.  CL = 3: a cluster of three CTAs works on the three 128-channel chunks of ONE layer over the
// same samples; every X (layer-input) block is fetched from L2 once per cluster and multicast into all three CTAs' shared
// memory, which halves the L2->SM traffic that bounds this kernel (~35 B/cycle/SM ingest).
template <int CL>
__global__ void __launch_bounds__(kWThreads, 1) mlp_wgrad_kernel(const WArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ WShared sh;
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
    const int item0 = (int)blockIdx.x / CL, item_stride = (int)gridDim.x / CL;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);

    if (tid == 0) {
        for (int i = 0; i < kWStages; ++i) { mbar_init(smem_u32(&sh.full[i]), 1); mbar_init(smem_u32(&sh.empty[i]), CL); }
        mbar_init(smem_u32(&sh.acc_full), 1);
        mbar_init(smem_u32(&sh.acc_empty), 128);
        sh.abort = 0;
        mbar_fence_init();
    }
    // "ones" operand: 16 rows x 64 samples of 1.0h (K-major image; every element equal, so swizzling is moot)
    for (int i = tid; i < 2048 / 4; i += kWThreads)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem + kWOffOnes + i * 4), "r"(0x3C003C00u) : "memory");
    fence_async_smem();
    if (warp == 1) tmem_alloc<512>(smem_u32(&sh.tmem_base));
    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                              // peers' barriers exist before any multicast can signal them
    tc_fence_after_sync();
    const uint32_t tmem_base = sh.tmem_base;

    if (warp == 0) {
        // ======================= producers =======================
        // The bulk copies of ONE thread complete one after the other (~700 cycles apiece under load, whatever their size), and a
        // stage is 4-9 copies of 8 KiB: a lone producer thread caps the kernel at ~2400 cycles per stage.  kWProd lanes walk the
        // same schedule and take the stage's copies round-robin; lane 0 posts the byte count (a copy that lands before that only
        // drives the transaction count negative for a moment: the phase cannot complete before lane 0's arrival).
        if (lane < kWProd) {
            uint32_t sc = 0;
            for (int it = item0; it < a.n_items && !sh.abort; it += item_stride) {
                const WItem w = a.items[it * CL + rank];
                if (w.tile1 <= w.tile0) continue;                  // padding entry of the balanced schedule
                const uint8_t* gsrc = w.g_dfeat ? a.dfeat_image : a.grads;
                const uint32_t bytes = (2 + w.n_x) * kHalfBytes;
                for (int tile = w.tile0; tile < w.tile1; ++tile) {
                    for (int half = 0; half < 2; ++half, ++sc) {
                        const uint32_t stage = sc % kWStages, par = (sc / kWStages) & 1;
                        if (!wwait(&sh.empty[stage], par ^ 1, &sh.abort, a.status, 701)) break;
                        const uint32_t fb = smem_u32(&sh.full[stage]);
#if HN_WEXP == 2 || HN_WEXP == 4                                    // diagnostic builds: no operand loads (MMA path alone)
                        if (lane == 0) mbar_arrive(fb);
                        continue;
#endif
                        if (lane == 0) mbar_arrive_expect_tx(fb, bytes);
                        const uint32_t dst = smem + stage * kWStageBytes;
                        // the same pieces kWPrefetch stages ahead are pulled DRAM -> L2 now: shared memory holds only three stages, far
                        // too few bytes in flight to cover DRAM latency, but L2 has room for dozens
                        const int ahead = 2 * (tile - w.tile0) + half + kWPrefetch;
                        const int pf_tile = w.tile0 + (ahead >> 1), pf_half = ahead & 1;
                        const bool pf = kWPrefetch > 0 && pf_tile < w.tile1;
                        int j = 0;
                        for (int k = 0; k < 2; ++k)
                            if ((j++ % kWProd) == lane) {
                                bulk_g2s(dst + k * kHalfBytes, gsrc + ((size_t)(w.g_blk + k) * a.n_tiles + tile) * kUnitBytes + half * kHalfBytes, kHalfBytes, fb);
                                if (pf) bulk_prefetch_l2(gsrc + ((size_t)(w.g_blk + k) * a.n_tiles + pf_tile) * kUnitBytes + pf_half * kHalfBytes, kHalfBytes);
                            }
                        for (int k = 0; k < w.n_x; ++k) {
                            if (CL > 1 && k % CL != (int)rank) continue;
                            if ((j++ % kWProd) != lane) continue;
                            const uint8_t* xs = a.act + ((size_t)w.x_blk[k] * a.n_tiles + tile) * kUnitBytes + half * kHalfBytes;
                            if (CL == 1) bulk_g2s(dst + (2 + k) * kHalfBytes, xs, kHalfBytes, fb);
                            else bulk_g2s_multicast(dst + (2 + k) * kHalfBytes, xs, kHalfBytes, fb, kMask);
                            if (pf) bulk_prefetch_l2(a.act + ((size_t)w.x_blk[k] * a.n_tiles + pf_tile) * kUnitBytes + pf_half * kHalfBytes, kHalfBytes);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        // One thread issues 12 MMAs per 64-sample stage; measured with the no-load diagnostic build, a generic issue loop
        // (64-bit descriptors rebuilt per MMA, runtime block loop) took ~2000 cycles per stage - more than the MMAs themselves.
        // Here the low descriptor words are computed once per stage and stepped with an add (K step of 16 samples = 2 KiB = 128
        // in descriptor units), and the block loop is resolved per item into at most two MMAs of fixed N.
        if (lane == 0) {
            uint32_t sc = 0, n_item = 0;
            const uint32_t idesc_bias = umma_idesc(128, 16, kF16, kF16, 1, 0);
            const uint32_t ones_lo = desc_lo(smem + kWOffOnes, 16);
            for (int it = item0; it < a.n_items && !sh.abort; it += item_stride) {
                const WItem w = a.items[it * CL + rank];
                if (w.tile1 <= w.tile0) continue;
                bool ok = wwait(&sh.acc_empty, (n_item & 1) ^ 1, &sh.abort, a.status, 710);
                ++n_item;
                const int n1 = w.n_x < 4 ? w.n_x : 4, n2 = w.n_x - n1;              // X blocks of the first / second MMA
                const uint32_t idesc1 = umma_idesc(128, (uint32_t)(n1 > 0 ? n1 : 1) * 64, kF16, kF16, 1, 1);
                const uint32_t idesc2 = umma_idesc(128, (uint32_t)(n2 > 0 ? n2 : 1) * 64, kF16, kF16, 1, 1);
                uint32_t first = 0;                                                   // accumulate flag of the item's first K step
                for (int tile = w.tile0; tile < w.tile1 && ok; ++tile) {
                    for (int half = 0; half < 2; ++half, ++sc) {
                        const uint32_t stage = sc % kWStages, par = (sc / kWStages) & 1;
                        ok = wwait(&sh.full[stage], par, &sh.abort, a.status, 711);
                        if (!ok) break;
                        tc_fence_after_sync();
#if HN_WEXP != 1                                                    // diagnostic build 1: no MMAs (load path alone)
                        const uint32_t g_addr = smem + stage * kWStageBytes;
                        const uint32_t g_lo = desc_lo(g_addr, kHalfBytes);
                        const uint32_t x_lo = desc_lo(g_addr + 2 * kHalfBytes, kHalfBytes), x2_lo = desc_lo(g_addr + 6 * kHalfBytes, kHalfBytes);
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) {
                            const uint32_t acc = ks == 0 ? first : 1u;
                            if (n1 > 0) umma_f16_lohi(tmem_base, g_lo + ks * 128, x_lo + ks * 128, idesc1, acc);
                            if (n2 > 0) umma_f16_lohi(tmem_base + 256, g_lo + ks * 128, x2_lo + ks * 128, idesc2, acc);
#if HN_WEXP != 4
                            umma_f16_lohi(tmem_base + kBiasCol, g_lo + ks * 128, ones_lo + ks * 2, idesc_bias, acc);
#endif
                        }
#endif
                        first = 1;
                        if (CL == 1) umma_commit(smem_u32(&sh.empty[stage]));
                        else umma_commit_multicast(smem_u32(&sh.empty[stage]), kMask);     // a stage is refilled by all peers: all must release it
                    }
                }
                umma_commit(smem_u32(&sh.acc_full));
            }
        }
    } else {
        // ======================= flush: TMEM -> atomic adds into dW / dbias =======================
        const int row = (warp & 3) * 32 + lane;                    // TMEM lane (a warp may only read its own quarter) = dW row in the chunk
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float inv_scale = 1.0f / __ldg(a.grad_scale);
        uint32_t n_item = 0;
        for (int it = item0; it < a.n_items && !sh.abort; it += item_stride) {
            const WItem w = a.items[it * CL + rank];
            if (w.tile1 <= w.tile0) continue;
            wwait(&sh.acc_full, n_item & 1, &sh.abort, a.status, 720);
            ++n_item;
            tc_fence_after_sync();
            const bool row_ok = row < w.rows;
            for (int k = 0; k < w.n_x; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_base + k * 64 + h * 32, v);
                    tmem_ld_wait();
                    if (row_ok && w.w_idx >= 0 && a.dw[w.w_idx]) {
                        float* dst = a.dw[w.w_idx] + (size_t)(w.row0 + row) * a.ld[w.w_idx] + w.x_col[k] + h * 32;
                        const int nvalid = w.x_valid[k] - h * 32;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nvalid) atomicAdd(dst + i, __uint_as_float(v[i]) * inv_scale);
                    }
                }
            }
            {
                uint32_t v[32];                                   // bias columns (all 16 equal); x32 load stays inside the 512 columns
                tmem_ld32(tmem_base + lane_base + kBiasCol, v);
                tmem_ld_wait();
                if (row_ok && w.bias_off >= 0 && a.dbias)
                    atomicAdd(a.dbias + (size_t)w.b * HN_BIAS_STRIDE + w.bias_off + row, __uint_as_float(v[0]) * inv_scale);
            }
            tc_fence_before_sync();
            mbar_arrive(smem_u32(&sh.acc_empty));
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                              // peers may still multicast into / signal this CTA
    if (warp == 1) tmem_free<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant for the 384 x 384 layers (cta_group::2, M = 256): the pair accumulates the dW rows of chunks 0 and 1 of ONE
// layer over the same samples.  CTA r loads its own dZ chunk and only HALF of the layer-input columns (X blocks 2r, 2r+1 and
// 4+r): the tensor core reads the B operand's halves from both SMs, so 40 KiB instead of 64 KiB land in each SM's shared memory
// per 64-sample stage.  Chunk 2 of those layers has no partner (384 = 256 + 128) and stays on the single-CTA kernel.
// EXPERIMENT, opt-in with HN_WGRAD_PAIRS=1: correct, but slower than the 3-CTA multicast clusters (see hn_mlp_bwd_weights).
constexpr int kPairStages = 5;
constexpr uint32_t kPairStageBytes = 5 * kHalfBytes;          // dZ chunk (2 half-blocks) + 3 X half-blocks = 40 KiB
constexpr uint32_t kPairOffOnes = kPairStages * kPairStageBytes;
constexpr uint32_t kPairSmem = kPairOffOnes + 2048 + 1024;

struct WPairShared {
    uint64_t full[kPairStages], peer_full[kPairStages], empty[kPairStages], acc_full, acc_empty;
    uint32_t tmem_base;
    volatile int abort;
};

__device__ __forceinline__ bool wwait_cluster(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* status, int code) {
    const uint32_t b = smem_u32(bar);
    if (mbar_try_wait_cluster(b, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(b, parity)) {
        if (*abort_flag) return false;
        if (clock64() - t0 > 2000000000ll) { *abort_flag = 1; atomicCAS(status, 0, code); return false; }
    }
    return true;
}

__global__ void __launch_bounds__(kWThreads, 1) mlp_wgrad_pair_kernel(const WArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ WPairShared sh;
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const int item0 = (int)blockIdx.x >> 1, item_stride = (int)gridDim.x >> 1;

    if (tid == 0) {
        for (int i = 0; i < kPairStages; ++i) { mbar_init(smem_u32(&sh.full[i]), 1); mbar_init(smem_u32(&sh.peer_full[i]), 1); mbar_init(smem_u32(&sh.empty[i]), 1); }
        mbar_init(smem_u32(&sh.acc_full), 1);
        mbar_init(smem_u32(&sh.acc_empty), 256);                  // the flush threads of BOTH CTAs release the leader's issuer
        sh.abort = 0;
        mbar_fence_init();
    }
    for (int i = tid; i < 2048 / 4; i += kWThreads)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem + kPairOffOnes + i * 4), "r"(0x3C003C00u) : "memory");
    fence_async_smem();
    if (warp == 1) tmem_alloc_pair<512>(smem_u32(&sh.tmem_base));
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = sh.tmem_base;

    if (warp == 0) {
        // ======================= producer (both CTAs): own dZ chunk + own half of the X columns =======================
        if (lane == 0) {
            uint32_t sc = 0;
            for (int it = item0; it < a.n_items && !sh.abort; it += item_stride) {
                const WItem w = a.items[it];
                const int xb[3] = {w.x_blk[2 * rank], w.x_blk[2 * rank + 1], w.x_blk[4 + rank]};
                for (int tile = w.tile0; tile < w.tile1; ++tile) {
                    for (int half = 0; half < 2; ++half, ++sc) {
                        const uint32_t stage = sc % kPairStages, par = (sc / kPairStages) & 1;
                        if (!wwait_cluster(&sh.empty[stage], par ^ 1, &sh.abort, a.status, 741)) break;
                        const uint32_t fb = smem_u32(&sh.full[stage]);
                        mbar_arrive_expect_tx(fb, kPairStageBytes);
                        const uint32_t dst = smem + stage * kPairStageBytes;
                        for (int k = 0; k < 2; ++k)
                            bulk_g2s(dst + k * kHalfBytes, a.grads + ((size_t)(w.g_blk + 2 * rank + k) * a.n_tiles + tile) * kUnitBytes + half * kHalfBytes, kHalfBytes, fb);
                        for (int k = 0; k < 3; ++k)
                            bulk_g2s(dst + (2 + k) * kHalfBytes, a.act + ((size_t)xb[k] * a.n_tiles + tile) * kUnitBytes + half * kHalfBytes, kHalfBytes, fb);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 1) {
            // ======================= peer: relay "my stage has landed" to the leader =======================
            uint32_t sc = 0;
            for (int it = item0; it < a.n_items && !sh.abort; it += item_stride) {
                const WItem w = a.items[it];
                for (int n = 2 * (w.tile1 - w.tile0); n > 0; --n, ++sc) {
                    const uint32_t stage = sc % kPairStages, par = (sc / kPairStages) & 1;
                    if (!wwait(&sh.full[stage], par, &sh.abort, a.status, 750)) break;
                    mbar_arrive_cluster(smem_u32(&sh.peer_full[stage]), 0);
                }
            }
        } else if (lane == 0) {
            // ======================= leader: MMA issuer =======================
            uint32_t sc = 0, n_item = 0;
            const uint32_t idesc1 = umma_idesc(256, 256, kF16, kF16, 1, 1), idesc2 = umma_idesc(256, 128, kF16, kF16, 1, 1);
            const uint32_t idesc_bias = umma_idesc(256, 16, kF16, kF16, 1, 0);
            for (int it = item0; it < a.n_items && !sh.abort; it += item_stride, ++n_item) {
                const WItem w = a.items[it];
                bool ok = wwait_cluster(&sh.acc_empty, (n_item & 1) ^ 1, &sh.abort, a.status, 742);
                bool first = true;
                for (int tile = w.tile0; tile < w.tile1 && ok; ++tile) {
                    for (int half = 0; half < 2; ++half, ++sc) {
                        const uint32_t stage = sc % kPairStages, par = (sc / kPairStages) & 1;
                        ok = wwait(&sh.full[stage], par, &sh.abort, a.status, 743);
                        if (ok) ok = wwait_cluster(&sh.peer_full[stage], par, &sh.abort, a.status, 744);
                        if (!ok) break;
                        tc_fence_after_sync();
                        const uint32_t g_addr = smem + stage * kPairStageBytes;
                        const uint32_t x_addr = g_addr + 2 * kHalfBytes;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t ad = umma_desc_mnmajor(g_addr, ks, kHalfBytes);
                            const bool acc = !(first && ks == 0);
                            umma_f16_pair(tmem_base, ad, umma_desc_mnmajor(x_addr, ks, kHalfBytes), idesc1, acc);
                            umma_f16_pair(tmem_base + 256, ad, umma_desc_mnmajor(x_addr + 2 * kHalfBytes, ks, kHalfBytes), idesc2, acc);
                            umma_f16_pair(tmem_base + kBiasCol, ad, umma_desc_kmajor(smem + kPairOffOnes, ks), idesc_bias, acc);
                        }
                        first = false;
                        umma_commit_pair(smem_u32(&sh.empty[stage]));          // both CTAs' stages are free once these MMAs retire
                    }
                }
                umma_commit_pair(smem_u32(&sh.acc_full));
            }
        }
    } else {
        // ======================= flush (both CTAs): own 128 rows of the pair's accumulator =======================
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float inv_scale = 1.0f / __ldg(a.grad_scale);
        uint32_t n_item = 0;
        for (int it = item0; it < a.n_items && !sh.abort; it += item_stride, ++n_item) {
            const WItem w = a.items[it];
            wwait_cluster(&sh.acc_full, n_item & 1, &sh.abort, a.status, 745);
            tc_fence_after_sync();
            float* dw = (w.w_idx >= 0) ? a.dw[w.w_idx] : nullptr;
            for (int k = 0; k < 6; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_base + k * 64 + h * 32, v);
                    tmem_ld_wait();
                    if (dw) {
                        float* dst = dw + (size_t)(w.row0 + 128 * rank + row) * a.ld[w.w_idx] + w.x_col[k] + h * 32;
#pragma unroll
                        for (int i = 0; i < 32; ++i) atomicAdd(dst + i, __uint_as_float(v[i]) * inv_scale);
                    }
                }
            }
            {
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_base + kBiasCol, v);
                tmem_ld_wait();
                if (w.bias_off >= 0 && a.dbias)
                    atomicAdd(a.dbias + (size_t)w.b * HN_BIAS_STRIDE + w.bias_off + 128 * rank + row, __uint_as_float(v[0]) * inv_scale);
            }
            tc_fence_before_sync();
            mbar_arrive_cluster(smem_u32(&sh.acc_empty), 0);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_free_pair<512>(tmem_base);
}

static std::mutex g_w_mu;
static bool g_w_ready[64] = {};

// host: enumerate work items.  `cluster` = items for the 3-CTA multicast kernel (three consecutive entries = the three
// 128-channel chunks of one layer over one sample range); `single` = everything else.
// `side` / n_side: single-CTA items for the SMs the resident 3-CTA clusters leave idle (148 - 3 * 45 = 13 on this part): they run on
// a second stream while the cluster kernel runs, the rest of the single-CTA work (`single`) follows on all SMs.
static void build_items(const hn_mlp_bwd_weights_t& a, int n_sm, int n_clusters, int n_pairs, int n_side, std::vector<WItem>& cluster,
                        std::vector<WItem>& single, std::vector<WItem>& pairs, std::vector<WItem>& side) {
    const int tiles_per_item = (int)(((int64_t)a.n_rays * a.n_samples) / HN_TILE);
    bool want_w = false;
    for (int i = 0; i < 12; ++i) want_w = want_w || (a.dw[i] != nullptr);
    struct LayerW { int w_idx; int n_out; int g_blk; int g_dfeat; int bias_off; int x_slot; int n_xblk; bool pe; int x_col0; };
    std::vector<LayerW> layers;
    for (int l = 0; l < 8; ++l)
        layers.push_back({l, HN_HIDDEN, HN_GSLOT_Z0 + 6 * l, 0, l * HN_HIDDEN, l == 0 ? -1 : HN_SLOT_H0 + 6 * (l - 1), l == 0 ? 0 : 6,
                          l == 0 || l == 5, l == 5 ? a.l5_hidden_col : 0});
    layers.push_back({W_R0, HN_HIDDEN, HN_GSLOT_R0, 0, HN_BIAS_OFF_R0, HN_SLOT_H0 + 6 * 7, 6, false, 0});
    layers.push_back({W_R1, HN_RGB1, HN_GSLOT_R1, 0, HN_BIAS_OFF_R1, HN_SLOT_R0, 6, false, 0});
    layers.push_back({W_R2, HN_FEAT, 0, 1, HN_BIAS_OFF_R2, HN_SLOT_X, 3, false, 0});
    layers.push_back({W_DENSITY, 1, HN_GSLOT_DENS, 0, HN_BIAS_OFF_DENSITY, HN_SLOT_H0 + 6 * 7, 6, false, 0});   // density pseudo layer
    auto active = [&](const LayerW& L) { return want_w || a.want_all_bias || L.w_idx == W_L0 || L.w_idx == W_L5 || L.w_idx == W_R1; };
    // CTA pairs: 384 x 384 layers (six hidden X blocks, no PE block); their chunk 2 goes to the single-CTA list
    auto paired = [&](const LayerW& L) { return want_w && n_pairs > 0 && L.n_out == HN_HIDDEN && L.n_xblk == 6 && !L.pe && a.dw[L.w_idx] != nullptr; };
    auto clustered = [&](const LayerW& L) { return !paired(L) && want_w && n_clusters > 0 && L.n_out == HN_HIDDEN && a.dw[L.w_idx] != nullptr; };
    // A unit = one layer (cluster list: its three chunks side by side) or one chunk (single list) of one batch item, over all of
    // the item's tiles, with a cost per tile ~ MMA cycles / operand bytes of a stage.  The units of a list are laid end to end and
    // cut into one equal-cost share per worker (a 3-CTA cluster or a CTA): a worker gets one or two sample ranges, every worker
    // finishes at the same time, and each accumulator is flushed (atomics) only once or twice.  [Equal sample splits per layer
    // left 108 equal items for 45 resident clusters: a makespan of 3 items against a mean of 2.4.]
    // Measured (no-MMA / no-load diagnostic builds, ncu): a stage takes ~2400 cycles almost regardless of how many operand
    // blocks it carries - the kernel is bound by DRAM latency against the bytes three stages keep in flight, not by tensor or
    // byte throughput - so a tile costs about the same whatever the layer width; the block count only adds a small slope.
    static const double slope = [] { const char* e = getenv("HN_WGRAD_SLOPE"); return e ? atof(e) : 40.0; }();   // tuning knob (cycles per operand block)
    auto stage_cost = [](int n_x) { return 1000.0 + slope * n_x; };
    struct Unit { std::vector<WItem> tmpl; int b; double cost; int t0, t1; };      // tiles [t0, t1) of batch item b
    std::vector<Unit> units_c, units_s;
    int layers_pair = 0;
    auto make_item = [&](const LayerW& L, int j, int b) {
        WItem w{};
        w.w_idx = (int16_t)((want_w && a.dw[L.w_idx]) ? L.w_idx : -1);
        w.row0 = (int16_t)(128 * j);
        w.rows = (int16_t)std::min(128, L.n_out - 128 * j);
        w.g_blk = L.g_blk + 2 * j;
        w.g_dfeat = (int16_t)L.g_dfeat;
        w.bias_off = (int16_t)(L.bias_off + 128 * j);
        w.b = b;
        int n = 0;
        if (w.w_idx >= 0) {
            for (int k = 0; k < L.n_xblk; ++k) { w.x_blk[n] = L.x_slot + k; w.x_col[n] = (int16_t)(L.x_col0 + 64 * k); w.x_valid[n] = 64; ++n; }
            if (L.pe) { w.x_blk[n] = HN_SLOT_PE; w.x_col[n] = 0; w.x_valid[n] = HN_PE; ++n; }
        }
        w.n_x = (int16_t)n;
        return w;
    };
    for (const LayerW& L : layers) {
        if (!active(L)) continue;
        const bool cl = clustered(L), pr = paired(L);
        if (pr) ++layers_pair;
        for (int b = 0; b < a.B; ++b) {
            if (cl) {
                Unit u; u.b = b; u.t0 = 0; u.t1 = tiles_per_item;
                for (int j = 0; j < 3; ++j) u.tmpl.push_back(make_item(L, j, b));
                u.cost = stage_cost(u.tmpl[0].n_x);
                units_c.push_back(u);
            } else {
                for (int j = pr ? 2 : 0; j * 128 < L.n_out; ++j) {
                    Unit u; u.b = b; u.t0 = 0; u.t1 = tiles_per_item;
                    u.tmpl.push_back(make_item(L, j, b));
                    u.cost = stage_cost(u.tmpl[0].n_x);
                    units_s.push_back(u);
                }
            }
        }
    }
    auto partition = [&](const std::vector<Unit>& units, int workers, int group, std::vector<WItem>& out) {
        if (units.empty() || workers <= 0) return;
        double total = 0;
        for (const Unit& u : units) total += u.cost * (u.t1 - u.t0);
        const double quota = total / workers;
        std::vector<std::vector<WItem>> lists(workers);               // `group` consecutive entries per piece
        int wk = 0;
        double need = quota;
        for (const Unit& u : units) {
            int t = u.t0;
            while (t < u.t1) {
                int take = (int)(need / u.cost + 0.5);
                take = take < 1 ? 1 : take;
                if (take > u.t1 - t || wk == workers - 1) take = u.t1 - t;
                if (u.t1 - t - take > 0 && (u.t1 - t - take) * u.cost < 0.03 * quota) take = u.t1 - t;   // no crumbs
                for (int r = 0; r < group; ++r) {
                    WItem w = u.tmpl[r];
                    w.tile0 = u.b * tiles_per_item + t;
                    w.tile1 = w.tile0 + take;
                    lists[wk].push_back(w);
                }
                need -= take * u.cost;
                t += take;
                if (need < 0.03 * quota && wk < workers - 1) { ++wk; need += quota; }
            }
        }
        size_t rounds = 0;
        for (const auto& l : lists) rounds = std::max(rounds, l.size() / group);
        if (getenv("HN_WGRAD_DEBUG")) {
            for (int w = 0; w < workers; ++w) {
                double c = 0;
                fprintf(stderr, "worker %d:", w);
                for (size_t i = 0; i < lists[w].size(); i += group) {
                    const WItem& it = lists[w][i];
                    c += stage_cost(it.n_x) * (it.tile1 - it.tile0);
                    fprintf(stderr, " [w%d b%d g%d nx%d tiles %d-%d]", it.w_idx, it.b, it.g_blk, it.n_x, it.tile0, it.tile1);
                }
                fprintf(stderr, "  cost %.0f (quota %.0f)\n", c, quota);
            }
        }
        // round-robin order: the j-th piece of worker w sits at index j * workers + w (padding entries have no tiles)
        for (size_t j = 0; j < rounds; ++j)
            for (int w = 0; w < workers; ++w)
                for (int r = 0; r < group; ++r)
                    out.push_back(j * group + r < lists[w].size() ? lists[w][j * group + r] : WItem{});
    };
    partition(units_c, n_clusters, 3, cluster);
    if (n_side > 0 && !units_c.empty() && !units_s.empty()) {
        // share of the single-CTA work that n_side SMs finish in about the time the cluster kernel takes (cost ~ stages)
        double cost_c = 0, cost_s = 0;
        for (const Unit& u : units_c) cost_c += u.cost * 3 * (u.t1 - u.t0);
        for (const Unit& u : units_s) cost_s += u.cost * (u.t1 - u.t0);
        static const double tune = [] { const char* e = getenv("HN_WGRAD_SIDE"); return e ? atof(e) : 0.85; }();
        double f = tune * (cost_c / (3.0 * n_clusters)) * n_side / cost_s;          // (time of the cluster phase) x n_side / (single work)
        f = f > 0.9 ? 0.9 : f;
        std::vector<Unit> first, rest;
        for (const Unit& u : units_s) {
            const int cut = u.t0 + (int)((u.t1 - u.t0) * f);
            if (cut > u.t0) { Unit p = u; p.t1 = cut; first.push_back(p); }
            if (cut < u.t1) { Unit p = u; p.t0 = cut; rest.push_back(p); }
        }
        partition(first, n_side, 1, side);
        partition(rest, n_sm, 1, single);
    } else {
        partition(units_s, n_sm, 1, single);
    }
    // opt-in CTA-pair items (uniform sample splits)
    if (layers_pair > 0) {
        int sp = (2 * n_pairs + layers_pair * a.B - 1) / (layers_pair * a.B);
        sp = sp < 1 ? 1 : (sp > tiles_per_item ? tiles_per_item : sp);
        for (const LayerW& L : layers) {
            if (!active(L) || !paired(L)) continue;
            for (int b = 0; b < a.B; ++b)
                for (int q = 0; q < sp; ++q) {
                    WItem w = make_item(L, 0, b);
                    w.tile0 = b * tiles_per_item + (int)((int64_t)tiles_per_item * q / sp);
                    w.tile1 = b * tiles_per_item + (int)((int64_t)tiles_per_item * (q + 1) / sp);
                    if (w.tile1 > w.tile0) pairs.push_back(w);
                }
        }
    }
}

static int launch_wgrad(const WArgs& k, int cl, int grid, cudaStream_t st) {
    if (cl == 1) {
        mlp_wgrad_kernel<1><<<grid, kWThreads, kWgradSmem, st>>>(k);
        return check_launch("hn_mlp_bwd_weights");
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kWThreads); cfg.dynamicSmemBytes = kWgradSmem; cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 3; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_wgrad_kernel<3>, k);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    return check_launch("hn_mlp_bwd_weights (3-CTA clusters)");
}

static int g_w_clusters[64] = {};
static int g_w_pairs[64] = {};
static cudaStream_t g_w_side[64] = {};
static cudaEvent_t g_w_fork[64] = {}, g_w_join[64] = {};

}  // namespace hn

extern "C" int hn_mlp_bwd_weights(const hn_mlp_bwd_weights_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->act || !a->grads || !a->dfeat_image || !a->grad_scale || !a->status)
        return set_error(HN_E_BADARG, "hn_mlp_bwd_weights: null pointer");
    if (int rc = check_geometry(a->B, a->n_rays, a->n_samples, "hn_mlp_bwd_weights")) return rc;
    if (!a->items_workspace || a->items_workspace_bytes < hn_wgrad_workspace_bytes(a->B))
        return set_error(HN_E_BADARG, "hn_mlp_bwd_weights: items workspace missing or too small (hn_wgrad_workspace_bytes)");
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_w_mu);
        if (dev < 64 && !g_w_ready[dev]) {
            cudaError_t e = cudaFuncSetAttribute(mlp_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem);
            if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
            // how many 3-CTA clusters can be resident at once (GPC sizes strand a few SMs)
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(3 * 49); cfg.blockDim = dim3(kWThreads); cfg.dynamicSmemBytes = kWgradSmem;
            cudaLaunchAttribute attr{};
            attr.id = cudaLaunchAttributeClusterDimension;
            attr.val.clusterDim.x = 3; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
            cfg.attrs = &attr; cfg.numAttrs = 1;
            int n = 0;
            const char* env = getenv("HN_WGRAD_CLUSTERS");
            if (env && atoi(env) == 0) n = 0;
            else if (cudaOccupancyMaxActiveClusters(&n, mlp_wgrad_kernel<3>, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
            g_w_clusters[dev] = n;
            // resident CTA pairs of the pair kernel.  Opt-in (HN_WGRAD_PAIRS=1): validated by the same parity tests, but measured
            // SLOWER on B200 (2.90 vs 2.21 ms per Reso64 batch-2 pass) - what bounds this kernel is L2 -> SM read traffic, which the
            // 3-CTA multicast clusters already cut to the compulsory 32 KiB per chunk and stage; a pair reads 40 KiB and strands chunk 2.
            int np = 0;
            const char* envp = getenv("HN_WGRAD_PAIRS");
            if (envp && atoi(envp) != 0) {
                e = cudaFuncSetAttribute(mlp_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem);
                if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
                cfg.gridDim = dim3(2 * 74); cfg.dynamicSmemBytes = kPairSmem;
                attr.val.clusterDim.x = 2;
                if (cudaOccupancyMaxActiveClusters(&np, mlp_wgrad_pair_kernel, &cfg) != cudaSuccess) { np = 0; cudaGetLastError(); }
            }
            g_w_pairs[dev] = np;
            // a second stream (+ two events) per device: the single-CTA items that fill the SMs the clusters leave idle run on it,
            // forked from and joined back into the caller's stream (HN_WGRAD_SIDE=0 disables it)
            const char* envs = getenv("HN_WGRAD_SIDE");
            if (!(envs && atof(envs) == 0.0)) {
                if (cudaStreamCreateWithFlags(&g_w_side[dev], cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&g_w_fork[dev], cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&g_w_join[dev], cudaEventDisableTiming) != cudaSuccess) { g_w_side[dev] = nullptr; cudaGetLastError(); }
            }
            g_w_ready[dev] = true;
        }
    }
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int n_clusters = dev < 64 ? g_w_clusters[dev] : 0;
    const int n_pairs = dev < 64 ? g_w_pairs[dev] : 0;
    std::vector<WItem> cluster, single, pairs, side;
    cudaStream_t side_stream = dev < 64 ? g_w_side[dev] : nullptr;
    const int n_idle = n_sm - 3 * n_clusters;
    const int n_side = (side_stream && n_pairs == 0 && n_clusters > 0 && n_idle >= 4) ? n_idle : 0;
    build_items(*a, n_sm, n_clusters, n_pairs, n_side, cluster, single, pairs, side);
    if (cluster.empty() && single.empty() && pairs.empty() && side.empty()) return HN_OK;
    if ((cluster.size() + single.size() + pairs.size() + side.size()) * sizeof(WItem) > a->items_workspace_bytes)
        return set_error(HN_E_BADARG, "hn_mlp_bwd_weights: items workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<WItem> all(cluster);
    all.insert(all.end(), single.begin(), single.end());
    all.insert(all.end(), pairs.begin(), pairs.end());
    all.insert(all.end(), side.begin(), side.end());
    // the item table is tiny (<100 KiB); pageable -> device copy is stream-ordered and returns after staging
    cudaError_t e = cudaMemcpyAsync(a->items_workspace, all.data(), all.size() * sizeof(WItem), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    WArgs k{};
    k.act = (const uint8_t*)a->act; k.grads = (const uint8_t*)a->grads; k.dfeat_image = (const uint8_t*)a->dfeat_image;
    k.grad_scale = a->grad_scale;
    for (int i = 0; i < 12; ++i) { k.dw[i] = a->dw[i]; k.ld[i] = a->ld[i]; }
    k.dbias = a->dbias;
    k.n_tiles = (int)(total_samples(a->B, a->n_rays, a->n_samples) / HN_TILE);
    k.status = a->status;
    if (!pairs.empty()) {
        k.items = (const WItem*)a->items_workspace + cluster.size() + single.size(); k.n_items = (int)pairs.size();
        const int np = k.n_items < n_pairs ? k.n_items : n_pairs;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * np); cfg.blockDim = dim3(kWThreads); cfg.dynamicSmemBytes = kPairSmem; cfg.stream = st;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        cudaError_t pe = cudaLaunchKernelEx(&cfg, mlp_wgrad_pair_kernel, k);
        if (pe != cudaSuccess) return set_error((int)pe, cudaGetErrorString(pe));
        if (int rc = check_launch("hn_mlp_bwd_weights (CTA pairs)")) return rc;
    }
    const bool forked = !side.empty() && !cluster.empty();
    // the side stream and its two events are per device, shared by all callers: the fork / launch / join sequence is serialised
    std::unique_lock<std::mutex> side_lock(g_w_mu, std::defer_lock);
    if (forked) side_lock.lock();
    if (forked && cudaEventRecord(g_w_fork[dev], st) != cudaSuccess) return set_error(HN_E_PROTOCOL, "hn_mlp_bwd_weights: event record failed");
    if (!cluster.empty()) {
        k.items = (const WItem*)a->items_workspace; k.n_items = (int)cluster.size() / 3;
        const int nc = n_clusters;                                  // the balanced schedule has one column per resident cluster
        if (int rc = launch_wgrad(k, 3, 3 * nc, st)) return rc;
    }
    if (forked) {
        // the idle SMs' share, concurrently with the clusters (launched after them: it can only take what they leave free)
        cudaStreamWaitEvent(side_stream, g_w_fork[dev], 0);
        WArgs ks = k;
        ks.items = (const WItem*)a->items_workspace + cluster.size() + single.size() + pairs.size(); ks.n_items = (int)side.size();
        if (int rc = launch_wgrad(ks, 1, n_side, side_stream)) return rc;
        cudaEventRecord(g_w_join[dev], side_stream);
        cudaStreamWaitEvent(st, g_w_join[dev], 0);
    }
    if (!single.empty()) {
        k.items = (const WItem*)a->items_workspace + cluster.size(); k.n_items = (int)single.size();
        const int grid = n_sm;
        if (int rc = launch_wgrad(k, 1, grid, st)) return rc;
    }
    return HN_OK;
}

extern "C" size_t hn_wgrad_workspace_bytes(int B) {
    // upper bound: 35 (layer, chunk) pairs x B x splits, splits chosen so that items <= 2*SMs + pairs*B
    // upper bound: per list (#workers + #units) pieces, padded to whole rounds: clusters 3 * 3 * (49 + 9 B), single 3 * (160 + 9 B), pairs
    return (size_t)(200 * (size_t)(B > 0 ? B : 1) + 2000) * sizeof(hn::WItem);
}
