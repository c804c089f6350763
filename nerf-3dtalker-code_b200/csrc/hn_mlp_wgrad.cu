// hn_mlp_wgrad.cu — weight gradients of fg_CD_predictor: dW_l = sum over samples of dZ_l^T x X_l, where dZ_l are
// the (loss-scaled, fp16) pre-activation gradients saved by hn_mlp_bwd_data and X_l the layer inputs saved by
// hn_mlp_fwd — both already stored as tensor-core operand images, consumed here as MN-major operands
// (contraction over the 128 sample rows of a block), so nothing is transposed or re-laid-out.
//
// Work item = (layer, 128-channel chunk of dZ, batch item, sample split).  A CTA accumulates the item's
// [128 x K_in] slice of dW in TMEM over all its tiles and flushes it once with atomic adds.  While the tensor
// core works through a stage, the four otherwise idle flush warps read the same shared-memory stage on the
// CUDA cores: the column sums of dZ (= the bias gradient, per item, from which the host derives the
// latent-code and folded weight-column gradients) and, in RGB_layer_0's items, the density head's weight
// gradient sum_m dsigma_m * h7_m (h7 is RGB_layer_0's layer input, already in shared memory) - no ones-operand
// MMAs, no separate pass over h7.  [Round 1 spent 17 % of the MMA time on N = 16 bias products and re-read
// h7 (6 blocks per tile) for a one-row "density pseudo layer"; that layer survives only as the fallback when
// RGB_layer_0's own weight gradient is not requested.]
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>
#include "hn_api.h"
#include "hn_mlp_sched.h"
#include "hn_tc.cuh"

#ifndef HN_WEXP
#define HN_WEXP 0
#endif

namespace hn {

// A ring stage holds kPieceRows sample rows of every operand block of the item: 64 rows x 3 stages by default.  32-row stages (6 of
// them, HN_WPIECES=4) keep five sixths instead of two thirds of the ring in flight, but were measured much SLOWER on B200 (3.0 vs
// 1.79 ms): a stage costs ~1500 cycles of producer / barrier / commit latency almost whatever it carries, so halving it doubles that.
#ifndef HN_WPIECES
#define HN_WPIECES 2
#endif
constexpr int kWPieces = HN_WPIECES;                         // stages per 128-sample tile: 2 or 4
constexpr int kPieceRows = 128 / kWPieces;
constexpr int kWKSteps = kPieceRows / 16;                    // MMA K steps (16 samples) per stage
constexpr int kWStages = 3 * kWPieces / 2;
constexpr int kHalfBytes = kUnitBytes / kWPieces;            // bytes of one operand block's share of a stage (8 KiB or 4 KiB)
constexpr int kWMaxX = 7;
constexpr uint32_t kWStageBytes = (2 + kWMaxX) * kHalfBytes;  // 72 KiB / 36 KiB
// Where the bias gradient (column sums of dZ) comes from: 0 = the CUDA-core readers, 1 = a 16-column "ones" product of the tensor core
// (round 1's way: one more read of the A operand per K step, but no LSU traffic on the operand stage), 2 = readers with half-precision
// partial sums over the 8 rows a thread owns.  Measured on B200 (Reso64 batch 2): see DESIGN.md.
#ifndef HN_WBIAS
#define HN_WBIAS 1
#endif
constexpr uint32_t kWOffOnes = kWStages * kWStageBytes;
constexpr uint32_t kWgradSmem = kWOffOnes + 2048 + 1024;
constexpr uint32_t kBiasCol = 448;
constexpr int kWThreads = 192;                              // warp 0 producer, warp 1 MMA (+TMEM alloc), warps 2..5 readers + flush
constexpr int kWReaders = 128;
#ifndef HN_WPREFETCH
#define HN_WPREFETCH 0
#endif
constexpr int kWPrefetch = HN_WPREFETCH;                    // L2 prefetch distance in 64-sample stages.  0 = off: measured SLOWER with it (2.24 ms at 4 stages, 2.39 at 8,
                                                            // 2.63 at 32, against 2.11 without) - the memory system is already saturated
constexpr int kWProd = kWPieces == 2 ? 5 : 9;               // producer lanes (see the producer role): one per copy of a stage at 32-row stages

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
    int16_t dens;             // 0: none; 1: density-head weight gradient over X blocks dens_x0, dens_x0 + 1; 2: ... and its bias gradient
    int16_t dens_x0;
    int32_t dens_blk;         // block of the density head's gradient column in `grads`
    int32_t slot;             // deterministic mode: index of this item's private partial-sum slice
    int32_t dens_row1;        // 1 + accumulator row that holds the density head's weight / bias gradient, 0 = none (see build_items)
};

struct WShared {
    uint64_t full[kWStages], empty[kWStages], acc_full, acc_empty;
    float dsr[kWStages][kPieceRows];  // density head: dL/d(pre-ReLU density) of the stage's samples (loss-scaled)
    int reader_quit;          // set by a reader whose wait failed; read by all readers behind a barrier (uniform exit: no hung bar.sync)
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
    float* dwf;               // fused RGB_layer_1 x RGB_layer_0 gradient [192, 384] (r0_fused), replaces dw[W_R1] as the destination
    float* dbias;
    const WItem* items; int n_items; int n_tiles;
    int* status;
    float* partials;          // deterministic mode: [slots][kDetSlotFloats] private partial sums (no atomics), else NULL
};

// deterministic mode: every work item flushes its accumulator (128 rows x up to 448 columns), its bias column sums (two reader
// warps per 64-column block write separately) and, for the density fold, its two partial rows into a private slice; a second
// kernel adds the slices of every destination in item order.  Same arithmetic, fixed summation order: run-to-run bit-identical.
constexpr int kDetAccFloats = 128 * 448;
constexpr int kDetSlotFloats = kDetAccFloats + 128 /*bias from the MMA path*/ + 2 * 128 /*bias, reader halves*/ + 2 * 128 /*density, reader halves*/ + 2 /*density bias*/ + 6;

// `count` arrivals at once on a barrier of this CTA
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// arrival on the same barrier in CTA `rank` of the cluster, default (CTA-scope release) semantics as in the multicast pipelines of
// the vendor libraries: the readers' shared-memory loads have returned before the named barrier that precedes this call
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t bar, uint32_t rank) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(bar, rank)) : "memory");
}

__device__ __forceinline__ void reader_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWReaders) : "memory"); }

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
        // a stage is released by the MMAs of every CTA of the cluster (their commits are multicast) and by every CTA's readers
        for (int i = 0; i < kWStages; ++i) { mbar_init(smem_u32(&sh.full[i]), 1); mbar_init(smem_u32(&sh.empty[i]), 2 * CL); }
        mbar_init(smem_u32(&sh.acc_full), 1);
        mbar_init(smem_u32(&sh.acc_empty), 128);
        sh.abort = 0;
        sh.reader_quit = 0;
        mbar_fence_init();
    }
#if HN_WBIAS == 1
    // "ones" operand: 16 rows x 64 samples of 1.0h (K-major image; every element equal, so swizzling is moot)
    for (int i = tid; i < 2048 / 4; i += kWThreads)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem + kWOffOnes + i * 4), "r"(0x3C003C00u) : "memory");
    fence_async_smem();
#endif
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
                    for (int half = 0; half < kWPieces; ++half, ++sc) {
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
                        const int ahead = kWPieces * (tile - w.tile0) + half + kWPrefetch;
                        const int pf_tile = w.tile0 + ahead / kWPieces, pf_half = ahead % kWPieces;
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
#if HN_WBIAS == 1
            const uint32_t idesc_bias = umma_idesc(128, 16, kF16, kF16, 1, 0);
            const uint32_t ones_lo = desc_lo(smem + kWOffOnes, 16);
#endif
            for (int it = item0; it < a.n_items && !sh.abort; it += item_stride) {
                const WItem w = a.items[it * CL + rank];
                if (w.tile1 <= w.tile0) continue;
                bool ok = wwait(&sh.acc_empty, (n_item & 1) ^ 1, &sh.abort, a.status, 710);
                ++n_item;
                const bool reading = w.dens || (HN_WBIAS != 1 && w.bias_off >= 0 && a.dbias != nullptr);   // (same predicate in the reader role)
                const int n1 = w.n_x < 4 ? w.n_x : 4, n2 = w.n_x - n1;              // X blocks of the first / second MMA
                const uint32_t idesc1 = umma_idesc(128, (uint32_t)(n1 > 0 ? n1 : 1) * 64, kF16, kF16, 1, 1);
                const uint32_t idesc2 = umma_idesc(128, (uint32_t)(n2 > 0 ? n2 : 1) * 64, kF16, kF16, 1, 1);
                uint32_t first = 0;                                                   // accumulate flag of the item's first K step
                for (int tile = w.tile0; tile < w.tile1 && ok; ++tile) {
                    for (int half = 0; half < kWPieces; ++half, ++sc) {
                        const uint32_t stage = sc % kWStages, par = (sc / kWStages) & 1;
                        ok = wwait(&sh.full[stage], par, &sh.abort, a.status, 711);
                        if (!ok) break;
                        tc_fence_after_sync();
#if HN_WEXP != 1                                                    // diagnostic build 1: no MMAs (load path alone)
                        const uint32_t g_addr = smem + stage * kWStageBytes;
                        const uint32_t g_lo = desc_lo(g_addr, kHalfBytes);
                        const uint32_t x_lo = desc_lo(g_addr + 2 * kHalfBytes, kHalfBytes), x2_lo = desc_lo(g_addr + 6 * kHalfBytes, kHalfBytes);
#pragma unroll
                        for (uint32_t ks = 0; ks < (uint32_t)kWKSteps; ++ks) {
                            const uint32_t acc = ks == 0 ? first : 1u;
                            if (n1 > 0) umma_f16_lohi(tmem_base, g_lo + ks * 128, x_lo + ks * 128, idesc1, acc);
                            if (n2 > 0) umma_f16_lohi(tmem_base + 256, g_lo + ks * 128, x2_lo + ks * 128, idesc2, acc);
#if HN_WBIAS == 1
                            if (w.bias_off >= 0) umma_f16_lohi(tmem_base + kBiasCol, g_lo + ks * 128, ones_lo + ks * 2, idesc_bias, acc);
#endif
                        }
#endif
                        first = 1;
                        if (CL == 1) umma_commit(smem_u32(&sh.empty[stage]));
                        else umma_commit_multicast(smem_u32(&sh.empty[stage]), kMask);     // a stage is refilled by all peers: all must release it
                        if (!reading) mbar_arrive_n(smem_u32(&sh.empty[stage]), (uint32_t)CL);   // no reader touches this item's stages: their share of the release
                    }
                }
                umma_commit(smem_u32(&sh.acc_full));
            }
        }
    } else {
        // ======================= readers (CUDA-core side sums over the staged operands), then flush =======================
        // 128 threads.  Thread -> (hb: which 64-column block of the 128-channel chunk, c8: which 16-byte chunk of the 128-byte
        // row, sg: which 8 of the stage's 64 sample rows); a quarter warp covers the eight chunks of one row, so the 128-bit
        // shared-memory loads are conflict-free in the swizzled image.
        const int rt = tid - 64;                                   // 0..127
        const int rw = rt >> 5;
        const int hb = rw & 1, c8 = lane & 7, sg = (lane >> 3) + 4 * (rw >> 1);
        const int row = (warp & 3) * 32 + lane;                    // TMEM lane (a warp may only read its own quarter) = dW row in the chunk
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const float inv_scale = 1.0f / __ldg(a.grad_scale);
        uint32_t n_item = 0, sc = 0;
        bool quit = false;
        for (int it = item0; it < a.n_items && !quit; it += item_stride) {       // (no per-thread abort test here: the exit must be uniform)
            const WItem w = a.items[it * CL + rank];
            if (w.tile1 <= w.tile0) continue;
            float bsum[8], dsum[8], dsr_acc = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { bsum[j] = 0.f; dsum[j] = 0.f; }
            const bool want_bias = (w.bias_off >= 0 && a.dbias != nullptr);
            const bool read_bias = want_bias && HN_WBIAS != 1;
            const bool reading = w.dens || read_bias;
            if (!reading) sc += (uint32_t)kWPieces * (uint32_t)(w.tile1 - w.tile0);    // the MMA thread releases these stages on the readers' behalf
            for (int tile = w.tile0; tile < w.tile1 && !quit && reading; ++tile) {
                for (int half = 0; half < kWPieces; ++half, ++sc) {
                    const uint32_t stage = sc % kWStages, par = (sc / kWStages) & 1;
                    if (w.dens && rt < kPieceRows) {
                        // dL/d(pre-ReLU density) of sample rt of this stage: column 0 of the density gradient block (fetched while
                        // the stage's bulk copies are still in flight)
                        const int r = half * kPieceRows + rt;
                        const __half hv = *reinterpret_cast<const __half*>(a.grads + ((size_t)w.dens_blk * a.n_tiles + tile) * kUnitBytes + image_offset((uint32_t)r, 0u));
                        const float v = __half2float(hv);
                        sh.dsr[stage][rt] = v;
                        dsr_acc += v;
                    }
                    // ONE thread polls the stage's barrier (with a pause: every mbarrier operation of the CTA goes through one unit, and
                    // the MMA issuer's own waits must not queue behind 128 pollers - a try_wait per reader thread and stage was measured:
                    // 4.2 ms instead of 1.85); the others join at the named barrier, which orders their reads behind its observation
                    if (rt == 0) {
                        const uint32_t fb = smem_u32(&sh.full[stage]);
                        bool got = mbar_try_wait(fb, par);
                        const long long t0 = clock64();
                        while (!got) {
                            __nanosleep(32);
                            got = mbar_try_wait(fb, par);
                            if (!got && (sh.abort || clock64() - t0 > 2000000000ll)) { sh.abort = 1; atomicCAS(a.status, 0, 730); break; }
                        }
                        if (!got) sh.reader_quit = 1;
                    }
                    reader_sync();                                 // the stage's dsr values and the quit flag are visible to all readers
                    quit = sh.reader_quit != 0;                    // (written only before this barrier: every reader sees the same value)
#if HN_WEXP != 8                                                   // diagnostic build 8: no reader arithmetic
                    constexpr int RPT = kPieceRows / 8;            // sample rows per reader thread and stage
                    const uint32_t st_base = smem + stage * kWStageBytes;
                    auto row_off = [&](int i) -> uint32_t {        // byte offset of this thread's 16-byte chunk of its i-th row (swizzled image)
                        const uint32_t r = (uint32_t)(sg * RPT + i);
                        return (r >> 3) * 1024u + (r & 7u) * 128u + (((uint32_t)c8 ^ (r & 7u)) << 4);
                    };
                    if (read_bias && !quit) {
                        const uint32_t gb = st_base + (uint32_t)hb * kHalfBytes;
#if HN_WBIAS == 2
                        __half2 h0 = __float2half2_rn(0.f), h1 = h0, h2 = h0, h3 = h0;
#pragma unroll
                        for (int i = 0; i < RPT; ++i) {
                            const uint4 q = ld_shared_v4(gb + row_off(i));
                            h0 = __hadd2(h0, *reinterpret_cast<const __half2*>(&q.x)); h1 = __hadd2(h1, *reinterpret_cast<const __half2*>(&q.y));
                            h2 = __hadd2(h2, *reinterpret_cast<const __half2*>(&q.z)); h3 = __hadd2(h3, *reinterpret_cast<const __half2*>(&q.w));
                        }
                        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
                        bsum[0] += f0.x; bsum[1] += f0.y; bsum[2] += f1.x; bsum[3] += f1.y; bsum[4] += f2.x; bsum[5] += f2.y; bsum[6] += f3.x; bsum[7] += f3.y;
#else
#pragma unroll
                        for (int i = 0; i < RPT; ++i) {
                            const uint4 q = ld_shared_v4(gb + row_off(i));
                            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&q.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
                            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&q.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&q.w));
                            bsum[0] += f0.x; bsum[1] += f0.y; bsum[2] += f1.x; bsum[3] += f1.y; bsum[4] += f2.x; bsum[5] += f2.y; bsum[6] += f3.x; bsum[7] += f3.y;
                        }
#endif
                    }
                    if (w.dens && !quit) {
                        const uint32_t xb = st_base + (uint32_t)(2 + w.dens_x0 + hb) * kHalfBytes;
#pragma unroll
                        for (int i = 0; i < RPT; ++i) {
                            const float ds = sh.dsr[stage][sg * RPT + i];
                            const uint4 q = ld_shared_v4(xb + row_off(i));
                            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&q.x)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
                            const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&q.z)), f3 = __half22float2(*reinterpret_cast<const __half2*>(&q.w));
                            dsum[0] = fmaf(ds, f0.x, dsum[0]); dsum[1] = fmaf(ds, f0.y, dsum[1]); dsum[2] = fmaf(ds, f1.x, dsum[2]); dsum[3] = fmaf(ds, f1.y, dsum[3]);
                            dsum[4] = fmaf(ds, f2.x, dsum[4]); dsum[5] = fmaf(ds, f2.y, dsum[5]); dsum[6] = fmaf(ds, f3.x, dsum[6]); dsum[7] = fmaf(ds, f3.y, dsum[7]);
                        }
                    }
#endif
                    // every reader is done with the stage: one thread releases it in every CTA of the cluster (peers multicast
                    // layer-input blocks into this CTA's stage, so their producers must see this CTA's readers too)
                    reader_sync();
                    if (quit) break;
                    // Release.  Only the density items read blocks that PEER CTAs multicast into this stage (layer inputs); every
                    // other item reads its own dZ blocks only, so its CL reader arrivals can all be local - a remote release-arrive
                    // costs a cluster-scope fence (three of them from one thread per stage took the kernel from 1.85 to 4.1 ms).
                    if (CL == 1 || !w.dens) {
                        if (rt == 0) mbar_arrive_n(smem_u32(&sh.empty[stage]), (uint32_t)CL);
                    } else if (rt < CL) {
                        mbar_arrive_remote_relaxed(smem_u32(&sh.empty[stage]), (uint32_t)rt);     // one thread per peer, in parallel
                    }
                }
            }
            // ---- side sums: fold the four sample groups of a warp, then one atomic per column and warp
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                bsum[j] += __shfl_xor_sync(0xffffffffu, bsum[j], 8); bsum[j] += __shfl_xor_sync(0xffffffffu, bsum[j], 16);
                dsum[j] += __shfl_xor_sync(0xffffffffu, dsum[j], 8); dsum[j] += __shfl_xor_sync(0xffffffffu, dsum[j], 16);
            }
            float* const slot = a.partials ? a.partials + (size_t)w.slot * kDetSlotFloats : nullptr;
            if (slot && lane < 8) {
                const int ch = hb * 64 + c8 * 8, sgh = rw >> 1;     // the two sample-group halves of a block write separate rows
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (read_bias) slot[kDetAccFloats + 128 + sgh * 128 + ch + j] = bsum[j];
                    if (w.dens) slot[kDetAccFloats + 384 + sgh * 128 + ch + j] = dsum[j];
                }
            } else if (lane < 8) {
                const int ch = hb * 64 + c8 * 8;                    // first of this lane's eight channels / layer-input columns
                if (read_bias) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (ch + j < w.rows) atomicAdd(a.dbias + (size_t)w.b * HN_BIAS_STRIDE + w.bias_off + ch + j, bsum[j] * inv_scale);
                }
                if (w.dens && a.dw[W_DENSITY]) {
                    float* dst = a.dw[W_DENSITY] + w.x_col[w.dens_x0 + hb] + c8 * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) atomicAdd(dst + j, dsum[j] * inv_scale);
                }
            }
            if (w.dens == 2 && rt < 64 && a.dbias) {                 // (threads kPieceRows..63 hold zeros)
#pragma unroll
                for (int sft = 16; sft >= 1; sft >>= 1) dsr_acc += __shfl_xor_sync(0xffffffffu, dsr_acc, sft);
                if (lane == 0) {
                    if (slot) slot[kDetAccFloats + 640 + rw] = dsr_acc;
                    else atomicAdd(a.dbias + (size_t)w.b * HN_BIAS_STRIDE + HN_BIAS_OFF_DENSITY, dsr_acc * inv_scale);
                }
            }
            if (quit) break;
            // ---- flush: TMEM -> atomic adds into dW
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
                    if (slot) {                                    // private slice, plain 128-bit stores
                        float4* dst = reinterpret_cast<float4*>(slot + (size_t)row * 448 + k * 64 + h * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
                    } else if (w.dens_row1 && row == w.dens_row1 - 1) {
                        float* dst = a.dw[W_DENSITY] + w.x_col[k] + h * 32;     // the free density row (build_items): dW_density = sum dsigma x h7
#pragma unroll
                        for (int i = 0; i < 32; ++i) atomicAdd(dst + i, __uint_as_float(v[i]) * inv_scale);
                    } else if (row_ok && w.w_idx >= 0 && a.dw[w.w_idx]) {
                        const bool to_f = (w.w_idx == W_R1 && a.dwf != nullptr);
                        float* dst = (to_f ? a.dwf : a.dw[w.w_idx]) + (size_t)(w.row0 + row) * (to_f ? HN_HIDDEN : a.ld[w.w_idx]) + w.x_col[k] + h * 32;
                        const int nvalid = w.x_valid[k] - h * 32;
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i < nvalid) atomicAdd(dst + i, __uint_as_float(v[i]) * inv_scale);
                    }
                }
            }
#if HN_WBIAS == 1
            {
                uint32_t v[32];                                   // bias columns (all 16 equal); x32 load stays inside the 512 columns
                tmem_ld32(tmem_base + lane_base + kBiasCol, v);
                tmem_ld_wait();
                if (slot) slot[kDetAccFloats + row] = __uint_as_float(v[0]);
                else if (w.dens_row1 && row == w.dens_row1 - 1 && a.dbias)
                    atomicAdd(a.dbias + (size_t)w.b * HN_BIAS_STRIDE + HN_BIAS_OFF_DENSITY, __uint_as_float(v[0]) * inv_scale);
                else if (row_ok && want_bias)
                    atomicAdd(a.dbias + (size_t)w.b * HN_BIAS_STRIDE + w.bias_off + row, __uint_as_float(v[0]) * inv_scale);
            }
#endif
            tc_fence_before_sync();
            mbar_arrive(smem_u32(&sh.acc_empty));
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                              // peers may still multicast into / signal this CTA
    if (warp == 1) tmem_free<512>(tmem_base);
}

// ---- deterministic mode, second pass: one CTA per (destination chunk of dW | bias row of one item), fixed item order
struct WDst {
    int32_t kind;             // 0: dW chunk (w_idx, row0, rows, columns from x_col / x_valid); 1: bias rows of item b
    int32_t w_idx, row0, rows, n_x, b, bias_off;
    int16_t x_col[kWMaxX], x_valid[kWMaxX];
    int32_t first, count;     // range in the slot list
    int32_t src_row;          // first accumulator row of the slices that feeds this destination (64 for the free density row)
};

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WArgs a, const WDst* dsts, const int* slot_list, const int bias_mode) {
    const WDst d = dsts[blockIdx.x];
    const float inv_scale = 1.0f / __ldg(a.grad_scale);
    const int* sl = slot_list + d.first;
    if (d.kind == 0) {
        const bool to_f = (d.w_idx == W_R1 && a.dwf != nullptr);
        float* dw = to_f ? a.dwf : a.dw[d.w_idx];
        const int ld = to_f ? HN_HIDDEN : a.ld[d.w_idx];
        for (int e = threadIdx.x; e < d.rows * d.n_x * 64; e += blockDim.x) {
            const int r = e / (d.n_x * 64), c = e % (d.n_x * 64), k = c >> 6, cc = c & 63;
            if (cc >= d.x_valid[k]) continue;
            float s = 0.f;
            for (int i = 0; i < d.count; ++i) s += a.partials[(size_t)sl[i] * kDetSlotFloats + (size_t)(d.src_row + r) * 448 + c];
            dw[(size_t)(d.row0 + r) * ld + d.x_col[k] + cc] += s * inv_scale;
        }
    } else if (d.kind == 1) {
        for (int r = threadIdx.x; r < d.rows; r += blockDim.x) {
            float s = 0.f;
            for (int i = 0; i < d.count; ++i) {
                const float* p = a.partials + (size_t)sl[i] * kDetSlotFloats + kDetAccFloats;
                s += bias_mode == 1 ? p[d.src_row + r] : (p[128 + d.src_row + r] + p[256 + d.src_row + r]);
            }
            a.dbias[(size_t)d.b * HN_BIAS_STRIDE + d.bias_off + r] += s * inv_scale;
        }
    }
}

static std::mutex g_w_mu;
static bool g_w_ready[64] = {};

// host: enumerate work items.
//   cluster : items of the 3-CTA multicast kernel (three consecutive entries = the three 128-channel chunks of one 384-wide
//             layer over one sample range);
//   duo     : items of the 2-CTA multicast kernel (two consecutive entries = the two chunks of RGB_layer_1 (128 + 64 rows) or
//             RGB_layer_2 (128 + 128) over one sample range) - each layer-input block is read from L2 once per pair instead of
//             once per chunk;
//   single  : everything else, one CTA per entry (bias-only passes, and all layers when clusters cannot be resident);
//   side    : single-CTA items for the SMs the resident 3-CTA clusters leave idle (148 - 3 * 45 = 13 on this part): they run on a
//             second stream while the cluster kernel runs.
static void build_items(const hn_mlp_bwd_weights_t& a, int n_sm, int n_clusters, int n_duos, int n_side, std::vector<WItem>& cluster,
                        std::vector<WItem>& single, std::vector<WItem>& duo, std::vector<WItem>& side) {
    const int tiles_per_item = (int)(((int64_t)a.n_rays * a.n_samples) / HN_TILE);
    bool want_w = false;
    for (int i = 0; i < 12; ++i) want_w = want_w || (a.dw[i] != nullptr);
    struct LayerW { int w_idx; int n_out; int g_blk; int g_dfeat; int bias_off; int x_slot; int n_xblk; bool pe; int x_col0; };
    std::vector<LayerW> layers;
    for (int l = 0; l < 8; ++l)
        layers.push_back({l, HN_HIDDEN, HN_GSLOT_Z0 + 6 * l, 0, l * HN_HIDDEN, l == 0 ? -1 : HN_SLOT_H0 + 6 * (l - 1), l == 0 ? 0 : 6,
                          l == 0 || l == 5, l == 5 ? a.l5_hidden_col : 0});
    // r0_fused (the fast chains): RGB_layer_0 is not a layer (hn_mlp_sched.h) - RGB_layer_1's layer input is FeaExt_module_7's output
    const bool fused = a.r0_fused != 0;
    if (!fused) layers.push_back({W_R0, HN_HIDDEN, HN_GSLOT_R0, 0, HN_BIAS_OFF_R0, HN_SLOT_H0 + 6 * 7, 6, false, 0});
    layers.push_back({W_R1, HN_RGB1, HN_GSLOT_R1, 0, HN_BIAS_OFF_R1, fused ? HN_SLOT_H0 + 6 * 7 : HN_SLOT_R0, 6, false, 0});
    layers.push_back({W_R2, HN_FEAT, 0, 1, HN_BIAS_OFF_R2, HN_SLOT_X, 3, false, 0});
    // The density head's weight gradient rides in RGB_layer_0's items (same layer input h7, read by the CUDA-core readers) whenever
    // that layer's own weight gradient is computed; otherwise it is a one-channel pseudo layer of its own.
    // Opt-in (HN_WGRAD_DENS_FOLD=1): measured on B200, the readers slow RGB_layer_0's stages by more than the pseudo layer costs (cluster
    // kernel 1.45 -> 1.79 ms against 0.095 ms for the separate density items), so the default keeps the pseudo layer.
    static const bool fold_env = [] { const char* e = getenv("HN_WGRAD_DENS_FOLD"); return e && atoi(e) != 0; }();
    const bool dens_in_r0 = fold_env && !fused && !a.det_workspace && want_w && a.dw[W_R0] != nullptr && a.dw[W_DENSITY] != nullptr;   // (the fold has no deterministic reduction)
    // With RGB_layer_0 folded away, RGB_layer_1's layer input IS h7, and the second gradient block of its 64-row second chunk is
    // the density head's gradient block (slot HN_GSLOT_R1 + 3 = HN_GSLOT_DENS, column 0 = dL/d(pre-ReLU density)): row 64 of that
    // chunk's accumulator already holds sum_m dsigma_m h7_m - the density head's weight gradient - and its bias column the bias
    // gradient.  The flush writes that row out and the one-channel pseudo layer (6 blocks of h7 re-read, an M = 128 MMA for one
    // row) is not needed.
    const bool dens_free = fused && want_w && a.dwf != nullptr && a.dw[W_R1] != nullptr && a.dw[W_DENSITY] != nullptr;
    static_assert(HN_GSLOT_R1 + 3 == HN_GSLOT_DENS, "the density gradient block must follow RGB_layer_1's three blocks");
    if (!dens_in_r0 && !dens_free) layers.push_back({W_DENSITY, 1, HN_GSLOT_DENS, 0, HN_BIAS_OFF_DENSITY, HN_SLOT_H0 + 6 * 7, 6, false, 0});
    auto active = [&](const LayerW& L) { return want_w || a.want_all_bias || L.w_idx == W_L0 || L.w_idx == W_L5 || L.w_idx == W_R1; };
    auto clustered = [&](const LayerW& L) { return want_w && n_clusters > 0 && L.n_out == HN_HIDDEN && a.dw[L.w_idx] != nullptr; };
    auto duoed = [&](const LayerW& L) { return want_w && n_duos > 0 && (L.w_idx == W_R1 || L.w_idx == W_R2) && a.dw[L.w_idx] != nullptr; };
    // A unit = one layer (cluster / duo lists: its chunks side by side) or one chunk (single list) of one batch item, over all of
    // the item's tiles, with a cost per tile ~ MMA cycles / operand bytes of a stage.  The units of a list are laid end to end and
    // cut into one equal-cost share per worker (a cluster or a CTA): a worker gets one or two sample ranges, every worker
    // finishes at the same time, and each accumulator is flushed (atomics) only once or twice.
    static const double slope = [] { const char* e = getenv("HN_WGRAD_SLOPE"); return e ? atof(e) : 40.0; }();   // tuning knob (cycles per operand block)
    auto stage_cost = [](int n_x) { return 1000.0 + slope * n_x; };
    struct Unit { std::vector<WItem> tmpl; int b; double cost; int t0, t1; };      // tiles [t0, t1) of batch item b
    std::vector<Unit> units_c, units_d, units_s;
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
        if (dens_free && L.w_idx == W_R1 && j == 1) w.dens_row1 = 64 + 1;
        if (dens_in_r0 && L.w_idx == W_R0) {                        // chunk j covers h7 columns [128 j, 128 j + 128)
            w.dens = (int16_t)(j == 0 ? 2 : 1);
            w.dens_x0 = (int16_t)(2 * j);
            w.dens_blk = HN_GSLOT_DENS;
        }
        return w;
    };
    for (const LayerW& L : layers) {
        if (!active(L)) continue;
        const bool cl = clustered(L), du = !cl && duoed(L);
        for (int b = 0; b < a.B; ++b) {
            if (cl || du) {
                Unit u; u.b = b; u.t0 = 0; u.t1 = tiles_per_item;
                for (int j = 0; j < (cl ? 3 : 2); ++j) u.tmpl.push_back(make_item(L, j, b));
                u.cost = stage_cost(u.tmpl[0].n_x);
                (cl ? units_c : units_d).push_back(u);
            } else {
                for (int j = 0; j * 128 < L.n_out; ++j) {
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
    if (n_side > 0 && !units_c.empty() && !units_d.empty()) {
        // share of the two-chunk layers that n_side SMs finish, as single-CTA items, in about the time the cluster kernel takes
        double cost_c = 0, cost_d = 0;
        for (const Unit& u : units_c) cost_c += u.cost * 3 * (u.t1 - u.t0);
        for (const Unit& u : units_d) cost_d += u.cost * 2 * (u.t1 - u.t0);
        static const double tune = [] { const char* e = getenv("HN_WGRAD_SIDE"); return e ? atof(e) : 0.85; }();
        double f = tune * (cost_c / (3.0 * n_clusters)) * n_side / cost_d;          // (time of the cluster phase) x n_side / (duo work)
        f = f > 0.9 ? 0.9 : f;
        std::vector<Unit> first, rest;
        for (const Unit& u : units_d) {
            const int cut = u.t0 + (int)((u.t1 - u.t0) * f);
            if (cut > u.t0)
                for (size_t j = 0; j < u.tmpl.size(); ++j) { Unit p; p.b = u.b; p.cost = u.cost; p.t0 = u.t0; p.t1 = cut; p.tmpl.push_back(u.tmpl[j]); first.push_back(p); }
            if (cut < u.t1) { Unit p = u; p.t0 = cut; rest.push_back(p); }
        }
        partition(first, n_side, 1, side);
        partition(rest, n_duos, 2, duo);
    } else {
        partition(units_d, n_duos, 2, duo);
    }
    partition(units_s, n_sm, 1, single);
}

static inline bool k_has_dw(const hn_mlp_bwd_weights_t& a, int w_idx) { return w_idx >= 0 && a.dw[w_idx] != nullptr; }

template <int CL>
static int launch_wgrad(const WArgs& k, int grid, cudaStream_t st) {
    if (CL == 1) {
        mlp_wgrad_kernel<1><<<grid, kWThreads, kWgradSmem, st>>>(k);
        return check_launch("hn_mlp_bwd_weights");
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kWThreads); cfg.dynamicSmemBytes = kWgradSmem; cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = CL; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_wgrad_kernel<CL>, k);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    return check_launch(CL == 3 ? "hn_mlp_bwd_weights (3-CTA clusters)" : "hn_mlp_bwd_weights (2-CTA clusters)");
}

template <int CL>
static int resident_clusters(int n_sm) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CL * (n_sm / CL)); cfg.blockDim = dim3(kWThreads); cfg.dynamicSmemBytes = kWgradSmem;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = CL; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, mlp_wgrad_kernel<CL>, &cfg) != cudaSuccess) { n = 0; cudaGetLastError(); }
    return n;
}

static int g_w_clusters[64] = {};
static int g_w_duos[64] = {};
static cudaStream_t g_w_side[64] = {};
static cudaEvent_t g_w_fork[64] = {}, g_w_join[64] = {};

}  // namespace hn

extern "C" int hn_mlp_bwd_weights(const hn_mlp_bwd_weights_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->act || !a->grads || !a->dfeat_image || !a->grad_scale || !a->status)
        return set_error(HN_E_BADARG, "hn_mlp_bwd_weights: null pointer");
    if (int rc = check_geometry(a->B, a->n_rays, a->n_samples, "hn_mlp_bwd_weights")) return rc;
    if (a->r0_fused && a->dw[W_R1] && !a->dwf)
        return set_error(HN_E_BADARG, "hn_mlp_bwd_weights: r0_fused needs the dwf buffer for RGB_layer_1's (fused) weight gradient");
    if (!a->items_workspace || a->items_workspace_bytes < hn_wgrad_workspace_bytes(a->B))
        return set_error(HN_E_BADARG, "hn_mlp_bwd_weights: items workspace missing or too small (hn_wgrad_workspace_bytes)");
    int dev = 0;
    cudaGetDevice(&dev);
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    {
        std::lock_guard<std::mutex> lk(g_w_mu);
        if (dev < 64 && !g_w_ready[dev]) {
            cudaError_t e = cudaFuncSetAttribute(mlp_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmem);
            if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
            // how many clusters can be resident at once (GPC sizes strand a few SMs); HN_WGRAD_CLUSTERS=0 / HN_WGRAD_DUOS=0: single-CTA items only
            const char* env = getenv("HN_WGRAD_CLUSTERS");
            g_w_clusters[dev] = (env && atoi(env) == 0) ? 0 : resident_clusters<3>(n_sm);
            const char* envd = getenv("HN_WGRAD_DUOS");
            g_w_duos[dev] = (envd && atoi(envd) == 0) ? 0 : resident_clusters<2>(n_sm);
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
    const int n_clusters = dev < 64 ? g_w_clusters[dev] : 0;
    const int n_duos = dev < 64 ? g_w_duos[dev] : 0;
    std::vector<WItem> cluster, single, duo, side;
    cudaStream_t side_stream = dev < 64 ? g_w_side[dev] : nullptr;
    const int n_idle = n_sm - 3 * n_clusters;
    const int n_side = (side_stream && n_clusters > 0 && n_duos > 0 && n_idle >= 4) ? n_idle : 0;
    build_items(*a, n_sm, n_clusters, n_duos, n_side, cluster, single, duo, side);
    if (cluster.empty() && single.empty() && duo.empty() && side.empty()) return HN_OK;
    if ((cluster.size() + single.size() + duo.size() + side.size()) * sizeof(WItem) > a->items_workspace_bytes)
        return set_error(HN_E_BADARG, "hn_mlp_bwd_weights: items workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<WItem> all(cluster);
    all.insert(all.end(), single.begin(), single.end());
    all.insert(all.end(), duo.begin(), duo.end());
    all.insert(all.end(), side.begin(), side.end());
    // ---- deterministic mode: a private slice per work item, and the list of slices per destination in item order
    const bool det = a->det_workspace != nullptr;
    std::vector<WDst> dsts;
    std::vector<int> slot_list;
    if (det) {
        int n_slots = 0;
        for (WItem& w : all) if (w.tile1 > w.tile0) w.slot = n_slots++;
        const size_t need = (size_t)n_slots * kDetSlotFloats * sizeof(float) + (all.size() + 64) * (sizeof(WDst) + sizeof(int));
        if (need > a->det_workspace_bytes) return set_error(HN_E_BADARG, "hn_mlp_bwd_weights: deterministic workspace too small (hn_wgrad_det_workspace_bytes)");
        // destinations in order of first appearance; their slices in item order (two passes: count, then fill)
        std::map<std::pair<int, int>, int> by_w, by_b;
        std::vector<std::vector<int>> lists;
        for (const WItem& w : all) {
            if (w.tile1 <= w.tile0) continue;
            if (k_has_dw(*a, w.w_idx)) {
                auto it = by_w.find({w.w_idx, w.row0});
                if (it == by_w.end()) {
                    WDst d{}; d.kind = 0; d.w_idx = w.w_idx; d.row0 = w.row0; d.rows = w.rows; d.n_x = w.n_x;
                    for (int q = 0; q < w.n_x; ++q) { d.x_col[q] = w.x_col[q]; d.x_valid[q] = w.x_valid[q]; }
                    it = by_w.emplace(std::make_pair((int)w.w_idx, (int)w.row0), (int)dsts.size()).first;
                    dsts.push_back(d); lists.emplace_back();
                }
                lists[it->second].push_back(w.slot);
            }
            if (w.dens_row1) {                                           // the free density row of RGB_layer_1's second chunk
                auto it = by_w.find({(int)W_DENSITY, 0});
                if (it == by_w.end()) {
                    WDst d{}; d.kind = 0; d.w_idx = W_DENSITY; d.row0 = 0; d.rows = 1; d.n_x = w.n_x; d.src_row = w.dens_row1 - 1;
                    for (int q = 0; q < w.n_x; ++q) { d.x_col[q] = w.x_col[q]; d.x_valid[q] = w.x_valid[q]; }
                    it = by_w.emplace(std::make_pair((int)W_DENSITY, 0), (int)dsts.size()).first;
                    dsts.push_back(d); lists.emplace_back();
                }
                lists[it->second].push_back(w.slot);
                if (a->dbias) {
                    auto ib = by_b.find({w.b, (int)HN_BIAS_OFF_DENSITY});
                    if (ib == by_b.end()) {
                        WDst d{}; d.kind = 1; d.b = w.b; d.bias_off = HN_BIAS_OFF_DENSITY; d.rows = 1; d.src_row = w.dens_row1 - 1;
                        ib = by_b.emplace(std::make_pair((int)w.b, (int)HN_BIAS_OFF_DENSITY), (int)dsts.size()).first;
                        dsts.push_back(d); lists.emplace_back();
                    }
                    lists[ib->second].push_back(w.slot);
                }
            }
            if (w.bias_off >= 0 && a->dbias) {
                auto it = by_b.find({w.b, w.bias_off});
                if (it == by_b.end()) {
                    WDst d{}; d.kind = 1; d.b = w.b; d.bias_off = w.bias_off; d.rows = w.rows;
                    it = by_b.emplace(std::make_pair((int)w.b, (int)w.bias_off), (int)dsts.size()).first;
                    dsts.push_back(d); lists.emplace_back();
                }
                lists[it->second].push_back(w.slot);
            }
        }
        for (size_t i = 0; i < dsts.size(); ++i) {
            dsts[i].first = (int)slot_list.size(); dsts[i].count = (int)lists[i].size();
            slot_list.insert(slot_list.end(), lists[i].begin(), lists[i].end());
        }
    }
    // the item table is tiny (<100 KiB); pageable -> device copy is stream-ordered and returns after staging
    cudaError_t e = cudaMemcpyAsync(a->items_workspace, all.data(), all.size() * sizeof(WItem), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
    WArgs k{};
    k.act = (const uint8_t*)a->act; k.grads = (const uint8_t*)a->grads; k.dfeat_image = (const uint8_t*)a->dfeat_image;
    k.grad_scale = a->grad_scale;
    for (int i = 0; i < 12; ++i) { k.dw[i] = a->dw[i]; k.ld[i] = a->ld[i]; }
    k.dwf = a->r0_fused ? a->dwf : nullptr;
    k.dbias = a->dbias;
    k.n_tiles = (int)(total_samples(a->B, a->n_rays, a->n_samples) / HN_TILE);
    k.status = a->status;
    const WDst* d_dsts = nullptr;
    const int* d_slots = nullptr;
    if (det) {
        // layout of the deterministic workspace: [destinations][slot list][slices]
        uint8_t* wsb = (uint8_t*)a->det_workspace;
        const size_t off_slots = (dsts.size() * sizeof(WDst) + 255) & ~(size_t)255;
        const size_t off_part = (off_slots + slot_list.size() * sizeof(int) + 255) & ~(size_t)255;
        e = cudaMemcpyAsync(wsb, dsts.data(), dsts.size() * sizeof(WDst), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(wsb + off_slots, slot_list.data(), slot_list.size() * sizeof(int), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        d_dsts = (const WDst*)wsb; d_slots = (const int*)(wsb + off_slots);
        k.partials = (float*)(wsb + off_part);
    }
    const WItem* base = (const WItem*)a->items_workspace;
    const bool forked = !side.empty() && !cluster.empty();
    // the side stream and its two events are per device, shared by all callers: the fork / launch / join sequence is serialised
    std::unique_lock<std::mutex> side_lock(g_w_mu, std::defer_lock);
    if (forked) side_lock.lock();
    if (forked && cudaEventRecord(g_w_fork[dev], st) != cudaSuccess) return set_error(HN_E_PROTOCOL, "hn_mlp_bwd_weights: event record failed");
    if (!cluster.empty()) {
        k.items = base; k.n_items = (int)cluster.size() / 3;
        if (int rc = launch_wgrad<3>(k, 3 * n_clusters, st)) return rc;      // the balanced schedule has one column per resident cluster
    }
    if (forked) {
        // the idle SMs' share, concurrently with the clusters (launched after them: it can only take what they leave free)
        cudaStreamWaitEvent(side_stream, g_w_fork[dev], 0);
        WArgs ks = k;
        ks.items = base + cluster.size() + single.size() + duo.size(); ks.n_items = (int)side.size();
        if (int rc = launch_wgrad<1>(ks, n_side, side_stream)) return rc;
        cudaEventRecord(g_w_join[dev], side_stream);
        cudaStreamWaitEvent(st, g_w_join[dev], 0);
    }
    if (forked) side_lock.unlock();
    if (!duo.empty()) {
        k.items = base + cluster.size() + single.size(); k.n_items = (int)duo.size() / 2;
        if (int rc = launch_wgrad<2>(k, 2 * n_duos, st)) return rc;
    }
    if (!single.empty()) {
        k.items = base + cluster.size(); k.n_items = (int)single.size();
        if (int rc = launch_wgrad<1>(k, n_sm, st)) return rc;
    }
    if (det && !dsts.empty()) {
        wgrad_reduce_kernel<<<(unsigned)dsts.size(), 256, 0, st>>>(k, d_dsts, d_slots, HN_WBIAS);
        if (int rc = check_launch("hn_mlp_bwd_weights (deterministic reduction)")) return rc;
    }
    return HN_OK;
}

extern "C" size_t hn_wgrad_det_workspace_bytes(int B) {
    // one slice per work item (same bound as the item table) + the destination / slot tables
    const size_t items = 200 * (size_t)(B > 0 ? B : 1) + 2000;
    return items * (hn::kDetSlotFloats * sizeof(float) + sizeof(hn::WDst) + sizeof(int)) + 4096;
}

extern "C" size_t hn_wgrad_workspace_bytes(int B) {
    // upper bound: per list (#workers + #units) pieces, padded to whole rounds: clusters 3 * 3 * (49 + 9 B), duos 2 * 2 * (74 + 2 B),
    // single 3 * (160 + 12 B), side (13 + 4 B)
    return (size_t)(200 * (size_t)(B > 0 ? B : 1) + 2000) * sizeof(hn::WItem);
}
