// hn_mlp_fwd.cu — fused sampling + positional encoding + fg_CD_predictor forward for sm_100a.
//
// One persistent CTA per SM walks 128-sample tiles.  Per tile the whole 11-GEMM chain runs on-chip:
//   epilogue warps  : build the tile's PE operand block from the camera (ray -> stratified sample -> sin/cos),
//                     then for every accumulator chunk: TMEM -> registers -> +bias, ReLU -> fp16 -> the
//                     activation buffer in shared memory (the next GEMM's A operand), plus the density head
//                     (fp32 dot product on CUDA cores) and the final [128 x 256] feature rows to HBM
//   MMA issuer      : one thread, tcgen05.mma (128 x N x 16, fp16 in, fp32 accumulate in TMEM), K-outer over
//                     128-column K chunks so a chunk's epilogue overlaps the next MMAs; 4 rotating 128-column
//                     accumulators
//   weight producer : one thread, streams the packed weight units L2 -> smem ring with cp.async.bulk
// Activations never leave the SM (unless saved for backward: then each finished operand block is also
// bulk-stored to HBM in the same image layout, together with 1-bit ReLU masks).
// Reference semantics: NetWorks/utils.py:43-51,147-161; NetWorks/models.py:62-87; HeadNeRFNet.py:139-152.
#include <mutex>
#include "hn_api.h"
#include "hn_mlp_common.cuh"
#include "hn_mlp_sched.h"
#include "hn_sample.cuh"

namespace hn {

__constant__ FwdTables c_fwd;

constexpr int kStages = 3;                                  // weight ring: 3 stages of two 64-wide K blocks (32 KiB)
constexpr uint32_t kStageBytes = 2 * kUnitBytes;
constexpr uint32_t kOffA = 0;                               // 6 activation blocks
constexpr uint32_t kOffPE = 6 * kUnitBytes;                 // 1 PE block
constexpr uint32_t kOffW = 7 * kUnitBytes;                  // weight ring
constexpr uint32_t kOffBias = 7 * kUnitBytes + kStages * kStageBytes;   // this item's effective bias row (15.3 KiB): the ~15 KiB
                                                                          // of L1 left beside the smem carve-out cannot hold it
constexpr uint32_t kBiasBytes = HN_BIAS_STRIDE * 4;
constexpr uint32_t kFwdSmem = kOffBias + kBiasBytes + 1024;
constexpr uint32_t kTmemCols = 512;

struct FwdShared {
    uint64_t w_full[6], w_empty[6], w_peer[6];
    uint64_t a_ready[3], pe_ready, acc_full[4], acc_empty[4];
    uint64_t a_ready_p[3], pe_ready_p, acc_empty_p[4];      // pair mode, leader only: one arrival each, forwarded by the peer CTA
    float dens[128];                    // density head: the four column groups add their partial dot products here
    float w_density[HN_HIDDEN];
    uint32_t tmem_base;
    volatile int abort;
};

// positional encoding (NetWorks/utils.py:20-51): columns [16*CG, 16*CG+16) of row `row` of the PE operand block.
// channel order: p(3), then per frequency 2^k: sin(3), cos(3); column 63 is the zero pad of the 64-wide K block.
template <int CG>
__device__ __forceinline__ void write_pe_part(uint32_t pe_block, int row, const float (&p)[3]) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        constexpr int c0 = 16 * CG;
        const int c = c0 + i;
        if (c < 3) v[i] = p[c];
        else if (c == 63) v[i] = 0.f;
        else {
            const int k = (c - 3) / 6, t = (c - 3) % 6;
            const float arg = p[t % 3] * (float)(1 << k);
            v[i] = (t < 3) ? sinf(arg) : cosf(arg);
        }
    }
    const uint32_t row_addr = pe_block + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
    for (int h = 0; h < 2; ++h)
        st_shared_v4(row_addr + (((2 * CG + h) ^ (row & 7)) << 4),
                     pack_h2(v[8 * h + 0], v[8 * h + 1]), pack_h2(v[8 * h + 2], v[8 * h + 3]),
                     pack_h2(v[8 * h + 4], v[8 * h + 5]), pack_h2(v[8 * h + 6], v[8 * h + 7]));
}

// PAIR = true: two CTAs of a cluster (one TPC) process two tiles in lockstep with cta_group::2 MMAs of M = 256; each
// CTA streams only HALF of every weight unit (its half of the B operand).  Work item w -> tile 2w + rank.
template <bool PAIR>
__global__ void __launch_bounds__(kFusedThreads, 1) mlp_fwd_kernel(const hn_mlp_fwd_t a, const int n_tiles, const int tiles_per_item) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ FwdShared sh;
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool saving = (a.act != nullptr);
    // weight ring geometry: single CTA 3 x 32 KiB (two 64-wide K blocks of 128 rows); pair mode 6 x 16 KiB (this CTA's
    // half of the rows of both K blocks) - same bytes in flight per CTA, i.e. twice the prefetch depth per weight byte needed
    constexpr int STAGES = PAIR ? 6 : 3;
    constexpr uint32_t STAGE_BYTES = PAIR ? kUnitBytes : 2 * kUnitBytes;
    constexpr uint32_t KB_STRIDE = PAIR ? kUnitBytes / 2 : kUnitBytes;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int work0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int work_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_work = PAIR ? n_tiles / 2 : n_tiles;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(&sh.w_full[i]), 1); mbar_init(smem_u32(&sh.w_empty[i]), 1); mbar_init(smem_u32(&sh.w_peer[i]), 1);
        }
        for (int i = 0; i < 3; ++i) { mbar_init(smem_u32(&sh.a_ready[i]), kEpiWarps); mbar_init(smem_u32(&sh.a_ready_p[i]), 1); }
        mbar_init(smem_u32(&sh.pe_ready), kEpiWarps); mbar_init(smem_u32(&sh.pe_ready_p), 1);
        for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&sh.acc_full[i]), 1); mbar_init(smem_u32(&sh.acc_empty[i]), kEpiWarps); mbar_init(smem_u32(&sh.acc_empty_p[i]), 1); }
        sh.abort = 0;
        mbar_fence_init();
    }
    if (warp == 2) { if (PAIR) tmem_alloc_pair<kTmemCols>(smem_u32(&sh.tmem_base)); else tmem_alloc<kTmemCols>(smem_u32(&sh.tmem_base)); }
    tc_fence_before_sync();
    __syncthreads();
    if (PAIR) cluster_sync_all();                               // peer barriers initialised before any remote arrive
    tc_fence_after_sync();
    const uint32_t tmem_base = sh.tmem_base;
    const int n_ops = c_fwd.n_ops;

    if (warp == 0) {
        // ======================= weight producer =======================
        if (lane == 0) {
            uint32_t uc = 0;
            const uint8_t* packed = (const uint8_t*)a.packed;
            for (int w = work0; w < n_work && !sh.abort; w += work_stride) {
                for (int u = 0; u < n_ops; ++u, ++uc) {
                    const uint32_t stage = uc % STAGES, par = (uc / STAGES) & 1;
                    if (!wait_or_abort(&sh.w_empty[stage], par ^ 1, &sh.abort, a.status, 101)) break;
                    const MmaOp op = c_fwd.mma[u];
                    // pair mode: this CTA holds rows [rank*N/2, rank*N/2 + N/2) of the unit (its half of the B operand)
                    const uint32_t bytes = (uint32_t)op.n8 * 8 * 128 / (PAIR ? 2 : 1);
                    const uint32_t fb = smem_u32(&sh.w_full[stage]);
                    mbar_arrive_expect_tx(fb, bytes * op.nkb);
                    for (int k = 0; k < op.nkb; ++k)
                        bulk_g2s(smem + kOffW + stage * STAGE_BYTES + k * KB_STRIDE,
                                 packed + (size_t)(op.unit + k) * kUnitBytes + (PAIR ? rank * bytes : 0), bytes, fb);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (PAIR && rank == 1) {
            // peer CTA: no MMA issue; relay "my half of the weights has landed" to the leader
            if (lane == 0) {
                uint32_t uc = 0;
                for (int w = work0; w < n_work && !sh.abort; w += work_stride)
                    for (int u = 0; u < n_ops; ++u, ++uc) {
                        const uint32_t stage = uc % STAGES, par = (uc / STAGES) & 1;
                        if (!wait_or_abort(&sh.w_full[stage], par, &sh.abort, a.status, 230)) break;
                        mbar_arrive_cluster(smem_u32(&sh.w_peer[stage]), 0);
                    }
            }
        } else {
            // the whole warp walks the schedule (uniform control flow keeps descriptors in uniform registers); one
            // elected lane issues the MMAs and commits
            uint32_t uc = 0, par_ready = 0, par_pe = 0, par_empty = 0;
            HN_PC_DECL(pc, 8);
            for (int w = work0; w < n_work && !sh.abort; w += work_stride) {
                MmaOp op = c_fwd.mma[0];
                for (int u = 0; u < n_ops; ++u, ++uc) {
                    const MmaOp nxt = c_fwd.mma[u + 1 < n_ops ? u + 1 : 0];       // table read off the critical path
                    bool ok = true;
                    HN_PC_T0(pc);
                    if (op.wait_src == 4) { ok = wait_or_abort(&sh.pe_ready, par_pe, &sh.abort, a.status, 201);
                        if (PAIR && ok) ok = wait_or_abort_x<true>(&sh.pe_ready_p, par_pe, &sh.abort, a.status, 241);
                        par_pe ^= 1; HN_PC_LAP(pc, 1); }
                    else if (op.wait_src) {
                        const int c = op.wait_src - 1;
                        ok = wait_or_abort(&sh.a_ready[c], (par_ready >> c) & 1, &sh.abort, a.status, 202 + c);
                        if (PAIR && ok) ok = wait_or_abort_x<true>(&sh.a_ready_p[c], (par_ready >> c) & 1, &sh.abort, a.status, 242 + c);
                        par_ready ^= 1u << c;
                        HN_PC_LAP(pc, 2);
                    }
                    if (ok && op.wait_empty) {
                        ok = wait_or_abort(&sh.acc_empty[op.q], ((par_empty >> op.q) & 1) ^ 1, &sh.abort, a.status, 210 + op.q);
                        if (PAIR && ok) ok = wait_or_abort_x<true>(&sh.acc_empty_p[op.q], ((par_empty >> op.q) & 1) ^ 1, &sh.abort, a.status, 246 + op.q);
                        par_empty ^= 1u << op.q;
                        HN_PC_LAP(pc, 3);
                    }
                    const uint32_t stage = uc % STAGES, par = (uc / STAGES) & 1;
                    if (ok) ok = wait_or_abort(&sh.w_full[stage], par, &sh.abort, a.status, 220);
                    HN_PC_LAP(pc, 4);
                    if (PAIR && ok) ok = wait_or_abort_x<true>(&sh.w_peer[stage], par, &sh.abort, a.status, 221);
                    HN_PC_LAP(pc, 5);
                    if (!ok) break;
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem + (op.a_blk == kPeBlk ? kOffPE : kOffA + op.a_blk * kUnitBytes);
                    const uint32_t b_addr = smem + kOffW + stage * STAGE_BYTES;
                    const uint32_t idesc = umma_idesc(PAIR ? 256 : 128, (uint32_t)op.n8 * 8, kF16, kF16, 0, 0);
                    const uint32_t d_addr = tmem_base + (uint32_t)op.tmem_col8 * 8;
                    const uint32_t a_lo = desc_lo(a_addr, 16), b_lo = desc_lo(b_addr, 16);
                    const uint32_t first = op.first, nkb = op.nkb;
                    if (elect_one()) {
                        for (uint32_t k = 0; k < nkb; ++k) {
#pragma unroll
                            for (uint32_t ks = 0; ks < 4; ++ks)
                                umma_lohi_x<PAIR>(d_addr, a_lo + k * (kUnitBytes >> 4) + ks * 2, b_lo + k * (KB_STRIDE >> 4) + ks * 2, idesc,
                                                  (first && k == 0 && ks == 0) ? 0u : 1u);
                        }
                        HN_PC_LAP(pc, 6);
                        umma_commit_x<PAIR>(smem_u32(&sh.w_empty[stage]));
                        if (op.commit) umma_commit_x<PAIR>(smem_u32(&sh.acc_full[op.q]));
                    }
                    __syncwarp();
                    HN_PC_LAP(pc, 7);
                    op = nxt;
                }
            }
            HN_PC_FLUSH(pc, 8, a.status + 2, blockIdx.x == 0);
        }
    } else if (warp == 3) {
        // ======================= pair mode, peer CTA: barrier forwarder =======================
        // The peer's epilogue warps arrive on their OWN CTA's barriers (cheap); this thread relays every completed phase
        // to the leader with a single remote arrive, keeping cluster-scope release traffic off the epilogue's critical path.
        if (PAIR && rank == 1 && lane == 0) {
            const int my_tiles = (n_work - work0 + work_stride - 1) / work_stride;
            uint64_t* local[8] = {&sh.a_ready[0], &sh.a_ready[1], &sh.a_ready[2], &sh.acc_empty[0], &sh.acc_empty[1], &sh.acc_empty[2], &sh.acc_empty[3], &sh.pe_ready};
            uint64_t* remote[8] = {&sh.a_ready_p[0], &sh.a_ready_p[1], &sh.a_ready_p[2], &sh.acc_empty_p[0], &sh.acc_empty_p[1], &sh.acc_empty_p[2], &sh.acc_empty_p[3], &sh.pe_ready_p};
            int left[8];
            for (int i = 0; i < 3; ++i) left[i] = c_fwd.n_ready[i] * my_tiles;
            for (int i = 0; i < 4; ++i) left[3 + i] = c_fwd.n_empty[i] * my_tiles;
            left[7] = my_tiles;
            uint32_t par = 0;
            int total = 0;
            for (int i = 0; i < 8; ++i) total += left[i];
            const long long t0 = clock64();
            while (total > 0 && !sh.abort) {
                for (int i = 0; i < 8; ++i) {
                    if (left[i] > 0 && mbar_try_wait(smem_u32(local[i]), (par >> i) & 1)) {
                        mbar_arrive_cluster(smem_u32(remote[i]), 0);
                        par ^= 1u << i; --left[i]; --total;
                    }
                }
                if (clock64() - t0 > 20000000000ll) { sh.abort = 1; atomicCAS(a.status, 0, 260); }
            }
        }
    } else if (warp >= kCtrlWarps) {
        // ======================= PE producer + epilogue: 16 warps, 32 rows x 32 columns each =======================
        const int ew = warp - kCtrlWarps;
        const int cg = ew >> 2;                                   // column group inside a chunk
        const int row = (ew & 3) * 32 + lane;                     // tile row = TMEM lane
        const uint32_t lane_base = (uint32_t)((ew & 3) * 32) << 16;
        uint32_t par_full = 0;
        const bool leader = (ew == 0 && lane == 0);
        const int pe_after = c_fwd.pe_after_epi;
        int cached_b = -1;
        for (int i = tid - kCtrlWarps * 32; i < HN_HIDDEN; i += kEpiThreads) sh.w_density[i] = __ldg(a.w_density + i);
        if (tid - kCtrlWarps * 32 < 128) sh.dens[tid - kCtrlWarps * 32] = 0.f;
        named_sync(3, kEpiThreads);

        // sampling + positional encoding of tile `t`: the first GEMM's operand is generated, not loaded
        auto produce_pe = [&](int t) {
            const size_t mm = (size_t)t * HN_TILE + row;
            const int bb = t / tiles_per_item;
            const int ns = a.cam.n_samples;
            const size_t ray_idx = mm / ns;
            const int s = (int)(mm % ns), r = (int)(ray_idx % a.cam.n_rays);
            const Ray ray = make_ray(a.cam, bb, r);
            const Sample q = make_sample(a.cam, ray, bb, r, s);
            if (cg == 0) {
                a.delta[mm] = q.zdist;
                if (a.zvals) a.zvals[mm] = q.zval;
            }
            if (saving) { if (leader) bulk_wait_read<1>(); named_sync(1, kEpiThreads); }
            const float p[3] = {q.px, q.py, q.pz};
            switch (cg) {
                case 0: write_pe_part<0>(smem + kOffPE, row, p); break;
                case 1: write_pe_part<1>(smem + kOffPE, row, p); break;
                case 2: write_pe_part<2>(smem + kOffPE, row, p); break;
                default: write_pe_part<3>(smem + kOffPE, row, p); break;
            }
            fence_async_smem();
            if (saving) {
                named_sync(1, kEpiThreads);
                if (leader) {
                    bulk_s2g((uint8_t*)a.act + ((size_t)HN_SLOT_PE * n_tiles + t) * kUnitBytes, smem + kOffPE, kUnitBytes);
                    bulk_commit();
                }
            }
            warp_arrive(smem_u32(&sh.pe_ready), lane);
        };

        auto tile_of = [&](int w) { return PAIR ? 2 * w + (int)rank : w; };
        HN_PC_DECL(ec, 6);
        if (work0 < n_work) produce_pe(tile_of(work0));
        for (int w = work0; w < n_work && !sh.abort; w += work_stride) {
            const int tile = tile_of(w);
            const size_t m = (size_t)tile * HN_TILE + row;
            const int b = tile / tiles_per_item;
            if (b != cached_b) {                                    // (re)load the item's bias row; epilogue warps run in lockstep
                named_sync(3, kEpiThreads);
                const float4* src = reinterpret_cast<const float4*>(a.bias + (size_t)b * HN_BIAS_STRIDE);
                for (int i = tid - kCtrlWarps * 32; i < HN_BIAS_STRIDE / 4; i += kEpiThreads) {
                    const float4 v4 = __ldg(src + i);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(smem + kOffBias + i * 16), "f"(v4.x), "f"(v4.y), "f"(v4.z), "f"(v4.w) : "memory");
                }
                named_sync(3, kEpiThreads);
                cached_b = b;
            }
            const uint32_t bias_row = smem + kOffBias;
            float dens = 0.f;
            for (int e = 0; e < kFwdEpis; ++e) {
                const EpiOp op = c_fwd.epi[e];
                // on a pipeline fault every later wait returns at once; the loop still runs to its end so that all
                // epilogue threads keep meeting at the same named barriers
                HN_PC_T0(ec);
                wait_or_abort(&sh.acc_full[op.q], (par_full >> op.q) & 1, &sh.abort, a.status, 300 + e);
                par_full ^= 1u << op.q;
                HN_PC_LAP(ec, 1);
                tc_fence_after_sync();
                const bool to_smem = (op.kind != EPI_FEAT);
                const bool active = cg < op.width32;
                const int col = cg * 32;                           // column inside the chunk
                if (to_smem && saving) { if (leader) bulk_wait_read<1>(); named_sync(1, kEpiThreads); }
                float y[32];
                if (active) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_base + (uint32_t)op.tmem_col8 * 8 + col, v);
                    tmem_ld_wait();
                    const uint32_t bp = bias_row + (op.bias_off + col) * 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 bb;
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w) : "r"(bp + i * 16));
                        y[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bb.x;
                        y[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bb.y;
                        y[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bb.z;
                        y[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bb.w;
                    }
                }
                HN_PC_LAP(ec, 2);
                tc_fence_before_sync();
                warp_arrive(smem_u32(&sh.acc_empty[op.q]), lane);  // accumulator read: hand it back to the MMA issuer
                if (active) {
                    if (op.kind == EPI_HIDDEN) {
                        if (a.masks && op.mask_word != 0xFFFF)
                            a.masks[m * HN_MASK_WORDS + op.mask_word + cg] = positive_mask32(y);
                        if (op.density) {                          // density head on the fp32 activations (models.py:78,83)
                            const float4* wp = reinterpret_cast<const float4*>(sh.w_density + op.col0 + col);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 ww = wp[i];
                                dens = fmaf(fmaxf(y[4 * i + 0], 0.f), ww.x, dens); dens = fmaf(fmaxf(y[4 * i + 1], 0.f), ww.y, dens);
                                dens = fmaf(fmaxf(y[4 * i + 2], 0.f), ww.z, dens); dens = fmaf(fmaxf(y[4 * i + 3], 0.f), ww.w, dens);
                            }
                        }
                        store_row32<true>(smem + kOffA + op.dst_blk * kUnitBytes, row, col, y);
                    } else if (op.kind == EPI_LINEAR) {
                        store_row32<false>(smem + kOffA + op.dst_blk * kUnitBytes, row, col, y);
                    } else if (a.feat) {
                        float4* dst = reinterpret_cast<float4*>(a.feat + m * HN_FEAT + op.col0 + col);
#pragma unroll
                        for (int i = 0; i < 8; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                    }
                }
                HN_PC_LAP(ec, 3);
                if (op.density == 2) {                             // combine the four column groups' partial dot products
                    atomicAdd(&sh.dens[row], dens);
                    dens = 0.f;
                    named_sync(2, kEpiThreads);
                    if (cg == 0) {
                        a.sigma[m] = fmaxf(sh.dens[row] + __ldg(a.bias + (size_t)b * HN_BIAS_STRIDE + HN_BIAS_OFF_DENSITY), 0.f);
                        sh.dens[row] = 0.f;                        // next use is a whole tile (and many barrier hops) away
                    }
                }
                if (to_smem) {
                    fence_async_smem();
                    if (saving) {
                        named_sync(1, kEpiThreads);
                        if (leader && op.save_blk != 0xFFFF) {
                            const int nblk = (op.width32 + 1) / 2;
                            for (int k = 0; k < nblk; ++k)
                                bulk_s2g((uint8_t*)a.act + ((size_t)(op.save_blk + k) * n_tiles + tile) * kUnitBytes,
                                         smem + kOffA + (op.dst_blk + k) * kUnitBytes, kUnitBytes);
                            bulk_commit();
                        }
                    }
                    warp_arrive(smem_u32(&sh.a_ready[op.ready_idx]), lane);
                }
                // FeaExt_module_5 (the last reader of the PE block) is done: build the NEXT tile's PE operand now, while
                // the tensor core still has this tile's remaining layers queued
                HN_PC_LAP(ec, 4);
                if (e == pe_after && w + work_stride < n_work) produce_pe(tile_of(w + work_stride));
                HN_PC_LAP(ec, 5);
            }
        }
        if (saving && leader) bulk_wait_all<0>();
        HN_PC_FLUSH(ec, 6, a.status + 18, blockIdx.x == 0 && leader);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (PAIR) cluster_sync_all();                               // the peer may still be reading this CTA's operands / barriers
    if (warp == 2) { if (PAIR) tmem_free_pair<kTmemCols>(tmem_base); else tmem_free<kTmemCols>(tmem_base); }
}

static std::mutex g_fwd_mu;
static bool g_fwd_ready[64] = {};

}  // namespace hn

extern "C" int hn_mlp_fwd(const hn_mlp_fwd_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->cam.xy || !a->cam.Rmats || !a->cam.Tvecs || !a->cam.inv_inmats || !a->bias || !a->w_density ||
        !a->packed || !a->sigma || !a->delta || !a->status)
        return set_error(HN_E_BADARG, "hn_mlp_fwd: null pointer");
    if (int rc = check_geometry(a->cam.B, a->cam.n_rays, a->cam.n_samples, "hn_mlp_fwd")) return rc;
    if ((a->act == nullptr) != (a->masks == nullptr))
        return set_error(HN_E_BADARG, "hn_mlp_fwd: act and masks must both be given (backward) or both be NULL");
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_fwd_mu);
        if (dev < 64 && !g_fwd_ready[dev]) {
            cudaError_t e = cudaMemcpyToSymbol(c_fwd, &host_schedules().fwd, sizeof(FwdTables));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
            if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
            g_fwd_ready[dev] = true;
        }
    }
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int64_t M = total_samples(a->cam.B, a->cam.n_rays, a->cam.n_samples);
    const int n_tiles = (int)(M / HN_TILE);
    const int tiles_per_item = (int)(((int64_t)a->cam.n_rays * a->cam.n_samples) / HN_TILE);
    if (use_cta_pairs(n_tiles)) {
        const int n_pairs = (n_tiles / 2) < (n_sm / 2) ? (n_tiles / 2) : (n_sm / 2);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * n_pairs); cfg.blockDim = dim3(kFusedThreads); cfg.dynamicSmemBytes = kFwdSmem; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, mlp_fwd_kernel<true>, *a, n_tiles, tiles_per_item);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        return check_launch("hn_mlp_fwd (cta pairs)");
    }
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    mlp_fwd_kernel<false><<<grid, kFusedThreads, kFwdSmem, (cudaStream_t)stream>>>(*a, n_tiles, tiles_per_item);
    return check_launch("hn_mlp_fwd");
}
