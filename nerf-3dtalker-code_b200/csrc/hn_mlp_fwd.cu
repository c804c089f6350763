// hn_mlp_fwd.cu — fused sampling + positional encoding + fg_CD_predictor forward for sm_100a.
//
// One persistent CTA per SM walks 128-sample tiles.  Per tile the whole 11-GEMM chain runs on-chip and the
// activations never leave TENSOR MEMORY: every GEMM is a tcgen05.mma with the A operand (128 samples x K) read
// from TMEM and the B operand (weights) streamed through shared memory, so shared-memory bandwidth carries only
// the weights (an A operand in shared memory would double the operand traffic and cap the MMA rate at ~60 %).
//   epilogue warps   : build the tile's PE operand block from the camera (ray -> stratified sample -> sin/cos) in
//                      shared memory (the only A operand read from there), then for every accumulator chunk:
//                      TMEM -> registers -> +bias, ReLU -> packed f16 pairs -> TMEM slot that the next GEMM reads
//                      as its A operand; density head (fp32 dot product) and the final feature rows go to HBM
//   MMA issuer       : one elected thread, tcgen05.mma 128 x N x 16 (f16 in, fp32 accumulate), chunk by chunk over
//                      all K blocks (N-outer), two accumulator chunks in flight so a chunk's epilogue overlaps the
//                      next chunk's MMAs; TMEM slots per csrc/hn_mlp_sched.h
//   weight producers : bulk copies of one thread complete one after the other (~700 cycles each under load), so
//                      up to three threads of different warps stream the 16 KiB weight units L2 -> smem ring
//   saver            : (backward needed) every finished activation chunk is also staged in shared memory as an
//                      operand image and bulk-stored to HBM, with 1-bit ReLU masks written by the epilogue
// Reference semantics: NetWorks/utils.py:43-51,147-161; NetWorks/models.py:62-87; HeadNeRFNet.py:139-152.
#include <mutex>
#include "hn_api.h"
#include "hn_mlp_common.cuh"
#include "hn_mlp_sched.h"
#include "hn_sample.cuh"

namespace hn {

__constant__ FwdTables c_fwd;

constexpr int kStages = 8;                                  // weight ring: one 16 KiB unit per stage
constexpr uint32_t kOffPE = 0;                              // PE operand block
constexpr uint32_t kOffW = kUnitBytes;                      // weight ring
constexpr uint32_t kOffStg = kOffW + kStages * kUnitBytes;  // 2 staging buffers of two blocks (saved activations)
constexpr uint32_t kOffBias = kOffStg + 4 * kUnitBytes;     // this item's effective bias row
constexpr uint32_t kBiasBytes = HN_BIAS_STRIDE * 4;
constexpr uint32_t kOffShared = kOffBias + kBiasBytes;      // barriers and small arrays (FwdShared) close the dynamic region
constexpr uint32_t kTmemCols = 512;
constexpr int kGroupWarps = kEpiWarps / 2;                  // two epilogue groups take alternate accumulator chunks

struct FwdShared {
    uint64_t w_full[kStages], w_empty[kStages];
    uint64_t a_ready[3], pe_ready, pe_free, acc_full[2], acc_empty[2];
    uint64_t stg_full[2], stg_free[2];  // one staging buffer per epilogue group
    uint64_t ld_done[2][4], dens_done;  // accumulator loads of a group's chunk k (rotating over 4); density partials of a tile
    alignas(16) float dens[2][128];     // density head: every warp adds its partial dot products here (by tile parity)
    alignas(16) float w_density[HN_HIDDEN];   // read as float4
    uint32_t tmem_base;
    volatile int abort;
};
constexpr uint32_t kFwdSmem = kOffShared + sizeof(FwdShared);
static_assert(kOffShared % 16 == 0 && kFwdSmem <= 232448, "shared-memory budget (227 KiB per CTA)");

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

// sampling + positional encoding of row `row` of tile `t`: the first GEMM's operand is generated, not loaded.  `part` selects
// 16 of the 64 PE columns.  Kept out of line: the accurate sin/cos paths need registers and a little local memory that the
// epilogue loop around the call should not pay for.
__device__ __noinline__ void produce_pe_part(const hn_camera_t cam, float* delta, float* zvals, uint32_t pe_block, int t, int tiles_per_item,
                                             int row, int part, bool aux) {
    const size_t mm = (size_t)t * HN_TILE + row;
    const int bb = t / tiles_per_item;
    const int ns = cam.n_samples;
    const size_t ray_idx = mm / ns;
    const int s = (int)(mm % ns), r = (int)(ray_idx % cam.n_rays);
    const Ray ray = make_ray(cam, bb, r);
    const Sample q = make_sample(cam, ray, bb, r, s);
    if (aux) {
        delta[mm] = q.zdist;
        if (zvals) zvals[mm] = q.zval;
    }
    const float p[3] = {q.px, q.py, q.pz};
    switch (part) {
        case 0: write_pe_part<0>(pe_block, row, p); break;
        case 1: write_pe_part<1>(pe_block, row, p); break;
        case 2: write_pe_part<2>(pe_block, row, p); break;
        default: write_pe_part<3>(pe_block, row, p); break;
    }
}

__global__ void __launch_bounds__(kFusedThreads, 1) mlp_fwd_kernel(const hn_mlp_fwd_t a, const int n_tiles, const int tiles_per_item) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];      // no static shared memory in this kernel: the window starts aligned
    FwdShared& sh = *reinterpret_cast<FwdShared*>(smem_raw + kOffShared);
    const uint32_t smem = smem_u32(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool saving = (a.act != nullptr);
    const int work0 = (int)blockIdx.x, work_stride = (int)gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(&sh.w_full[i]), 1); mbar_init(smem_u32(&sh.w_empty[i]), 1); }
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&sh.a_ready[i]), kGroupWarps);
        mbar_init(smem_u32(&sh.pe_ready), kEpiWarps); mbar_init(smem_u32(&sh.pe_free), 1);
        mbar_init(smem_u32(&sh.dens_done), kEpiWarps);
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&sh.acc_full[i]), 1); mbar_init(smem_u32(&sh.acc_empty[i]), kGroupWarps);
            mbar_init(smem_u32(&sh.stg_full[i]), kGroupWarps); mbar_init(smem_u32(&sh.stg_free[i]), 1);
            for (int k = 0; k < 4; ++k) mbar_init(smem_u32(&sh.ld_done[i][k]), kGroupWarps);
        }
        sh.abort = 0;
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(smem_u32(&sh.tmem_base));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = sh.tmem_base;
    const int n_ops = c_fwd.n_ops;

    if (warp == 0 || warp == 2 || (warp == 3 && !saving)) {
        // ======================= weight producers (lane 0 of warps 0, 2 and - when nothing is saved - 3) =======================
        if (lane == 0) {
            const uint32_t P = saving ? 2u : 3u, p = warp == 0 ? 0u : (uint32_t)(warp - 1);
            uint32_t uc = 0;
            const uint8_t* packed = (const uint8_t*)a.packed;
            for (int w = work0; w < n_tiles && !sh.abort; w += work_stride) {
                for (int u = 0; u < n_ops; ++u, ++uc) {
                    if (uc % P != p) continue;
                    const uint32_t stage = uc % kStages, par = (uc / kStages) & 1;
                    if (!wait_or_abort(&sh.w_empty[stage], par ^ 1, &sh.abort, a.status, 101)) break;
                    const uint32_t bytes = (uint32_t)c_fwd.mma[u].n8 * 8 * 128;
                    const uint32_t fb = smem_u32(&sh.w_full[stage]);
                    mbar_arrive_expect_tx(fb, bytes);
                    bulk_g2s(smem + kOffW + stage * kUnitBytes, packed + (size_t)u * kUnitBytes, bytes, fb);
                }
            }
        }
    } else if (warp == 3) {
        // ======================= saver: staged activation chunks + the PE block -> HBM operand images =======================
        if (lane == 0) {
            uint32_t use0 = 0, use1 = 0, par_pe = 0, base_n = 0;
            for (int w = work0; w < n_tiles && !sh.abort; w += work_stride, base_n += kFwdEpis) {
                const int tile = w;
                if (!wait_or_abort(&sh.pe_ready, par_pe, &sh.abort, a.status, 150)) break;
                par_pe ^= 1;
                bulk_s2g((uint8_t*)a.act + ((size_t)HN_SLOT_PE * n_tiles + tile) * kUnitBytes, smem + kOffPE, kUnitBytes);
                bulk_commit();
                bulk_wait_read<0>();
                mbar_arrive(smem_u32(&sh.pe_free));
                for (int e = 0; e < kFwdEpis; ++e) {
                    const EpiOp2 op = c_fwd.epi[e];
                    if (op.save_blk == 0xFFFF) continue;
                    const uint32_t sg = (base_n + e) & 1;             // staging buffer = epilogue group of the chunk
                    if (!wait_or_abort(&sh.stg_full[sg], (sg ? use1 : use0) & 1, &sh.abort, a.status, 151)) break;
                    if (sg) ++use1; else ++use0;
                    const int nblk = (op.width32 + 1) / 2;
                    for (int k = 0; k < nblk; ++k)
                        bulk_s2g((uint8_t*)a.act + ((size_t)(op.save_blk + k) * n_tiles + tile) * kUnitBytes,
                                 smem + kOffStg + (sg * 2 + k) * kUnitBytes, kUnitBytes);
                    bulk_commit();
                    bulk_wait_read<0>();                               // the buffer is free as soon as the engine has read it
                    mbar_arrive(smem_u32(&sh.stg_free[sg]));
                }
            }
            bulk_wait_all<0>();
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        // the whole warp walks the schedule (uniform control flow keeps descriptors in uniform registers); one
        // elected lane issues the MMAs and commits
        uint32_t uc = 0, par_ready = 0, par_pe = 0, chunk_n = 0;
        HN_PC_DECL(pc, 8);
        for (int w = work0; w < n_tiles && !sh.abort; w += work_stride) {
            MmaOp2 op = c_fwd.mma[0];
            for (int u = 0; u < n_ops; ++u, ++uc) {
                const MmaOp2 nxt = c_fwd.mma[u + 1 < n_ops ? u + 1 : 0];         // table read off the critical path
                bool ok = true;
                HN_PC_T0(pc);
                if (op.wait_src == 4) { ok = wait_or_abort(&sh.pe_ready, par_pe, &sh.abort, a.status, 201); par_pe ^= 1; HN_PC_LAP(pc, 1); }
                else if (op.wait_src) {
                    const int c = op.wait_src - 1;
                    ok = wait_or_abort(&sh.a_ready[c], (par_ready >> c) & 1, &sh.abort, a.status, 202 + c);
                    par_ready ^= 1u << c;
                    HN_PC_LAP(pc, 2);
                }
                // accumulator of chunk n: released by the epilogue of chunk n-2 (the first two chunks find fresh barriers)
                if (ok && op.first) ok = wait_or_abort(&sh.acc_empty[chunk_n & 1], ((chunk_n >> 1) & 1) ^ 1, &sh.abort, a.status, 210);
                HN_PC_LAP(pc, 3);
                const uint32_t stage = uc % kStages, par = (uc / kStages) & 1;
                if (ok) ok = wait_or_abort(&sh.w_full[stage], par, &sh.abort, a.status, 220);
                HN_PC_LAP(pc, 4);
                if (!ok) break;
                tc_fence_after_sync();
                const uint32_t idesc = umma_idesc(128, (uint32_t)op.n8 * 8, kF16, kF16, 0, 0);
                const uint32_t d_addr = tmem_base + op.acc_col;
                const uint32_t b_lo = desc_lo(smem + kOffW + stage * kUnitBytes, 16);
                const uint32_t first = op.first;
                if (op.a_src & kSrcSmem) {
                    const uint32_t a_lo = desc_lo(smem + kOffPE + (op.a_src & 0x7FFFu) * kUnitBytes, 16);
                    if (elect_one()) {
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_lohi(d_addr, a_lo + ks * 2, b_lo + ks * 2, idesc, (first && ks == 0) ? 0u : 1u);
                    }
                } else {
                    const uint32_t a_t = tmem_base + op.a_src;
                    if (elect_one()) {
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ts_lo(d_addr, a_t + ks * 8, b_lo + ks * 2, idesc, (first && ks == 0) ? 0u : 1u);
                    }
                }
                HN_PC_LAP(pc, 6);
                if (elect_one()) {
                    umma_commit(smem_u32(&sh.w_empty[stage]));
                    if (op.commit) umma_commit(smem_u32(&sh.acc_full[chunk_n & 1]));
                }
                __syncwarp();
                HN_PC_LAP(pc, 7);
                chunk_n += op.commit;
                op = nxt;
            }
        }
        HN_PC_FLUSH(pc, 8, a.status + 2, blockIdx.x == 0 && lane == 0);
    } else {
        // ======================= PE producers + epilogue =======================
        // Two groups of 8 warps take alternate accumulator chunks (group = chunk index & 1), so one group's TMEM loads /
        // conversions / stores overlap the other's.  Inside a group: warp = (TMEM lane quarter, column half); a warp drains
        // 32 rows x 64 columns of its chunk as two 32-column pieces.
        const int ew = warp - kCtrlWarps;
        const int g = ew >> 3, quarter = ew & 3, half = (ew >> 2) & 1;
        const int row = quarter * 32 + lane;                        // tile row = TMEM lane
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const int qbar = 4 + g * 4 + quarter;                       // named barrier of the two warps sharing (group, quarter)
        uint32_t base_n = 0, use_n = 0, pe_n = 0, tile_i = 0;       // first chunk index of the tile; staged chunks of this group
        const int pe_after = c_fwd.pe_after_epi;
        int cached_b = -1;
        for (int i = tid - kCtrlWarps * 32; i < HN_HIDDEN; i += kEpiThreads) sh.w_density[i] = __ldg(a.w_density + i);
        if (tid - kCtrlWarps * 32 < 256) (&sh.dens[0][0])[tid - kCtrlWarps * 32] = 0.f;
        named_sync(3, kEpiThreads);

        // each group writes 32 of the 64 PE columns of the next tile (16 per warp half)
        auto produce_pe = [&](int t) {
            // the previous tile's PE block must have been read by the saver's bulk store (its MMAs are long done)
            if (saving) wait_or_abort(&sh.pe_free, (pe_n & 1) ^ 1, &sh.abort, a.status, 160);
            ++pe_n;
            produce_pe_part(a.cam, a.delta, a.zvals, smem + kOffPE, t, tiles_per_item, row, 2 * g + half, g == 0 && half == 0);
            fence_async_smem();
            warp_arrive(smem_u32(&sh.pe_ready), lane);
        };

        HN_PC_DECL(ec, 16);
        if (work0 < n_tiles) produce_pe(work0);
        float dens = 0.f;
        for (int w = work0; w < n_tiles && !sh.abort; w += work_stride, base_n += kFwdEpis, ++tile_i) {
            const int tile = w;
            const size_t m = (size_t)tile * HN_TILE + row;
            const int b = tile / tiles_per_item;
            if (b != cached_b) {                                    // (re)load the item's bias row: both groups meet here
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
            // the group that drains FeaExt_module_5's last chunk builds its half of the next PE block right after it, the
            // other group after its next own chunk (by then that chunk's MMAs, the last readers of the PE block, are done)
            const int pe_e = (((base_n + pe_after) & 1) == (uint32_t)g) ? pe_after : pe_after + 1;
            for (int e = 0; e < kFwdEpis; ++e) {
                const uint32_t n = base_n + e;
                if ((n & 1) != (uint32_t)g) continue;
                const EpiOp2 op = c_fwd.epi[e];
                // on a pipeline fault every later wait returns at once; the loop still runs to its end so that all
                // epilogue threads keep meeting at the same named barriers
                HN_PC_T0(ec);
                wait_or_abort(&sh.acc_full[g], (n >> 1) & 1, &sh.abort, a.status, 300 + e);
                HN_PC_LAP(ec, 1);
                tc_fence_after_sync();
                const int col = half * 64;                         // first column of this warp inside the chunk
                const bool act0 = 2 * half < op.width32, act1 = 2 * half + 1 < op.width32;
                const bool save = saving && op.save_blk != 0xFFFF;
                const uint32_t bp = bias_row + (op.bias_off + col) * 4;
                const uint32_t acc_addr = tmem_base + lane_base + op.acc_col + col;
                // after the accumulator loads: both warps of this lane quarter have read their part, so packed outputs may be
                // stored over it (in-place slots) and the MMA issuer may reuse it two chunks later
                auto release_acc = [&]() {
                    tc_fence_before_sync();
                    named_sync(qbar, 64);
                    tc_fence_after_sync();
                    warp_arrive(smem_u32(&sh.acc_empty[g]), lane);
                    warp_arrive(smem_u32(&sh.ld_done[g][(n >> 1) & 3]), lane);
                    if (op.wait_prev && n > 0) {                   // output slot drained by chunk n-1: the other group's loads
                        wait_or_abort(&sh.ld_done[g ^ 1][((n - 1) >> 1) & 3], ((n - 1) >> 3) & 1, &sh.abort, a.status, 330);
                        tc_fence_after_sync();
                    }
                    // staging buffer of this group (saved chunk, or scratch of the final-feature store): previous bulk read done?
                    if (save || (saving && op.kind == EPI_FEAT)) wait_or_abort(&sh.stg_free[g], (use_n & 1) ^ 1, &sh.abort, a.status, 340);
                };
                if (op.kind == EPI_FEAT) {
                    // final features: stage the 32 rows x 32 columns of each piece in shared memory (swizzled 16-byte chunks),
                    // then write whole 128-byte row segments (8 lanes each) instead of 32 scattered 16-byte pieces
                    const uint32_t stg = smem + kOffStg + (uint32_t)ew * 4096;                  // 32 rows x 128 B per warp
                    float* gbase = a.feat + ((size_t)tile * HN_TILE + quarter * 32) * HN_FEAT + op.col0 + col;
#pragma unroll
                    for (int pc = 0; pc < 2; ++pc) {
                        uint32_t v[32];
                        tmem_ld32(acc_addr + pc * 32, v);
                        tmem_ld_wait();
                        if (pc == 1) { HN_PC_LAP(ec, 2); release_acc(); HN_PC_LAP(ec, 3); }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 bb;
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w) : "r"(bp + pc * 128 + i * 16));
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stg + lane * 128 + ((i ^ (lane & 7)) << 4)),
                                         "f"(__uint_as_float(v[4 * i + 0]) + bb.x), "f"(__uint_as_float(v[4 * i + 1]) + bb.y),
                                         "f"(__uint_as_float(v[4 * i + 2]) + bb.z), "f"(__uint_as_float(v[4 * i + 3]) + bb.w) : "memory");
                        }
                        __syncwarp();
                        if (a.feat) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int rr = 4 * j + (lane >> 3), ch = lane & 7;
                                float4 o;
                                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                                             : "r"(stg + rr * 128 + ((ch ^ (rr & 7)) << 4)));
                                *reinterpret_cast<float4*>(gbase + (size_t)rr * HN_FEAT + pc * 32 + ch * 4) = o;
                            }
                        }
                        __syncwarp();
                    }
                    HN_PC_LAP(ec, 4);
                } else {
                    // piece 0 is converted before piece 1 is loaded (registers); its packed columns are stored after the barrier
                    uint32_t pk0[16];
                    auto convert = [&](const uint32_t (&v)[32], int pc, uint32_t (&pk)[16]) {
                        float y[32];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 bb;
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w) : "r"(bp + pc * 128 + i * 16));
                            y[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bb.x;
                            y[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bb.y;
                            y[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bb.z;
                            y[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bb.w;
                        }
                        if (op.kind == EPI_HIDDEN) {
                            if (a.masks && op.mask_word != 0xFFFF)
                                a.masks[m * HN_MASK_WORDS + op.mask_word + 2 * half + pc] = positive_mask32(y);
                            if (op.density) {                      // density head on the fp32 activations (models.py:78,83)
                                const float4* wp = reinterpret_cast<const float4*>(sh.w_density + op.col0 + col + pc * 32);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 ww = wp[i];
                                    dens = fmaf(fmaxf(y[4 * i + 0], 0.f), ww.x, dens); dens = fmaf(fmaxf(y[4 * i + 1], 0.f), ww.y, dens);
                                    dens = fmaf(fmaxf(y[4 * i + 2], 0.f), ww.z, dens); dens = fmaf(fmaxf(y[4 * i + 3], 0.f), ww.w, dens);
                                }
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) pk[i] = pack_relu_sat(y[2 * i], y[2 * i + 1]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) pk[i] = pack_sat(y[2 * i], y[2 * i + 1]);
                        }
                    };
                    const uint32_t out_addr = tmem_base + lane_base + op.out_col + half * 32;
                    const uint32_t stg = smem + kOffStg + (uint32_t)g * 2 * kUnitBytes;
                    if (act0) {
                        uint32_t v[32];
                        tmem_ld32(acc_addr, v);
                        tmem_ld_wait();
                        convert(v, 0, pk0);
                    }
                    uint32_t v1[32];
                    if (act1) { tmem_ld32(acc_addr + 32, v1); tmem_ld_wait(); }
                    HN_PC_LAP(ec, 2);
                    release_acc();
                    HN_PC_LAP(ec, 3);
                    if (act0) {
                        tmem_st16(out_addr, pk0);
                        if (save) store_row_packed(stg, row, col, pk0);
                    }
                    if (act1) {
                        uint32_t pk1[16];
                        convert(v1, 1, pk1);
                        tmem_st16(out_addr + 16, pk1);
                        if (save) store_row_packed(stg, row, col + 32, pk1);
                    }
                    HN_PC_LAP(ec, 4);
                    tmem_st_wait();
                    HN_PC_LAP(ec, 5);
                }
                if (save) { fence_async_smem(); warp_arrive(smem_u32(&sh.stg_full[g]), lane); ++use_n; }
                if (op.ready_idx != 255) {
                    tc_fence_before_sync();
                    warp_arrive(smem_u32(&sh.a_ready[op.ready_idx]), lane);
                }
                HN_PC_LAP(ec, 6);
                if (op.density) {
                    // this warp's last density chunk of the tile: publish its partial dot products; the group that owns the
                    // layer's last chunk adds the bias and writes sigma once all 16 warps have contributed
                    const bool last_mine = (op.density == 2) || (e + 2 >= kFwdEpis) || !c_fwd.epi[e + 2].density;
                    if (last_mine) {
                        atomicAdd(&sh.dens[tile_i & 1][row], dens);
                        dens = 0.f;
                        __threadfence_block();
                        warp_arrive(smem_u32(&sh.dens_done), lane);
                    }
                    if (op.density == 2) {
                        wait_or_abort(&sh.dens_done, tile_i & 1, &sh.abort, a.status, 350);
                        if (half == 0) {
                            const float tot = *((volatile float*)&sh.dens[tile_i & 1][row]);
                            a.sigma[m] = fmaxf(tot + __ldg(a.bias + (size_t)b * HN_BIAS_STRIDE + HN_BIAS_OFF_DENSITY), 0.f);
                            sh.dens[tile_i & 1][row] = 0.f;        // next use is two tiles away
                        }
                    }
                }
                HN_PC_LAP(ec, 7);
                if (e == pe_e && w + work_stride < n_tiles) produce_pe(w + work_stride);
                HN_PC_LAP(ec, 8);
            }
        }
        HN_PC_FLUSH(ec, 16, a.status + 18, blockIdx.x == 0 && ew == 0 && lane == 0);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_free<kTmemCols>(tmem_base);
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
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
            if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
            g_fwd_ready[dev] = true;
        }
    }
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int64_t M = total_samples(a->cam.B, a->cam.n_rays, a->cam.n_samples);
    const int n_tiles = (int)(M / HN_TILE);
    const int tiles_per_item = (int)(((int64_t)a->cam.n_rays * a->cam.n_samples) / HN_TILE);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    mlp_fwd_kernel<<<grid, kFusedThreads, kFwdSmem, (cudaStream_t)stream>>>(*a, n_tiles, tiles_per_item);
    return check_launch("hn_mlp_fwd");
}
