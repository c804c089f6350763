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
constexpr uint32_t kFwdSmem = kOffBias + kBiasBytes + 1024;
constexpr uint32_t kTmemCols = 512;

struct FwdShared {
    uint64_t w_full[kStages], w_empty[kStages];
    uint64_t a_ready[3], pe_ready, pe_free, acc_full[2], acc_empty[2];
    uint64_t stg_full[2], stg_free[2];
    alignas(16) float dens[128];        // density head: the four column groups add their partial dot products here
    alignas(16) float w_density[HN_HIDDEN];   // read as float4
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

__global__ void __launch_bounds__(kFusedThreads, 1) mlp_fwd_kernel(const hn_mlp_fwd_t a, const int n_tiles, const int tiles_per_item) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ FwdShared sh;
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool saving = (a.act != nullptr);
    const int work0 = (int)blockIdx.x, work_stride = (int)gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(&sh.w_full[i]), 1); mbar_init(smem_u32(&sh.w_empty[i]), 1); }
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&sh.a_ready[i]), kEpiWarps);
        mbar_init(smem_u32(&sh.pe_ready), kEpiWarps); mbar_init(smem_u32(&sh.pe_free), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&sh.acc_full[i]), 1); mbar_init(smem_u32(&sh.acc_empty[i]), kEpiWarps);
            mbar_init(smem_u32(&sh.stg_full[i]), kEpiWarps); mbar_init(smem_u32(&sh.stg_free[i]), 1);
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
            uint32_t save_n = 0, par_pe = 0;
            int pending = -1;                                        // staging buffer whose bulk read is still in flight
            for (int w = work0; w < n_tiles && !sh.abort; w += work_stride) {
                const int tile = w;
                if (!wait_or_abort(&sh.pe_ready, par_pe, &sh.abort, a.status, 150)) break;
                par_pe ^= 1;
                bulk_s2g((uint8_t*)a.act + ((size_t)HN_SLOT_PE * n_tiles + tile) * kUnitBytes, smem + kOffPE, kUnitBytes);
                bulk_commit();
                bulk_wait_read<0>();
                if (pending >= 0) { mbar_arrive(smem_u32(&sh.stg_free[pending])); pending = -1; }
                mbar_arrive(smem_u32(&sh.pe_free));
                for (int e = 0; e < kFwdEpis; ++e) {
                    const EpiOp2 op = c_fwd.epi[e];
                    if (op.save_blk == 0xFFFF) continue;
                    const uint32_t sb = save_n & 1;
                    if (!wait_or_abort(&sh.stg_full[sb], (save_n >> 1) & 1, &sh.abort, a.status, 151)) break;
                    const int nblk = (op.width32 + 1) / 2;
                    for (int k = 0; k < nblk; ++k)
                        bulk_s2g((uint8_t*)a.act + ((size_t)(op.save_blk + k) * n_tiles + tile) * kUnitBytes,
                                 smem + kOffStg + (sb * 2 + k) * kUnitBytes, kUnitBytes);
                    bulk_commit();
                    if (pending >= 0) { bulk_wait_read<1>(); mbar_arrive(smem_u32(&sh.stg_free[pending])); }
                    pending = (int)sb;
                    ++save_n;
                }
            }
            bulk_wait_all<0>();
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        // the whole warp walks the schedule (uniform control flow keeps descriptors in uniform registers); one
        // elected lane issues the MMAs and commits
        uint32_t uc = 0, par_ready = 0, par_pe = 0, chunk_n = 0;
        for (int w = work0; w < n_tiles && !sh.abort; w += work_stride) {
            MmaOp2 op = c_fwd.mma[0];
            for (int u = 0; u < n_ops; ++u, ++uc) {
                const MmaOp2 nxt = c_fwd.mma[u + 1 < n_ops ? u + 1 : 0];         // table read off the critical path
                bool ok = true;
                if (op.wait_src == 4) { ok = wait_or_abort(&sh.pe_ready, par_pe, &sh.abort, a.status, 201); par_pe ^= 1; }
                else if (op.wait_src) {
                    const int c = op.wait_src - 1;
                    ok = wait_or_abort(&sh.a_ready[c], (par_ready >> c) & 1, &sh.abort, a.status, 202 + c);
                    par_ready ^= 1u << c;
                }
                // accumulator of chunk n: released by the epilogue of chunk n-2 (the first two chunks find fresh barriers)
                if (ok && op.first) ok = wait_or_abort(&sh.acc_empty[chunk_n & 1], ((chunk_n >> 1) & 1) ^ 1, &sh.abort, a.status, 210);
                const uint32_t stage = uc % kStages, par = (uc / kStages) & 1;
                if (ok) ok = wait_or_abort(&sh.w_full[stage], par, &sh.abort, a.status, 220);
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
                if (elect_one()) {
                    umma_commit(smem_u32(&sh.w_empty[stage]));
                    if (op.commit) umma_commit(smem_u32(&sh.acc_full[chunk_n & 1]));
                }
                __syncwarp();
                chunk_n += op.commit;
                op = nxt;
            }
        }
    } else {
        // ======================= PE producer + epilogue: 16 warps, 32 rows x 32 columns each =======================
        const int ew = warp - kCtrlWarps;
        const int cg = ew >> 2, quarter = ew & 3;                   // column group inside a chunk; TMEM lane quarter
        const int row = quarter * 32 + lane;                        // tile row = TMEM lane
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t chunk_n = 0, save_n = 0, pe_n = 0;
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
            // the previous tile's PE block must have been read by the saver's bulk store (its MMAs are long done)
            if (saving) wait_or_abort(&sh.pe_free, (pe_n & 1) ^ 1, &sh.abort, a.status, 160);
            ++pe_n;
            const float p[3] = {q.px, q.py, q.pz};
            switch (cg) {
                case 0: write_pe_part<0>(smem + kOffPE, row, p); break;
                case 1: write_pe_part<1>(smem + kOffPE, row, p); break;
                case 2: write_pe_part<2>(smem + kOffPE, row, p); break;
                default: write_pe_part<3>(smem + kOffPE, row, p); break;
            }
            fence_async_smem();
            warp_arrive(smem_u32(&sh.pe_ready), lane);
        };

        if (work0 < n_tiles) produce_pe(work0);
        for (int w = work0; w < n_tiles && !sh.abort; w += work_stride) {
            const int tile = w;
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
            for (int e = 0; e < kFwdEpis; ++e, ++chunk_n) {
                const EpiOp2 op = c_fwd.epi[e];
                // on a pipeline fault every later wait returns at once; the loop still runs to its end so that all
                // epilogue threads keep meeting at the same named barriers
                wait_or_abort(&sh.acc_full[chunk_n & 1], (chunk_n >> 1) & 1, &sh.abort, a.status, 300 + e);
                tc_fence_after_sync();
                const bool active = cg < op.width32;
                const int col = cg * 32;                           // column inside the chunk
                float y[32];
                if (active) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_base + op.acc_col + col, v);
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
                // every warp of this lane quarter has read its part of the accumulator: packed outputs may now be stored over
                // it (in-place slots), and the MMA issuer may reuse it two chunks later
                tc_fence_before_sync();
                named_sync(4 + quarter, 128);
                tc_fence_after_sync();
                warp_arrive(smem_u32(&sh.acc_empty[chunk_n & 1]), lane);
                const bool save = saving && op.save_blk != 0xFFFF;
                const uint32_t sb = save_n & 1;
                if (save) wait_or_abort(&sh.stg_free[sb], ((save_n >> 1) & 1) ^ 1, &sh.abort, a.status, 340);
                if (active) {
                    if (op.kind == EPI_FEAT) {
                        if (a.feat) {
                            float4* dst = reinterpret_cast<float4*>(a.feat + m * HN_FEAT + op.col0 + col);
#pragma unroll
                            for (int i = 0; i < 8; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                        }
                    } else {
                        uint32_t pk[16];
                        if (op.kind == EPI_HIDDEN) {
                            if (a.masks && op.mask_word != 0xFFFF)
                                a.masks[m * HN_MASK_WORDS + op.mask_word + cg] = positive_mask32(y);
                            if (op.density) {                      // density head on the fp32 activations (models.py:78,83)
                                const float4* wp = reinterpret_cast<const float4*>(sh.w_density + op.col0 + col);
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
                        tmem_st16(tmem_base + lane_base + op.out_col + cg * 16, pk);
                        if (save) store_row_packed(smem + kOffStg + sb * 2 * kUnitBytes, row, col, pk);
                        tmem_st_wait();
                    }
                }
                if (op.density == 2) {                             // combine the four column groups' partial dot products
                    atomicAdd(&sh.dens[row], dens);
                    dens = 0.f;
                    named_sync(2, kEpiThreads);
                    if (cg == 0) {
                        a.sigma[m] = fmaxf(sh.dens[row] + __ldg(a.bias + (size_t)b * HN_BIAS_STRIDE + HN_BIAS_OFF_DENSITY), 0.f);
                        sh.dens[row] = 0.f;                        // next use is a whole tile (and many barrier hops) away
                    }
                }
                if (save) { fence_async_smem(); warp_arrive(smem_u32(&sh.stg_full[sb]), lane); ++save_n; }
                if (op.ready_idx != 255) {
                    tc_fence_before_sync();
                    warp_arrive(smem_u32(&sh.a_ready[op.ready_idx]), lane);
                }
                // FeaExt_module_5 (the last reader of the PE block) is done: build the NEXT tile's PE operand now, while
                // the tensor core still has this tile's remaining layers queued
                if (e == pe_after && w + work_stride < n_tiles) produce_pe(w + work_stride);
            }
        }
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
