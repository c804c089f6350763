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
constexpr uint32_t kFwdSmem = 7 * kUnitBytes + kStages * kStageBytes + 1024;
constexpr uint32_t kTmemCols = 512;

struct FwdShared {
    uint64_t w_full[kStages], w_empty[kStages];
    uint64_t a_ready[3], pe_ready, acc_full[4], acc_empty[4];
    float dens[4][128];
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

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(&sh.w_full[i]), 1); mbar_init(smem_u32(&sh.w_empty[i]), 1); }
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&sh.a_ready[i]), kEpiWarps);
        mbar_init(smem_u32(&sh.pe_ready), kEpiWarps);
        for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&sh.acc_full[i]), 1); mbar_init(smem_u32(&sh.acc_empty[i]), kEpiWarps); }
        sh.abort = 0;
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(smem_u32(&sh.tmem_base));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = sh.tmem_base;
    const int n_ops = c_fwd.n_ops;

    if (warp == 0) {
        // ======================= weight producer =======================
        if (lane == 0) {
            uint32_t uc = 0;
            const uint8_t* packed = (const uint8_t*)a.packed;
            for (int tile = blockIdx.x; tile < n_tiles && !sh.abort; tile += gridDim.x) {
                for (int u = 0; u < n_ops; ++u, ++uc) {
                    const uint32_t stage = uc % kStages, par = (uc / kStages) & 1;
                    if (!wait_or_abort(&sh.w_empty[stage], par ^ 1, &sh.abort, a.status, 101)) break;
                    const MmaOp op = c_fwd.mma[u];
                    const uint32_t bytes = (uint32_t)op.n8 * 8 * 128;
                    const uint32_t fb = smem_u32(&sh.w_full[stage]);
                    mbar_arrive_expect_tx(fb, bytes * op.nkb);
                    for (int k = 0; k < op.nkb; ++k)
                        bulk_g2s(smem + kOffW + stage * kStageBytes + k * kUnitBytes, packed + (size_t)(op.unit + k) * kUnitBytes, bytes, fb);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            uint32_t uc = 0, par_ready = 0, par_pe = 0, par_empty = 0;
            for (int tile = blockIdx.x; tile < n_tiles && !sh.abort; tile += gridDim.x) {
                MmaOp op = c_fwd.mma[0];
                for (int u = 0; u < n_ops; ++u, ++uc) {
                    const MmaOp nxt = c_fwd.mma[u + 1 < n_ops ? u + 1 : 0];       // table read off the critical path
                    bool ok = true;
                    if (op.wait_src == 4) { ok = wait_or_abort(&sh.pe_ready, par_pe, &sh.abort, a.status, 201); par_pe ^= 1; }
                    else if (op.wait_src) {
                        const int c = op.wait_src - 1;
                        ok = wait_or_abort(&sh.a_ready[c], (par_ready >> c) & 1, &sh.abort, a.status, 202 + c);
                        par_ready ^= 1u << c;
                    }
                    if (ok && op.wait_empty) {
                        ok = wait_or_abort(&sh.acc_empty[op.q], ((par_empty >> op.q) & 1) ^ 1, &sh.abort, a.status, 210 + op.q);
                        par_empty ^= 1u << op.q;
                    }
                    const uint32_t stage = uc % kStages, par = (uc / kStages) & 1;
                    if (ok) ok = wait_or_abort(&sh.w_full[stage], par, &sh.abort, a.status, 220);
                    if (!ok) break;
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem + (op.a_blk == kPeBlk ? kOffPE : kOffA + op.a_blk * kUnitBytes);
                    const uint32_t b_addr = smem + kOffW + stage * kStageBytes;
                    const uint32_t idesc = umma_idesc(128, (uint32_t)op.n8 * 8, kF16, kF16, 0, 0);
                    const uint32_t d_addr = tmem_base + (uint32_t)op.tmem_col8 * 8;
                    for (int k = 0; k < op.nkb; ++k) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            umma_f16(d_addr, umma_desc_kmajor(a_addr + k * kUnitBytes, ks), umma_desc_kmajor(b_addr + k * kUnitBytes, ks), idesc,
                                     !(op.first && k == 0 && ks == 0));
                    }
                    umma_commit(smem_u32(&sh.w_empty[stage]));
                    if (op.commit) umma_commit(smem_u32(&sh.acc_full[op.q]));
                    op = nxt;
                }
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

        if ((int)blockIdx.x < n_tiles) produce_pe(blockIdx.x);
        for (int tile = blockIdx.x; tile < n_tiles && !sh.abort; tile += gridDim.x) {
            const size_t m = (size_t)tile * HN_TILE + row;
            const int b = tile / tiles_per_item;
            const float* bias_row = a.bias + (size_t)b * HN_BIAS_STRIDE;
            float dens = 0.f;
            for (int e = 0; e < kFwdEpis; ++e) {
                const EpiOp op = c_fwd.epi[e];
                // on a pipeline fault every later wait returns at once; the loop still runs to its end so that all
                // epilogue threads keep meeting at the same named barriers
                wait_or_abort(&sh.acc_full[op.q], (par_full >> op.q) & 1, &sh.abort, a.status, 300 + e);
                par_full ^= 1u << op.q;
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
                    const float4* bp = reinterpret_cast<const float4*>(bias_row + op.bias_off + col);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 bb = __ldg(bp + i);
                        y[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bb.x;
                        y[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bb.y;
                        y[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bb.z;
                        y[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bb.w;
                    }
                }
                tc_fence_before_sync();
                warp_arrive(smem_u32(&sh.acc_empty[op.q]), lane);  // accumulator read: hand it back to the MMA issuer
                if (active) {
                    if (op.kind == EPI_HIDDEN) {
                        if (a.masks && op.mask_word != 0xFFFF)
                            a.masks[m * HN_MASK_WORDS + op.mask_word + cg] = positive_mask32(y);
                        if (op.density) {                          // density head on the fp32 activations (models.py:78,83)
                            const float4* wp = reinterpret_cast<const float4*>(a.w_density + op.col0 + col);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 ww = __ldg(wp + i);
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
                if (op.density == 2) {                             // combine the four column groups' partial dot products
                    sh.dens[cg][row] = dens;
                    dens = 0.f;
                    named_sync(2, kEpiThreads);
                    if (cg == 0)
                        a.sigma[m] = fmaxf(sh.dens[0][row] + sh.dens[1][row] + sh.dens[2][row] + sh.dens[3][row] +
                                           __ldg(bias_row + HN_BIAS_OFF_DENSITY), 0.f);
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
                if (e == pe_after && tile + (int)gridDim.x < n_tiles) produce_pe(tile + gridDim.x);
            }
        }
        if (saving && leader) bulk_wait_all<0>();
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
