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
#include "hn_mlp_sched.h"
#include "hn_sample.cuh"
#include "hn_tc.cuh"

namespace hn {

__constant__ FwdTables c_fwd;

constexpr int kStages = 6;
constexpr int kFwdThreads = 256;
constexpr uint32_t kOffA = 0;                               // 6 activation blocks
constexpr uint32_t kOffPE = 6 * kUnitBytes;                 // 1 PE block
constexpr uint32_t kOffW = 7 * kUnitBytes;                  // weight ring
constexpr uint32_t kFwdSmem = (7 + kStages) * kUnitBytes + 1024;
constexpr uint32_t kTmemCols = 512;

struct FwdShared {
    uint64_t w_full[kStages], w_empty[kStages];
    uint64_t a_ready[3], pe_ready, acc_full[4], acc_empty[4];
    uint32_t tmem_base;
    volatile int abort;
};

__device__ __forceinline__ bool wait_or_abort(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* status, int code) {
    const uint32_t b = smem_u32(bar);
    if (mbar_try_wait(b, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(b, parity)) {
        if (*abort_flag) return false;
        if (clock64() - t0 > 2000000000ll) {             // ~1 s: a protocol bug must not hang the GPU
            *abort_flag = 1;
            atomicCAS(status, 0, code);
            return false;
        }
    }
    return true;
}

__device__ __forceinline__ void named_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// positional encoding of one sample into row `row` of the PE operand block (NetWorks/utils.py:20-51)
__device__ __forceinline__ void write_pe_row(uint32_t pe_block, int row, float px, float py, float pz) {
    float v[64];
    v[0] = px; v[1] = py; v[2] = pz;
    const float p[3] = {px, py, pz};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        const float f = (float)(1 << k);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float s, c;
            sincosf(p[d] * f, &s, &c);
            v[3 + 6 * k + d] = s;
            v[3 + 6 * k + 3 + d] = c;
        }
    }
    v[63] = 0.f;
    const uint32_t row_addr = pe_block + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch)
        st_shared_v4(row_addr + ((ch ^ (row & 7)) << 4),
                     pack_h2(v[8 * ch + 0], v[8 * ch + 1]), pack_h2(v[8 * ch + 2], v[8 * ch + 3]),
                     pack_h2(v[8 * ch + 4], v[8 * ch + 5]), pack_h2(v[8 * ch + 6], v[8 * ch + 7]));
}

__global__ void __launch_bounds__(kFwdThreads, 1) mlp_fwd_kernel(const hn_mlp_fwd_t a, const int n_tiles, const int tiles_per_item) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ FwdShared sh;
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool saving = (a.act != nullptr);

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(&sh.w_full[i]), 1); mbar_init(smem_u32(&sh.w_empty[i]), 1); }
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&sh.a_ready[i]), 128);
        mbar_init(smem_u32(&sh.pe_ready), 128);
        for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&sh.acc_full[i]), 1); mbar_init(smem_u32(&sh.acc_empty[i]), 128); }
        sh.abort = 0;
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<kTmemCols>(smem_u32(&sh.tmem_base));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = sh.tmem_base;

    if (warp == 0) {
        // ======================= weight producer =======================
        if (lane == 0) {
            uint32_t uc = 0;
            const uint8_t* packed = (const uint8_t*)a.packed;
            for (int tile = blockIdx.x; tile < n_tiles && !sh.abort; tile += gridDim.x) {
                for (int u = 0; u < kFwdUnits; ++u, ++uc) {
                    const uint32_t stage = uc % kStages, par = (uc / kStages) & 1;
                    if (!wait_or_abort(&sh.w_empty[stage], par ^ 1, &sh.abort, a.status, 101)) break;
                    const uint32_t bytes = (uint32_t)c_fwd.mma[u].n8 * 8 * 128;
                    mbar_arrive_expect_tx(smem_u32(&sh.w_full[stage]), bytes);
                    bulk_g2s(smem + kOffW + stage * kUnitBytes, packed + (size_t)u * kUnitBytes, bytes, smem_u32(&sh.w_full[stage]));
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            uint32_t uc = 0, par_ready = 0, par_pe = 0, par_empty = 0;
            for (int tile = blockIdx.x; tile < n_tiles && !sh.abort; tile += gridDim.x) {
                for (int u = 0; u < kFwdUnits; ++u, ++uc) {
                    const MmaOp op = c_fwd.mma[u];
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
                    const uint32_t b_addr = smem + kOffW + stage * kUnitBytes;
                    const uint32_t idesc = umma_idesc(128, (uint32_t)op.n8 * 8, kF16, kF16, 0, 0);
                    const uint32_t d_addr = tmem_base + (uint32_t)op.tmem_col8 * 8;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(d_addr, umma_desc_kmajor(a_addr, ks), umma_desc_kmajor(b_addr, ks), idesc, !(op.first && ks == 0));
                    umma_commit(smem_u32(&sh.w_empty[stage]));
                    if (op.commit) umma_commit(smem_u32(&sh.acc_full[op.q]));
                }
            }
        }
    } else if (warp >= 4) {
        // ======================= PE producer + epilogue (128 threads, thread = tile row = TMEM lane) =======================
        const int row = tid - 128;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t row_off = (row >> 3) * 1024 + (row & 7) * 128;
        const int rsw = row & 7;
        uint32_t par_full = 0;
        const bool leader = (row == 0);
        for (int tile = blockIdx.x; tile < n_tiles && !sh.abort; tile += gridDim.x) {
            const size_t m = (size_t)tile * HN_TILE + row;
            const int b = tile / tiles_per_item;
            // ---- sampling + positional encoding: the first GEMM's operand is generated, not loaded
            {
                const int ns = a.cam.n_samples;
                const size_t ray_idx = m / ns;
                const int s = (int)(m % ns), r = (int)(ray_idx % a.cam.n_rays);
                const Ray ray = make_ray(a.cam, b, r);
                const Sample q = make_sample(a.cam, ray, b, r, s);
                a.delta[m] = q.zdist;
                if (a.zvals) a.zvals[m] = q.zval;
                if (saving) { if (leader) bulk_wait_read<1>(); named_sync(1, 128); }
                write_pe_row(smem + kOffPE, row, q.px, q.py, q.pz);
                fence_async_smem();
                if (saving) {
                    named_sync(1, 128);
                    if (leader) {
                        bulk_s2g((uint8_t*)a.act + ((size_t)HN_SLOT_PE * n_tiles + tile) * kUnitBytes, smem + kOffPE, kUnitBytes);
                        bulk_commit();
                    }
                }
                mbar_arrive(smem_u32(&sh.pe_ready));
            }
            const float* bias_row = a.bias + (size_t)b * HN_BIAS_STRIDE;
            float dens = 0.f;
            for (int e = 0; e < kFwdEpis; ++e) {
                const EpiOp op = c_fwd.epi[e];
                // on a pipeline fault every later wait returns at once; the loop still runs to its end so that all
                // 128 threads keep meeting at the same named barriers
                wait_or_abort(&sh.acc_full[op.q], (par_full >> op.q) & 1, &sh.abort, a.status, 300 + e);
                par_full ^= 1u << op.q;
                tc_fence_after_sync();
                const bool to_smem = (op.kind != EPI_FEAT);
                if (to_smem && saving) { if (leader) bulk_wait_read<1>(); named_sync(1, 128); }
                uint32_t mask_words[4] = {0, 0, 0, 0};
                for (int g = 0; g < op.width32; ++g) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_base + (uint32_t)op.tmem_col8 * 8 + g * 32, v);
                    tmem_ld_wait();
                    if (g + 1 == op.width32) {            // accumulator fully read: hand it back to the MMA issuer
                        tc_fence_before_sync();
                        mbar_arrive(smem_u32(&sh.acc_empty[op.q]));
                    }
                    const float4* bp = reinterpret_cast<const float4*>(bias_row + op.bias_off + g * 32);
                    float y[32];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 bb = __ldg(bp + i);
                        y[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bb.x;
                        y[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bb.y;
                        y[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bb.z;
                        y[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bb.w;
                    }
                    if (op.kind == EPI_HIDDEN) {
                        uint32_t mw = 0;
#pragma unroll
                        for (int i = 0; i < 32; ++i) { mw |= (y[i] > 0.f ? 1u : 0u) << i; y[i] = fminf(fmaxf(y[i], 0.f), 65504.f); }
                        mask_words[g] = mw;
                    } else if (op.kind == EPI_LINEAR) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) y[i] = fminf(fmaxf(y[i], -65504.f), 65504.f);
                    }
                    if (op.density) {                     // density head on the fp32 activations (models.py:78,83)
                        const float4* wp = reinterpret_cast<const float4*>(a.w_density + op.col0 + g * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 ww = __ldg(wp + i);
                            dens = fmaf(y[4 * i + 0], ww.x, dens); dens = fmaf(y[4 * i + 1], ww.y, dens);
                            dens = fmaf(y[4 * i + 2], ww.z, dens); dens = fmaf(y[4 * i + 3], ww.w, dens);
                        }
                    }
                    if (to_smem) {
                        const int col = g * 32;                                  // column inside the 128-wide chunk
                        const uint32_t blk_addr = smem + kOffA + (op.dst_blk + (col >> 6)) * kUnitBytes + row_off;
                        const int ch0 = (col & 63) >> 3;
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            st_shared_v4(blk_addr + (((ch0 + c) ^ rsw) << 4),
                                         pack_h2(y[8 * c + 0], y[8 * c + 1]), pack_h2(y[8 * c + 2], y[8 * c + 3]),
                                         pack_h2(y[8 * c + 4], y[8 * c + 5]), pack_h2(y[8 * c + 6], y[8 * c + 7]));
                    } else if (a.feat) {
                        float4* dst = reinterpret_cast<float4*>(a.feat + m * HN_FEAT + op.col0 + g * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
                    }
                }
                if (op.density == 2) {
                    a.sigma[m] = fmaxf(dens + __ldg(bias_row + HN_BIAS_OFF_DENSITY), 0.f);
                    dens = 0.f;
                }
                if (a.masks && op.mask_word != 0xFFFF) {
                    uint32_t* mp = a.masks + m * HN_MASK_WORDS + op.mask_word;
                    if (op.width32 == 4) *reinterpret_cast<uint4*>(mp) = make_uint4(mask_words[0], mask_words[1], mask_words[2], mask_words[3]);
                    else *reinterpret_cast<uint2*>(mp) = make_uint2(mask_words[0], mask_words[1]);
                }
                if (to_smem) {
                    fence_async_smem();
                    if (saving) {
                        named_sync(1, 128);
                        if (leader && op.save_blk != 0xFFFF) {
                            const int nblk = (op.width32 + 1) / 2;
                            for (int k = 0; k < nblk; ++k)
                                bulk_s2g((uint8_t*)a.act + ((size_t)(op.save_blk + k) * n_tiles + tile) * kUnitBytes,
                                         smem + kOffA + (op.dst_blk + k) * kUnitBytes, kUnitBytes);
                            bulk_commit();
                        }
                    }
                    mbar_arrive(smem_u32(&sh.a_ready[op.ready_idx]));
                }
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
    mlp_fwd_kernel<<<grid, kFwdThreads, kFwdSmem, (cudaStream_t)stream>>>(*a, n_tiles, tiles_per_item);
    return check_launch("hn_mlp_fwd");
}
