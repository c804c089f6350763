// hn_mlp_bwd.cu — data-gradient chain of fg_CD_predictor for sm_100a (autograd of NetWorks/models.py:62-87,
// NetWorks/utils.py:43-51,80-86), the mirror image of hn_mlp_fwd.cu.
//
// Per 128-sample tile, one persistent CTA:
//   producer  : bulk-loads the tile's dL/dfeat operand image (written by hn_composite_bwd) into the gradient
//               buffer, then streams the transposed weight units (W^T) through the smem ring
//   MMA issuer: dX = dZ * W for RGB_layer_2, _1, _0, FeaExt_module_7..1 (tcgen05, fp16 operands holding
//               loss-scaled gradients, fp32 accumulate); dL/dPE accumulates in its own 64 TMEM columns
//               from FeaExt_module_5 and _0
//   epilogue  : TMEM -> registers -> (+ dsigma x w_density for the density head) -> ReLU mask from the
//               forward's bit masks -> fp16 -> gradient buffer (next GEMM's A operand) and, for the weight
//               pass, bulk-stored to HBM as operand images; finally dL/dPE -> dL/dpts (PE backward) ->
//               per-ray reductions (origin, direction*length, length) for the camera gradients
// Gradients are carried multiplied by the power-of-two *grad_scale so that fp16 keeps them in range.
#include <cstdlib>
#include <mutex>
#include "hn_api.h"
#include "hn_mlp_common.cuh"
#include "hn_mlp_sched.h"
#include "hn_sample.cuh"

namespace hn {

__constant__ BwdTables c_bwd[2];          // [0] without dL/dPE, [1] with

constexpr int kBStages = 3;                                 // weight ring: 3 stages of two 64-wide K blocks (32 KiB)
constexpr uint32_t kBStageBytes = 2 * kUnitBytes;
constexpr uint32_t kBOffZ = 0;
constexpr uint32_t kBOffW = 6 * kUnitBytes;
constexpr uint32_t kBwdSmem = 6 * kUnitBytes + kBStages * kBStageBytes + 1024;
constexpr uint32_t kBTmemCols = 512;

struct BwdShared {
    uint64_t w_full[3], w_empty[3];
    uint64_t a_ready[3], in_ready, z_free, acc_full[4], acc_empty[4];
    float dp_part[2][128][3];
    uint32_t tmem_base;
    volatile int abort;
};

__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// PE backward for PE columns [32*H, 32*H+32): partial dL/dpts of one sample.
// g[i] = dL/dPE[32*H + i] (still loss-scaled); p = sample position.
template <int H>
__device__ __forceinline__ void pe_backward_half(const uint32_t (&g)[32], const float (&p)[3], float (&dp)[3]) {
    dp[0] = dp[1] = dp[2] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int c = 32 * H + i;
        const float gi = __uint_as_float(g[i]);
        if (c < 3) dp[c] += gi;
        else if (c < 63) {
            const int k = (c - 3) / 6, t = (c - 3) % 6, d = t % 3;
            const float f = (float)(1 << k);
            const float arg = p[d] * f;
            dp[d] = fmaf(f * gi, (t < 3) ? cosf(arg) : -sinf(arg), dp[d]);
        }
    }
}

__global__ void __launch_bounds__(kFusedThreads, 1) mlp_bwd_kernel(const hn_mlp_bwd_data_t a, const int n_tiles, const int with_pe) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ BwdShared sh;
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const BwdTables& tb = c_bwd[with_pe];
    const bool saving = (a.grads != nullptr);
    // weight ring geometry: single CTA 3 x 32 KiB (two 64-wide K blocks of 128 rows); pair mode 6 x 16 KiB (this CTA's
    // half of the rows of both K blocks) - same bytes in flight per CTA, i.e. twice the prefetch depth per weight byte needed
    constexpr int STAGES = 3;
    constexpr uint32_t STAGE_BYTES = 2 * kUnitBytes;
    constexpr uint32_t KB_STRIDE = kUnitBytes;
    const int work0 = (int)blockIdx.x;
    const int work_stride = (int)gridDim.x;
    const int n_work = n_tiles;

    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(smem_u32(&sh.w_full[i]), 1); mbar_init(smem_u32(&sh.w_empty[i]), 1);
        }
        for (int i = 0; i < 3; ++i) { mbar_init(smem_u32(&sh.a_ready[i]), kEpiWarps); }
        mbar_init(smem_u32(&sh.in_ready), 1);
        mbar_init(smem_u32(&sh.z_free), kEpiWarps);
        for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&sh.acc_full[i]), 1); mbar_init(smem_u32(&sh.acc_empty[i]), kEpiWarps); }
        sh.abort = 0;
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<kBTmemCols>(smem_u32(&sh.tmem_base));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = sh.tmem_base;
    const int n_ops = tb.n_ops, n_epis = tb.n_epis;

    if (warp == 0) {
        // ======================= producer: dL/dfeat image + W^T units =======================
        if (lane == 0) {
            uint32_t uc = 0, par_free = 0;
            const uint8_t* wt = (const uint8_t*)a.packed + (size_t)kFwdUnits * kUnitBytes;
            const uint8_t* din = (const uint8_t*)a.dfeat_image;
            for (int w = work0; w < n_work && !sh.abort; w += work_stride) {
                const int tile = w;
                if (!wait_or_abort(&sh.z_free, par_free ^ 1, &sh.abort, a.status, 401)) break;
                par_free ^= 1;
                mbar_arrive_expect_tx(smem_u32(&sh.in_ready), 4 * kUnitBytes);
                for (int kb = 0; kb < 4; ++kb)
                    bulk_g2s(smem + kBOffZ + kb * kUnitBytes, din + ((size_t)kb * n_tiles + tile) * kUnitBytes, kUnitBytes, smem_u32(&sh.in_ready));
                for (int u = 0; u < n_ops; ++u, ++uc) {
                    const uint32_t stage = uc % STAGES, par = (uc / STAGES) & 1;
                    if (!wait_or_abort(&sh.w_empty[stage], par ^ 1, &sh.abort, a.status, 402)) break;
                    const MmaOp op = tb.mma[u];
                    const uint32_t bytes = (uint32_t)op.n8 * 8 * 128;
                    const uint32_t fb = smem_u32(&sh.w_full[stage]);
                    mbar_arrive_expect_tx(fb, bytes * op.nkb);
                    for (int k = 0; k < op.nkb; ++k)
                        bulk_g2s(smem + kBOffW + stage * STAGE_BYTES + k * KB_STRIDE,
                                 wt + (size_t)(op.unit + k) * kUnitBytes, bytes, fb);
                }
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        {
            // the whole warp walks the schedule (uniform control flow keeps descriptors in uniform registers); one
            // elected lane issues the MMAs and commits
            uint32_t uc = 0, par_ready = 0, par_in = 0, par_empty = 0;
            // The issuer shares its scheduler with four epilogue warps and gets a fraction of the issue slots, so what it
            // executes per MMA bounds the tensor pipe: dependencies are polled without clock reads, the NEXT unit's weight
            // barrier is queried before its answer is needed, and one elected-lane block issues the MMAs and the commits.
            bool pre = mbar_try_wait(smem_u32(&sh.w_full[0]), 0);
            MmaOp op = tb.mma[0];
            for (int w = work0; w < n_work && !sh.abort; w += work_stride) {
                for (int u = 0; u < n_ops; ++u, ++uc) {
                    const MmaOp nxt = tb.mma[u + 1 < n_ops ? u + 1 : 0];          // table read off the critical path
                    bool ok = true;
                    if (op.wait_src == 5) {
                        ok = wait_spin(&sh.in_ready, par_in, &sh.abort, a.status, 501);
                        par_in ^= 1;
                    } else if (op.wait_src) {
                        const int c = op.wait_src - 1;
                        ok = wait_spin(&sh.a_ready[c], (par_ready >> c) & 1, &sh.abort, a.status, 502 + c);
                        par_ready ^= 1u << c;
                    }
                    if (ok && op.wait_empty) {
                        ok = wait_spin(&sh.acc_empty[op.q], ((par_empty >> op.q) & 1) ^ 1, &sh.abort, a.status, 510 + op.q);
                        par_empty ^= 1u << op.q;
                    }
                    const uint32_t stage = uc % STAGES, par = (uc / STAGES) & 1;
                    if (ok && !pre) ok = wait_spin(&sh.w_full[stage], par, &sh.abort, a.status, 520);
                    if (!ok) break;
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem + kBOffZ + op.a_blk * kUnitBytes;
                    const uint32_t b_addr = smem + kBOffW + stage * STAGE_BYTES;
                    const uint32_t idesc = umma_idesc(128, (uint32_t)op.n8 * 8, kF16, kF16, 0, 0);
                    const uint32_t d_addr = tmem_base + (uint32_t)op.tmem_col8 * 8;
                    const uint32_t a_lo = desc_lo(a_addr, 16), b_lo = desc_lo(b_addr, 16);
                    const uint32_t first = op.first, nkb = op.nkb;
                    const uint32_t empty_bar = smem_u32(&sh.w_empty[stage]), full_bar = smem_u32(&sh.acc_full[op.q]);
                    if (elect_one()) {
                        for (uint32_t k = 0; k < nkb; ++k) {
#pragma unroll
                            for (uint32_t ks = 0; ks < 4; ++ks)
                                umma_f16_lohi(d_addr, a_lo + k * (kUnitBytes >> 4) + ks * 2, b_lo + k * (KB_STRIDE >> 4) + ks * 2, idesc,
                                                  (first && k == 0 && ks == 0) ? 0u : 1u);
                        }
                        umma_commit(empty_bar);
                        if (op.commit) umma_commit(full_bar);
                    }
                    __syncwarp();
                    // ask for the next unit's weights now; the answer is consumed at the top of the next iteration
                    pre = mbar_try_wait(smem_u32(&sh.w_full[(uc + 1) % STAGES]), ((uc + 1) / STAGES) & 1);
                    op = nxt;
                }
            }
        }
    } else if (warp >= kCtrlWarps) {
        // ======================= epilogue: 16 warps, 32 rows x 32 columns each =======================
        const int ew = warp - kCtrlWarps;
        const int cg = ew >> 2;
        const int row = (ew & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((ew & 3) * 32) << 16;
        const uint32_t row_off = (row >> 3) * 1024 + (row & 7) * 128;
        const int rsw = row & 7;
        uint32_t par_full = 0;
        const bool leader = (ew == 0 && lane == 0);
        const float scale = __ldg(a.grad_scale);
        const float inv_scale = 1.0f / scale;
        for (int w = work0; w < n_work && !sh.abort; w += work_stride) {
            const int tile = w;
            const size_t m = (size_t)tile * HN_TILE + row;
            const float dsr = (__ldg(a.sigma + m) > 0.f) ? __ldg(a.dsigma + m) * scale : 0.f;   // d/d(pre-ReLU density), scaled
            const uint32_t* mask_row = a.masks + m * HN_MASK_WORDS;
            for (int e = 0; e < n_epis; ++e) {
                const EpiOp op = tb.epi[e];
                wait_or_abort(&sh.acc_full[op.q], (par_full >> op.q) & 1, &sh.abort, a.status, 600 + e);
                par_full ^= 1u << op.q;
                tc_fence_after_sync();
                if (op.kind == EPI_GRAD_PE) {
                    // ---- dL/dPE (64 columns, scaled) -> dL/dpts -> per-ray sums
                    const int ns = a.cam.n_samples;
                    const size_t ray_idx = m / ns;
                    const int s = (int)(m % ns), r = (int)(ray_idx % a.cam.n_rays), b = (int)(ray_idx / a.cam.n_rays);
                    const Ray ray = make_ray(a.cam, b, r);
                    const float e0 = sample_edge(a.cam, ray.oz, b, r, s), e1 = sample_edge(a.cam, ray.oz, b, r, s + 1);
                    const float p[3] = {__fadd_rn(ray.ox, __fmul_rn(ray.vx, e0)), __fadd_rn(ray.oy, __fmul_rn(ray.vy, e0)),
                                        __fadd_rn(ray.oz, __fmul_rn(ray.vz, e0))};
                    if (cg < 2) {
                        uint32_t v[32];
                        tmem_ld32(tmem_base + lane_base + (uint32_t)op.tmem_col8 * 8 + cg * 32, v);
                        tmem_ld_wait();
                        float part[3];
                        if (cg == 0) pe_backward_half<0>(v, p, part); else pe_backward_half<1>(v, p, part);
                        sh.dp_part[cg][row][0] = part[0]; sh.dp_part[cg][row][1] = part[1]; sh.dp_part[cg][row][2] = part[2];
                    }
                    tc_fence_before_sync();
                    warp_arrive(smem_u32(&sh.acc_empty[op.q]), lane);
                    named_sync(2, kEpiThreads);
                    if (cg == 0) {
                        float dp[3];
#pragma unroll
                        for (int d = 0; d < 3; ++d) dp[d] = (sh.dp_part[0][row][d] + sh.dp_part[1][row][d]) * inv_scale;
                        const float dpv = dp[0] * ray.vx + dp[1] * ray.vy + dp[2] * ray.vz;      // through z_s = o_z - const
                        float red[7] = {dp[0], dp[1], dp[2] + dpv, e0 * dp[0], e0 * dp[1], e0 * dp[2],
                                        a.ddelta ? (e1 - e0) * __ldg(a.ddelta + m) : 0.f};
#pragma unroll
                        for (int i = 0; i < 7; ++i) red[i] = warp_sum32(red[i]);
                        if (lane == 0 && a.g_ray_o) {
                            atomicAdd(a.g_ray_o + ray_idx * 3 + 0, red[0]); atomicAdd(a.g_ray_o + ray_idx * 3 + 1, red[1]);
                            atomicAdd(a.g_ray_o + ray_idx * 3 + 2, red[2]);
                            atomicAdd(a.g_ray_v + ray_idx * 3 + 0, red[3]); atomicAdd(a.g_ray_v + ray_idx * 3 + 1, red[4]);
                            atomicAdd(a.g_ray_v + ray_idx * 3 + 2, red[5]);
                            atomicAdd(a.g_ray_l + ray_idx, red[6]);
                        }
                    }
                    continue;
                }
                const bool active = cg < op.width32;
                const int col = cg * 32;
                if (saving) { if (leader) bulk_wait_read<1>(); named_sync(1, kEpiThreads); }
                if (saving && op.kind == EPI_GRAD_DENSITY && op.col0 == 0 && cg == 0) {
                    // density head as a one-channel pseudo layer for the weight pass: row = [dsr, 0, ..., 0]
                    uint8_t* drow = (uint8_t*)a.grads + ((size_t)HN_GSLOT_DENS * n_tiles + tile) * kUnitBytes + row_off;
                    const uint32_t first = pack_sat(dsr, 0.f);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(drow + ((c ^ rsw) << 4)) = make_uint4(c == 0 ? first : 0u, 0u, 0u, 0u);
                }
                float y[32];
                if (active) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + lane_base + (uint32_t)op.tmem_col8 * 8 + col, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(v[i]);
                }
                tc_fence_before_sync();
                warp_arrive(smem_u32(&sh.acc_empty[op.q]), lane);
                if (active) {
                    if (op.kind == EPI_GRAD_DENSITY) {
                        const float4* wp = reinterpret_cast<const float4*>(a.w_density + op.col0 + col);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 ww = __ldg(wp + i);
                            y[4 * i + 0] = fmaf(dsr, ww.x, y[4 * i + 0]); y[4 * i + 1] = fmaf(dsr, ww.y, y[4 * i + 1]);
                            y[4 * i + 2] = fmaf(dsr, ww.z, y[4 * i + 2]); y[4 * i + 3] = fmaf(dsr, ww.w, y[4 * i + 3]);
                        }
                    }
                    if (op.kind != EPI_GRAD_LINEAR) {
                        const uint32_t mw = __ldg(mask_row + op.mask_word + cg);
#pragma unroll
                        for (int i = 0; i < 32; ++i) y[i] = (mw & (1u << i)) ? y[i] : 0.f;
                    }
                    store_row32<false>(smem + kBOffZ + op.dst_blk * kUnitBytes, row, col, y);
                }
                fence_async_smem();
                if (saving) {
                    named_sync(1, kEpiThreads);
                    if (leader && op.save_blk != 0xFFFF) {
                        const int nblk = (op.width32 + 1) / 2;
                        for (int k = 0; k < nblk; ++k)
                            bulk_s2g((uint8_t*)a.grads + ((size_t)(op.save_blk + k) * n_tiles + tile) * kUnitBytes,
                                     smem + kBOffZ + (op.dst_blk + k) * kUnitBytes, kUnitBytes);
                        bulk_commit();
                    }
                }
                if (op.ready_idx != 255) warp_arrive(smem_u32(&sh.a_ready[op.ready_idx]), lane);
            }
            // the tile's gradient buffer may now be overwritten by the next tile's dL/dfeat image
            if (saving && leader) bulk_wait_read<0>();
            warp_arrive(smem_u32(&sh.z_free), lane);
        }
        if (saving && leader) bulk_wait_all<0>();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 2) tmem_free<kBTmemCols>(tmem_base);
}

static std::mutex g_bwd_mu;
static bool g_bwd_ready[64] = {};

int launch_bwd_data_tmem(const hn_mlp_bwd_data_t* a, void* stream);   // hn_mlp_fwd.cu: the chain with gradients resident in tensor memory

}  // namespace hn

extern "C" int hn_mlp_bwd_data(const hn_mlp_bwd_data_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->cam.xy || !a->cam.Rmats || !a->cam.Tvecs || !a->cam.inv_inmats || !a->packed || !a->w_density ||
        !a->dfeat_image || !a->dsigma || !a->sigma || !a->grad_scale || !a->masks || !a->status)
        return set_error(HN_E_BADARG, "hn_mlp_bwd_data: null pointer");
    if (int rc = check_geometry(a->cam.B, a->cam.n_rays, a->cam.n_samples, "hn_mlp_bwd_data")) return rc;
    const bool with_pe = (a->g_ray_o != nullptr);
    if (with_pe && (!a->g_ray_v || !a->g_ray_l))
        return set_error(HN_E_BADARG, "hn_mlp_bwd_data: g_ray_o, g_ray_v and g_ray_l must be given together");
    const int64_t M_all = total_samples(a->cam.B, a->cam.n_rays, a->cam.n_samples);
    // the training step (no camera gradients) runs on the tensor-memory chain (hn_mlp_fwd.cu); dL/dPE needs this file's kernel,
    // which keeps a dedicated accumulator for it
    if (!with_pe) return launch_bwd_data_tmem(a, stream);
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_bwd_mu);
        if (dev < 64 && !g_bwd_ready[dev]) {
            const HostSchedules& hs = host_schedules();
            cudaError_t e = cudaMemcpyToSymbol(c_bwd, &hs.bwd_nope, sizeof(BwdTables), 0);
            if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_bwd, &hs.bwd, sizeof(BwdTables), sizeof(BwdTables));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem);
            if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
            g_bwd_ready[dev] = true;
        }
    }
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int64_t M = total_samples(a->cam.B, a->cam.n_rays, a->cam.n_samples);
    const int n_tiles = (int)(M / HN_TILE);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    mlp_bwd_kernel<<<grid, kFusedThreads, kBwdSmem, (cudaStream_t)stream>>>(*a, n_tiles, with_pe ? 1 : 0);
    return check_launch("hn_mlp_bwd_data");
}
