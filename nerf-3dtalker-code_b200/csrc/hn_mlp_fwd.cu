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
//
// The same kernel, instantiated with BWD = true, runs the DATA-GRADIENT chain of the training step (no dL/dPE):
// dX = dZ * W for RGB_layer_2, _1, _0, FeaExt_module_7..1 with the gradients resident in tensor memory.  Differences: the
// first GEMM's four K blocks (the dL/dfeat operand image written by hn_composite_bwd) stream from HBM through a two-block
// shared-memory ring (the PE warp becomes their loader); the epilogue applies the forward's saved 1-bit ReLU masks and the
// density head's rank-1 term instead of bias + ReLU; every chunk is saved as an operand image for the weight-gradient pass.
// (With camera gradients the older shared-memory-operand kernel of hn_mlp_bwd.cu runs instead: it also accumulates dL/dPE.)
#include <mutex>
#include "hn_api.h"
#include "hn_mlp_common.cuh"
#include "hn_mlp_sched.h"
#include "hn_sample.cuh"

namespace hn {

// HN_TRACE build (tools/trace_fwd.py): CTA 0 records the SM clock of pipeline events of its third tile into status[64..]
#ifdef HN_TRACE
#define HN_TR(cond, slot) do { if ((cond) && blockIdx.x == 0 && tile_k == 2) a.status[64 + (slot)] = (int)clock64(); } while (0)
#else
#define HN_TR(cond, slot) do {} while (0)
#endif

__constant__ __align__(16) FwdTables c_fwd;
__constant__ __align__(16) FwdTables c_bwdt;

// arguments of either chain (filled from hn_mlp_fwd_t / hn_mlp_bwd_data_t by the entry points)
struct ChainArgs {
    hn_camera_t cam;             // forward
    const float* bias;           // forward
    const float* w_density;
    const void* packed;          // weight units of this chain in stage order
    float* feat;                 // forward out
    float* sigma;                // forward: out; data gradients: the saved forward output (ReLU mask of the density head)
    float* delta;                // forward out
    float* zvals;                // forward out or NULL
    void* act;                   // saved operand images: activations (forward) / pre-activation gradients (data gradients), or NULL
    uint32_t* masks;             // forward: out; data gradients: in
    const void* dfeat_image;     // data gradients
    const float* dsigma;         // data gradients
    const float* grad_scale;     // data gradients
    int* status;
};

constexpr int kStages = 4;                                  // weight ring: two consecutive 16 KiB units (32 KiB) per stage
constexpr uint32_t kStageBytes = 2 * kUnitBytes;
constexpr uint32_t kOffPE = 0;                              // PE operand block (forward) / two-block dL/dfeat ring (data gradients)
constexpr uint32_t kBiasBytes = HN_BIAS_STRIDE * 4;
template <bool BWD> struct Lay {
    static constexpr uint32_t kOffW = (BWD ? 2 : 1) * kUnitBytes;           // weight ring
    static constexpr uint32_t kOffStg = kOffW + kStages * kStageBytes;      // 2 staging buffers of two blocks (saved chunks)
    static constexpr uint32_t kOffBias = kOffStg + 4 * kUnitBytes;          // this item's effective bias row (forward only)
    static constexpr uint32_t kOffShared = kOffBias + (BWD ? 0 : kBiasBytes);   // barriers and small arrays close the dynamic region
};
constexpr uint32_t kTmemCols = 512;
constexpr int kFwdEpiWarps = 8;                              // (4 TMEM lane quarters) x (2 column groups of a 128-column chunk); a 16-warp
                                                            // (4 column groups, 96 registers) variant was tried: slower issuer, no gain
constexpr int kPW = 16 / kFwdEpiWarps;                      // 32-column pieces per warp and chunk: 2 or 1
constexpr int kGroupWarps = kFwdEpiWarps / 2;
constexpr int kFwdEpiThreads = kFwdEpiWarps * 32;
constexpr int kFwdThreads = (kFwdEpiWarps + kCtrlWarps) * 32; // 384 = 12 warps (registers are granted per 4 warps: 13 would cost like 16)

struct FwdShared {
    uint64_t w_full[kStages], w_empty[kStages];
    uint64_t a_ready[3], pe_ready, pe_free, pe_consumed, dens_done;
    uint64_t in_full[2], in_empty[2];   // data gradients: the dL/dfeat ring
    // per accumulator chunk n, rotating over 8 (the issuer can run at most a few chunks ahead of the slowest epilogue warp):
    uint64_t acc_full[8];               //   committed by the MMA issuer
    uint64_t loaded[8];                 //   all eight epilogue warps have read their part of the accumulator
    uint64_t stg_full[2], stg_free[2];  // staging buffers of saved chunks, alternating per saved chunk
    alignas(16) float dens[2][128];     // density head: every warp adds its partial dot products here (by tile parity)
    alignas(16) float w_density[HN_HIDDEN];   // read as float4
    uint32_t tmem_base;
    volatile int abort;
};
template <bool BWD> constexpr uint32_t chain_smem() { return Lay<BWD>::kOffShared + sizeof(FwdShared); }
constexpr uint32_t kFwdSmem = chain_smem<false>();
static_assert(Lay<false>::kOffShared % 16 == 0 && Lay<true>::kOffShared % 16 == 0 && chain_smem<false>() <= 232448 && chain_smem<true>() <= 232448,
              "shared-memory budget (227 KiB per CTA)");

// one EpiOp2 (16 bytes) fetched with a single 128-bit constant load and decoded with shifts: the fields stay in registers
// instead of being re-read from the constant bank wherever they are used
struct EpiFields {
    uint32_t acc_col, out_col, bias_off, col0, save_blk, mask_word, width32, kind, ready_idx, density, wait_next, signal_p;
};
template <bool BWD> __device__ __forceinline__ EpiFields load_epi(int e) {
    const uint4 r = reinterpret_cast<const uint4*>(BWD ? c_bwdt.epi : c_fwd.epi)[e];
    EpiFields f;
    f.acc_col = r.x & 0xFFFFu; f.out_col = r.x >> 16;
    f.bias_off = r.y & 0xFFFFu; f.col0 = r.y >> 16;
    f.save_blk = r.z & 0xFFFFu; f.mask_word = r.z >> 16;
    f.width32 = r.w & 0xFFu; f.kind = (r.w >> 8) & 0xFFu; f.ready_idx = (r.w >> 16) & 0xFFu;
    f.density = (r.w >> 24) & 3u; f.wait_next = (r.w >> 26) & 1u; f.signal_p = (r.w >> 27) & 1u;
    return f;
}

// Sampling + positional encoding (NetWorks/utils.py:20-51,147-161) of row `row` of tile `t`: the first GEMM's operand is
// generated, not loaded.  Channel order: p(3), then per frequency 2^k: sin(3), cos(3); column 63 is the zero pad of the
// 64-wide K block.
__device__ __forceinline__ void produce_pe_row(const hn_camera_t cam, float* delta, float* zvals, uint32_t pe_block, int t, int tiles_per_item, int row) {
    const size_t mm = (size_t)t * HN_TILE + row;
    const int bb = t / tiles_per_item;
    const int ns = cam.n_samples;
    const size_t ray_idx = mm / ns;
    const int s = (int)(mm % ns), r = (int)(ray_idx % cam.n_rays);
    const Ray ray = make_ray(cam, bb, r);
    const Sample q = make_sample(cam, ray, bb, r, s);
    delta[mm] = q.zdist;
    if (zvals) zvals[mm] = q.zval;
    const float p[3] = {q.px, q.py, q.pz};
    float v[64];
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[63] = 0.f;
#ifdef HN_DIAG_NO_PE_TRIG                                            // diagnostic build (timing only): no trigonometry
#pragma unroll
    for (int k = 0; k < 60; ++k) v[3 + k] = p[k % 3];
#else
    // sin / cos of 2^k p for k = 0..9: accurate sincosf at k = 0 and k = 5, angle doubling in between (sin 2a = 2 sin a cos a,
    // cos 2a = 1 - 2 sin^2 a).  Four doublings grow the anchor's rounding error to <= ~2e-6, two orders of magnitude below the
    // half-precision rounding (2.4e-4) the operand undergoes next; 6 instead of 30 accurate evaluations per sample took 0.1 ms
    // off the kernel (the trigonometry competes with the MMA issuer and two epilogue warps for one scheduler).
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float sn, cs;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            if (k == 0 || k == 5) sincosf(p[d] * (float)(1 << k), &sn, &cs);
            else { const float s2 = 2.f * sn * cs; cs = fmaf(-2.f * sn, sn, 1.f); sn = s2; }
            v[3 + 6 * k + d] = sn; v[3 + 6 * k + 3 + d] = cs;
        }
    }
#endif
    const uint32_t row_addr = pe_block + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c)
        st_shared_v4(row_addr + ((c ^ (row & 7)) << 4), pack_h2(v[8 * c + 0], v[8 * c + 1]), pack_h2(v[8 * c + 2], v[8 * c + 3]),
                     pack_h2(v[8 * c + 4], v[8 * c + 5]), pack_h2(v[8 * c + 6], v[8 * c + 7]));
}

template <bool BWD>
__global__ void __launch_bounds__(kFwdThreads, 1) mlp_chain_kernel(const ChainArgs a, const int n_tiles, const int tiles_per_item) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];      // no static shared memory in this kernel: the window starts aligned
    constexpr uint32_t kOffW = Lay<BWD>::kOffW, kOffStg = Lay<BWD>::kOffStg, kOffBias = Lay<BWD>::kOffBias, kOffShared = Lay<BWD>::kOffShared;
    const FwdTables& T = BWD ? c_bwdt : c_fwd;
    FwdShared& sh = *reinterpret_cast<FwdShared*>(smem_raw + kOffShared);
    const uint32_t smem = smem_u32(smem_raw);
    // warps 0..7: epilogue (TMEM lane quarter = warp & 3); warps 8..11: control roles.  The warp scheduler favours the
    // highest warp id of a sub-partition, so the MMA issuer and the producers - short instruction streams that everything
    // else waits for - sit at the top.
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cw = warp - kFwdEpiWarps;                           // control role: 0 weight producers, 1 MMA issuer, 2 TMEM + positional encoding, 3 saver
    const bool saving = (a.act != nullptr);
    const int work0 = (int)blockIdx.x, work_stride = (int)gridDim.x;

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(&sh.w_full[i]), 1); mbar_init(smem_u32(&sh.w_empty[i]), 1); }
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&sh.a_ready[i]), kFwdEpiWarps);
        mbar_init(smem_u32(&sh.pe_ready), 1); mbar_init(smem_u32(&sh.pe_free), 1); mbar_init(smem_u32(&sh.pe_consumed), kFwdEpiWarps);
        mbar_init(smem_u32(&sh.dens_done), kFwdEpiWarps);
        for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&sh.acc_full[i]), 1); mbar_init(smem_u32(&sh.loaded[i]), kFwdEpiWarps); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&sh.stg_full[i]), kFwdEpiWarps); mbar_init(smem_u32(&sh.stg_free[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&sh.in_full[i]), 1); mbar_init(smem_u32(&sh.in_empty[i]), 1); }
        sh.abort = 0;
        mbar_fence_init();
    }
    if (cw == 2) tmem_alloc<kTmemCols>(smem_u32(&sh.tmem_base));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = sh.tmem_base;
    const int n_stages = T.n_stages, n_epis = T.n_epis;
    const uint32_t tile_flip = (uint32_t)T.tile_flip;

    if (cw == 0) {
        // ======================= weight producers: three lanes of one warp, each with its own copies in flight =======================
        if (lane < 3) {
            const uint32_t P = 3u, p = (uint32_t)lane;
            uint32_t pc = 0;                                         // stage (unit pair) counter
            const uint8_t* packed = (const uint8_t*)a.packed;
            for (int w = work0; w < n_tiles && !sh.abort; w += work_stride) {
                for (int st = 0; st < n_stages; ++st, ++pc) {
                    if (pc % P != p) continue;
                    const uint32_t stage = pc % kStages, par = (pc / kStages) & 1;
                    if (!wait_spin(&sh.w_empty[stage], par ^ 1, &sh.abort, a.status, 101)) break;
                    const uint32_t fb = smem_u32(&sh.w_full[stage]);
                    mbar_arrive_expect_tx(fb, kStageBytes);
                    // one 32 KiB copy per stage: a thread's bulk copies complete one after the other at ~700 cycles apiece
                    // whatever their size, so bigger copies are what buys bandwidth
                    bulk_g2s(smem + kOffW + stage * kStageBytes, packed + (size_t)st * kStageBytes, kStageBytes, fb);
                }
            }
        }
    } else if (cw == 3) {
        if (!saving) { /* nothing to save */ } else
        // ======================= saver: staged activation chunks + the PE block -> HBM operand images =======================
        if (lane == 0) {
            uint32_t sidx = 0, par_pe = 0;                           // saved chunks so far: buffer = sidx & 1, its use = sidx >> 1
            for (int w = work0; w < n_tiles && !sh.abort; w += work_stride) {
                const int tile = w;
                if (!BWD) {
                    if (!wait_spin(&sh.pe_ready, par_pe, &sh.abort, a.status, 150)) break;
                    par_pe ^= 1;
                    bulk_s2g((uint8_t*)a.act + ((size_t)HN_SLOT_PE * n_tiles + tile) * kUnitBytes, smem + kOffPE, kUnitBytes);
                    bulk_commit();
                    bulk_wait_read<0>();
                    mbar_arrive(smem_u32(&sh.pe_free));
                }
                for (int e = 0; e < n_epis; ++e) {
                    const EpiFields op = load_epi<BWD>(e);
                    if (op.save_blk == 0xFFFF) continue;
                    const uint32_t sb = sidx & 1;
                    if (!wait_spin(&sh.stg_full[sb], (sidx >> 1) & 1, &sh.abort, a.status, 151)) break;
                    ++sidx;
                    const int nblk = (op.width32 + 1) / 2;
                    for (int k = 0; k < nblk; ++k)
                        bulk_s2g((uint8_t*)a.act + ((size_t)(op.save_blk + k) * n_tiles + tile) * kUnitBytes,
                                 smem + kOffStg + (sb * 2 + k) * kUnitBytes, kUnitBytes);
                    bulk_commit();
                    bulk_wait_read<0>();                               // the buffer is free as soon as the engine has read it
                    mbar_arrive(smem_u32(&sh.stg_free[sb]));
                }
            }
            bulk_wait_all<0>();
        }
    } else if (cw == 1) {
        // ======================= MMA issuer =======================
        // the whole warp walks the schedule (uniform control flow keeps descriptors in uniform registers); one
        // elected lane issues the MMAs and commits
        uint32_t pc = 0, par_ready = 0, par_pe = 0, base_n = 0, in_use = 0;
        int tile_k = 0;
        HN_PC_DECL(pcn, 8);
        // The issuer shares its scheduler with four epilogue warps and gets only a fraction of the issue slots, so its
        // instruction stream per MMA is what bounds the tensor pipe: a stage is one 128-bit table word, four MMAs of N = 256
        // (or eight of N = 128), one release; the NEXT stage's weight barrier is queried before its answer is needed.
        bool pre = mbar_try_wait(smem_u32(&sh.w_full[0]), 0);
        const uint4* table = reinterpret_cast<const uint4*>(T.stage);
        uint4 cur = table[0];
        for (int w = work0; w < n_tiles && !sh.abort; w += work_stride, base_n += n_epis, ++tile_k) {
            const uint32_t hx = ((uint32_t)tile_k & tile_flip) << 8;   // odd tiles: the TMEM halves swap roles (odd layer counts)
            for (int st = 0; st < n_stages; ++st, ++pc) {
                const uint4 nxt = table[st + 1 < n_stages ? st + 1 : 0];           // table read off the critical path
                const uint32_t stage = pc % kStages, par = (pc / kStages) & 1;
                const uint32_t a0 = cur.x & 0xFFFFu, a1 = cur.x >> 16, acc = cur.y & 0xFFFFu, n8 = (cur.y >> 16) & 0xFFu, first = cur.y >> 24;
                const uint32_t commit = cur.z & 0xFFu, chunk = (cur.z >> 8) & 0xFFu;
                bool ok = true;
                HN_TR(lane == 0, st * 4 + 0);
                if (cur.z >> 16) {                                     // rare: an input slot or chunk 0's accumulator must be awaited
                    const uint32_t wait_src = (cur.z >> 16) & 0xFFu;
                    if (wait_src == 4) {
                        if (BWD) ok = wait_spin(&sh.in_full[in_use & 1], (in_use >> 1) & 1, &sh.abort, a.status, 205);   // streamed dL/dfeat block
                        else { ok = wait_spin(&sh.pe_ready, par_pe, &sh.abort, a.status, 201); par_pe ^= 1; }
                    }
                    else if (wait_src) {
                        const int c = (int)wait_src - 1;
                        ok = wait_spin(&sh.a_ready[c], (par_ready >> c) & 1, &sh.abort, a.status, 202 + c);
                        par_ready ^= 1u << c;
                    }
                    if (ok && (cur.z >> 24)) {                         // chunk 2 reuses chunk 0's accumulator: all of it must have been read
                        const uint32_t np = base_n + chunk - 2;
                        ok = wait_spin(&sh.loaded[np & 7], (np >> 3) & 1, &sh.abort, a.status, 210);
                    }
                }
                HN_TR(lane == 0, st * 4 + 1);
                if (ok && !pre) ok = wait_spin(&sh.w_full[stage], par, &sh.abort, a.status, 220);
                HN_TR(lane == 0, st * 4 + 2);
                if (!ok) break;
                tc_fence_after_sync();
                const uint32_t b_lo = desc_lo(smem + kOffW + stage * kStageBytes, 16);
                const uint32_t idesc = umma_idesc(128, n8 * 8, kF16, kF16, 0, 0);
                const uint32_t d = tmem_base + (acc ^ hx);
                const uint32_t acc0 = first ? 0u : 1u;
                const uint32_t n0 = base_n + chunk, n1 = n0 + 1;
                const uint32_t full0 = smem_u32(&sh.acc_full[n0 & 7]), full1 = smem_u32(&sh.acc_full[n1 & 7]);
                if (a0 & kSrcSmem) {
                    const uint32_t a_lo = desc_lo(smem + kOffPE + (BWD ? (a0 & 1u) : (a0 & 0x7FFFu)) * kUnitBytes, 16);
                    const uint32_t in_bar = smem_u32(&sh.in_empty[in_use & 1]);
                    if (elect_one()) {
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_lohi(d, a_lo + ks * 2, b_lo + ks * 2, idesc, ks == 0 ? acc0 : 1u);
                        if (a1 != kSrcNone) {                          // (both K blocks of a stage come from the same kind of source)
#pragma unroll
                            for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_lohi(d, a_lo + ks * 2, b_lo + (kUnitBytes >> 4) + ks * 2, idesc, 1u);
                        }
                        if (commit) umma_commit(full0);
                        if (commit == 2) umma_commit(full1);
                        umma_commit(smem_u32(&sh.w_empty[stage]));
                        if (BWD) umma_commit(in_bar);                   // the ring slot may be refilled once these MMAs retire
                    }
                    if (BWD) ++in_use;
                } else {
                    const uint32_t a_t0 = tmem_base + (a0 ^ hx), a_t1 = tmem_base + (a1 ^ hx);
                    if (elect_one()) {
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ts_lo(d, a_t0 + ks * 8, b_lo + ks * 2, idesc, ks == 0 ? acc0 : 1u);
                        if (a1 != kSrcNone) {
#pragma unroll
                            for (uint32_t ks = 0; ks < 4; ++ks) umma_f16_ts_lo(d, a_t1 + ks * 8, b_lo + (kUnitBytes >> 4) + ks * 2, idesc, 1u);
                        }
                        if (commit) umma_commit(full0);
                        if (commit == 2) umma_commit(full1);
                        umma_commit(smem_u32(&sh.w_empty[stage]));
                    }
                }
                __syncwarp();
                // ask for the next stage's weights now; the answer is consumed at the top of the next iteration
                pre = mbar_try_wait(smem_u32(&sh.w_full[(pc + 1) % kStages]), ((pc + 1) / kStages) & 1);
                HN_TR(lane == 0, st * 4 + 3);
                cur = nxt;
            }
        }
        HN_PC_FLUSH(pcn, 8, a.status + 2, blockIdx.x == 0 && lane == 0);
    } else if (cw == 2) {
        // ======================= PE warp: sampling + positional encoding of the NEXT tile, off everybody's critical path =======================
        // four rows per lane; the PE block is free once FeaExt_module_5's last chunk has been committed (its MMAs were the last
        // readers) and - when activations are saved - the saver's bulk store has read it
        if (BWD) {
            // data gradients: this warp's lane 0 streams the tile's four dL/dfeat K blocks through the two-block ring
            if (lane == 0) {
                uint32_t i = 0;
                const uint8_t* din = (const uint8_t*)a.dfeat_image;
                for (int w = work0; w < n_tiles && !sh.abort; w += work_stride) {
                    for (int kb = 0; kb < 4; ++kb, ++i) {
                        if (!wait_spin(&sh.in_empty[i & 1], ((i >> 1) & 1) ^ 1, &sh.abort, a.status, 162)) break;
                        const uint32_t fb = smem_u32(&sh.in_full[i & 1]);
                        mbar_arrive_expect_tx(fb, kUnitBytes);
                        bulk_g2s(smem + kOffPE + (i & 1) * kUnitBytes, din + ((size_t)kb * n_tiles + w) * kUnitBytes, kUnitBytes, fb);
                    }
                }
            }
        } else {
        uint32_t k = 0;
        for (int w = work0; w < n_tiles && !sh.abort; w += work_stride, ++k) {
            if (k > 0) {
                if (!wait_spin(&sh.pe_consumed, (k - 1) & 1, &sh.abort, a.status, 160)) break;
                if (saving && !wait_spin(&sh.pe_free, (k - 1) & 1, &sh.abort, a.status, 161)) break;
            }
#pragma unroll 1
            for (int i = 0; i < 4; ++i) produce_pe_row(a.cam, a.delta, a.zvals, smem + kOffPE, w, tiles_per_item, i * 32 + lane);
            fence_async_smem();
            warp_arrive(smem_u32(&sh.pe_ready), lane);
        }
        }
    } else {
        // ======================= epilogue =======================
        // Eight warps, every one on EVERY accumulator chunk: warp = (TMEM lane quarter, column half); it drains 32 rows x 64
        // columns as two 32-column pieces.  Splitting a chunk by columns (rather than handing whole chunks to alternating groups)
        // halves the latency from "chunk committed" to "accumulator read" - which the MMA issuer waits for before chunk 2 can
        // reuse chunk 0's columns - and to "outputs stored".  Few, fat warps on purpose: the issuer shares its scheduler with
        // two of them.
        const int ew = warp;
        const int g = ew >> 2, quarter = ew & 3;                    // g: columns [32 kPW g, 32 kPW (g+1)) of every chunk
        constexpr int CW = 32 * kPW;                                // accumulator columns per warp
        const int row = quarter * 32 + lane;                        // tile row = TMEM lane
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        uint32_t base_n = 0, sidx = 0, tile_i = 0;                  // first chunk index of the tile; saved chunks so far
        const int pe_after = T.pe_after_epi;
        int cached_b = -1;
        for (int i = tid; i < HN_HIDDEN; i += kFwdEpiThreads) sh.w_density[i] = __ldg(a.w_density + i);
        if (tid < 256) (&sh.dens[0][0])[tid] = 0.f;
        named_sync(3, kFwdEpiThreads);

        HN_PC_DECL(ec, 16);
        float dens = 0.f;
        const float gscale = BWD ? __ldg(a.grad_scale) : 1.0f;
        for (int w = work0; w < n_tiles && !sh.abort; w += work_stride, base_n += n_epis, ++tile_i) {
            const int tile = w;
            const size_t m = (size_t)tile * HN_TILE + row;
            const int b = tile / tiles_per_item;
            // data gradients: d/d(pre-ReLU density) of this row, loss-scaled (the density head's rank-1 term and pseudo layer)
            const float dsr = (BWD && __ldg(a.sigma + m) > 0.f) ? __ldg(a.dsigma + m) * gscale : 0.f;
            if (!BWD && b != cached_b) {                            // (re)load the item's bias row
                named_sync(3, kFwdEpiThreads);
                const float4* src = reinterpret_cast<const float4*>(a.bias + (size_t)b * HN_BIAS_STRIDE);
                for (int i = tid; i < HN_BIAS_STRIDE / 4; i += kFwdEpiThreads) {
                    const float4 v4 = __ldg(src + i);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(smem + kOffBias + i * 16), "f"(v4.x), "f"(v4.y), "f"(v4.z), "f"(v4.w) : "memory");
                }
                named_sync(3, kFwdEpiThreads);
                cached_b = b;
            }
            const uint32_t bias_row = smem + kOffBias;
            const uint32_t hx = (tile_i & tile_flip) << 8;          // odd tiles: the TMEM halves swap roles (odd layer counts)
            uint32_t hold[kPW][16];                                 // packed output of chunk 0 of a pair, stored with chunk 1's
            uint32_t hold_addr = 0;
            bool holding = false;
            for (int e = 0; e < n_epis; ++e) {
                const uint32_t n = base_n + e;
                const EpiFields op = load_epi<BWD>(e);
                uint32_t mwords[kPW];
                if (BWD && op.mask_word != 0xFFFF && kPW * g < (int)op.width32) {   // saved ReLU masks: fetched ahead of the accumulator
#pragma unroll
                    for (int pc = 0; pc < kPW; ++pc) mwords[pc] = __ldg(a.masks + m * HN_MASK_WORDS + op.mask_word + kPW * g + pc);
                }
                HN_PC_T0(ec);
                wait_spin(&sh.acc_full[n & 7], (n >> 3) & 1, &sh.abort, a.status, 300 + e);
                HN_PC_LAP(ec, 1);
                { const int tile_k = (int)tile_i; HN_TR(ew == 0 && lane == 0, 1024 + e * 4 + 0); }
                tc_fence_after_sync();
                // FeaExt_module_5's last chunk is complete: every MMA that reads the PE block has run
                if (!BWD && e == pe_after) warp_arrive(smem_u32(&sh.pe_consumed), lane);
                const bool save = saving && op.save_blk != 0xFFFF;
                const bool active = kPW * g < (int)op.width32;     // RGB_layer_1's second chunk has 64 columns: the upper column groups idle
                const uint32_t bp = bias_row + (op.bias_off + CW * g) * 4;
                const uint32_t acc_addr = tmem_base + lane_base + (op.acc_col ^ hx) + CW * g;
                uint32_t v[kPW][32];
                if (active) {
#pragma unroll
                    for (int pc = 0; pc < kPW; ++pc) tmem_ld32(acc_addr + 32 * pc, v[pc]);
                }
                tmem_ld_wait();
                tc_fence_before_sync();
                warp_arrive(smem_u32(&sh.loaded[n & 7]), lane);    // (the MMA issuer waits on this for chunk 0 before it starts chunk 2)
                { const int tile_k = (int)tile_i; HN_TR(ew == 0 && lane == 0, 1024 + e * 4 + 1); }
                HN_PC_LAP(ec, 2);
                const uint32_t sb = sidx & 1;                      // staging buffer of this saved chunk
                if (!BWD && op.kind == EPI_FEAT) {
                    // final features: stage 32 rows x 32 columns per piece in shared memory (swizzled 16-byte chunks), then write
                    // whole 128-byte row segments (8 lanes each) instead of 32 scattered 16-byte pieces (every thread storing its
                    // own row directly was re-measured on this kernel: 1.86 vs 1.73 ms - these stores sit on the tile-boundary path).  Scratch = the staging
                    // buffer whose last saved chunk is the older one (chunk 29 -> buffer of chunk 27, chunk 30 -> of chunk 28).
                    const uint32_t fb = (sidx + (uint32_t)(e & 1 ? 0 : 1)) & 1;     // e = 29: sidx & 1;  e = 30: (sidx + 1) & 1
                    const uint32_t fidx = sidx + (uint32_t)(e & 1 ? 0 : 1);
                    if (saving) {
                        wait_spin(&sh.stg_free[fb], ((fidx >> 1) & 1) ^ 1, &sh.abort, a.status, 340);
                        if (kFwdEpiWarps == 16) wait_spin(&sh.stg_free[fb ^ 1], (((fidx + 1) >> 1) & 1) ^ 1, &sh.abort, a.status, 342);   // scratch spans both buffers
                    }
                    const uint32_t stg = smem + kOffStg + (kFwdEpiWarps == 8 ? fb * 2 * kUnitBytes : 0u) + (uint32_t)ew * 4096;   // 32 rows x 128 B per warp
                    float* gbase = a.feat + ((size_t)tile * HN_TILE + quarter * 32) * HN_FEAT + op.col0 + CW * g;
                    auto emit = [&](const uint32_t (&v)[32], int pc) {
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
                    };
#pragma unroll
                    for (int pc = 0; pc < kPW; ++pc) emit(v[pc], pc);
                    HN_PC_LAP(ec, 9);
                } else {
                    auto convert = [&](const uint32_t (&v)[32], int pc, uint32_t (&pk)[16]) {
                        if (BWD) {
                            // dX = (acc + dsigma * w_density) * [forward pre-activation > 0]
                            float y[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(v[i]);
                            if (op.kind == EPI_GRAD_DENSITY) {
                                const float4* wp = reinterpret_cast<const float4*>(sh.w_density + op.col0 + CW * g + pc * 32);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 ww = wp[i];
                                    y[4 * i + 0] = fmaf(dsr, ww.x, y[4 * i + 0]); y[4 * i + 1] = fmaf(dsr, ww.y, y[4 * i + 1]);
                                    y[4 * i + 2] = fmaf(dsr, ww.z, y[4 * i + 2]); y[4 * i + 3] = fmaf(dsr, ww.w, y[4 * i + 3]);
                                }
                            }
                            if (op.kind != EPI_GRAD_LINEAR) {
                                const uint32_t mw = mwords[pc];
#pragma unroll
                                for (int i = 0; i < 32; ++i) y[i] = (mw & (1u << i)) ? y[i] : 0.f;
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) pk[i] = pack_sat(y[2 * i], y[2 * i + 1]);
                            return;
                        }
#ifdef HN_DIAG_NO_EPILOGUE_MATH                                     // diagnostic build (timing only): measured 0.2 ms of the forward
                        for (int i = 0; i < 16; ++i) pk[i] = v[i] ^ v[i + 16];
                        return;
#endif
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
                            if (!BWD && a.masks && op.mask_word != 0xFFFF)
                                a.masks[m * HN_MASK_WORDS + op.mask_word + kPW * g + pc] = positive_mask32(y);
                            if (op.density) {                      // density head on the fp32 activations (models.py:78,83)
                                const float4* wp = reinterpret_cast<const float4*>(sh.w_density + op.col0 + CW * g + pc * 32);
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
                    const uint32_t out_addr = tmem_base + lane_base + (op.out_col ^ hx) + (CW / 2) * g;
                    const bool has_out = !BWD || op.out_col != kNoCol;      // the chain's last data-gradient chunk is only saved
                    const uint32_t stg = smem + kOffStg + sb * 2 * kUnitBytes;
                    if (BWD && save && op.kind == EPI_GRAD_DENSITY && op.col0 == 0 && g == 0) {
                        // density head as a one-channel pseudo layer for the weight pass: row = [dsr, 0, ..., 0]
                        uint8_t* drow = (uint8_t*)a.act + ((size_t)HN_GSLOT_DENS * n_tiles + tile) * kUnitBytes + (row >> 3) * 1024 + (row & 7) * 128;
                        const uint32_t first = pack_sat(dsr, 0.f);
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            *reinterpret_cast<uint4*>(drow + ((c ^ (row & 7)) << 4)) = make_uint4(c == 0 ? first : 0u, 0u, 0u, 0u);
                    }
                    // staging buffer: has the bulk store of its previous saved chunk read it?
                    if (save) wait_spin(&sh.stg_free[sb], ((sidx >> 1) & 1) ^ 1, &sh.abort, a.status, 341);
                    if (op.wait_next) {
                        // chunk 0 of a pair: its output slot is the upper half of chunk 1's accumulator, which is read only in the
                        // next iteration - keep the packed columns in registers (the saved image can be staged now)
                        if (active) {
#pragma unroll
                            for (int pc = 0; pc < kPW; ++pc) {
                                convert(v[pc], pc, hold[pc]);
                                if (save) store_row_packed(stg, row, CW * g + 32 * pc, hold[pc]);
                            }
                        }
                        hold_addr = out_addr; holding = true;
                    } else {
                        // every warp has read its part of this accumulator (and, after a held chunk, of the previous one): stores
                        // into in-place or neighbouring columns are safe now
                        wait_spin(&sh.loaded[n & 7], (n >> 3) & 1, &sh.abort, a.status, 330);
                        tc_fence_after_sync();
                        if (holding) {
                            // chunk 0's output first, signalled at once: the next layer's first K blocks are what the MMA
                            // issuer will ask for as soon as chunk 2 is issued
#pragma unroll
                            for (int pc = 0; pc < kPW; ++pc) tmem_st16(hold_addr + 16 * pc, hold[pc]);
                            tmem_st_wait();
                            tc_fence_before_sync();
                            warp_arrive(smem_u32(&sh.a_ready[0]), lane);
                            holding = false;
                        }
                        if (active) {
#pragma unroll
                            for (int pc = 0; pc < kPW; ++pc) {
                                uint32_t pk[16];
                                convert(v[pc], pc, pk);
                                if (has_out) tmem_st16(out_addr + 16 * pc, pk);
                                if (save) store_row_packed(stg, row, CW * g + 32 * pc, pk);
                            }
                        }
                    }
                    HN_PC_LAP(ec, 4);
                    tmem_st_wait();
                    HN_PC_LAP(ec, 5);
                }
                if (save) { fence_async_smem(); warp_arrive(smem_u32(&sh.stg_full[sb]), lane); ++sidx; }
                if (op.kind != EPI_FEAT && !op.wait_next) {
                    tc_fence_before_sync();
                    if (op.ready_idx != 255) warp_arrive(smem_u32(&sh.a_ready[op.ready_idx]), lane);
                }
                HN_PC_LAP(ec, 6);
                { const int tile_k = (int)tile_i; HN_TR(ew == 0 && lane == 0, 1024 + e * 4 + 2); }
                if (!BWD && op.density == 2) {
                    // last density chunk of the tile: publish the partial dot products; once all 8 warps have contributed the
                    // lower-half warps add the bias and write sigma
                    atomicAdd(&sh.dens[tile_i & 1][row], dens);
                    dens = 0.f;
                    __threadfence_block();
                    warp_arrive(smem_u32(&sh.dens_done), lane);
                    if (g == 0) {
                        wait_spin(&sh.dens_done, tile_i & 1, &sh.abort, a.status, 350);
                        const float tot = *((volatile float*)&sh.dens[tile_i & 1][row]);
                        a.sigma[m] = fmaxf(tot + __ldg(a.bias + (size_t)b * HN_BIAS_STRIDE + HN_BIAS_OFF_DENSITY), 0.f);
                        sh.dens[tile_i & 1][row] = 0.f;            // next use is two tiles away
                    }
                }
                HN_PC_LAP(ec, 7);
            }
        }
        HN_PC_FLUSH(ec, 16, a.status + 18, blockIdx.x == 0 && ew == 0 && lane == 0);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (cw == 2) tmem_free<kTmemCols>(tmem_base);
}

static std::mutex g_fwd_mu;
static bool g_fwd_ready[64] = {};

}  // namespace hn

namespace hn {
static int prepare_chain_device(int* n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_fwd_mu);
        if (dev < 64 && !g_fwd_ready[dev]) {
            const HostSchedules& hs = host_schedules();
            cudaError_t e = cudaMemcpyToSymbol(c_fwd, &hs.fwd, sizeof(FwdTables));
            if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_bwdt, &hs.bwdt, sizeof(FwdTables));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem<false>());
            if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem<true>());
            if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
            g_fwd_ready[dev] = true;
        }
    }
    *n_sm = 148;
    cudaDeviceGetAttribute(n_sm, cudaDevAttrMultiProcessorCount, dev);
    return HN_OK;
}

// data-gradient chain without dL/dPE on the tensor-memory kernel (called by hn_mlp_bwd_data, hn_mlp_bwd.cu)
int launch_bwd_data_tmem(const hn_mlp_bwd_data_t* a, void* stream) {
    int n_sm = 148;
    if (int rc = prepare_chain_device(&n_sm)) return rc;
    const HostSchedules& hs = host_schedules();
    ChainArgs c{};
    c.cam = a->cam;
    c.w_density = a->w_density;
    c.packed = (const uint8_t*)a->packed + (size_t)(kFwdUnits + hs.n_bwd_pack_units) * kUnitBytes;
    c.sigma = const_cast<float*>(a->sigma);
    c.act = a->grads;
    c.masks = const_cast<uint32_t*>(a->masks);
    c.dfeat_image = a->dfeat_image;
    c.dsigma = a->dsigma;
    c.grad_scale = a->grad_scale;
    c.status = a->status;
    const int64_t M = total_samples(a->cam.B, a->cam.n_rays, a->cam.n_samples);
    const int n_tiles = (int)(M / HN_TILE);
    const int tiles_per_item = (int)(((int64_t)a->cam.n_rays * a->cam.n_samples) / HN_TILE);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    mlp_chain_kernel<true><<<grid, kFwdThreads, chain_smem<true>(), (cudaStream_t)stream>>>(c, n_tiles, tiles_per_item);
    return check_launch("hn_mlp_bwd_data (tensor-memory chain)");
}
}  // namespace hn

extern "C" int hn_mlp_fwd(const hn_mlp_fwd_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->cam.xy || !a->cam.Rmats || !a->cam.Tvecs || !a->cam.inv_inmats || !a->bias || !a->w_density ||
        !a->packed || !a->sigma || !a->delta || !a->status)
        return set_error(HN_E_BADARG, "hn_mlp_fwd: null pointer");
    if (int rc = check_geometry(a->cam.B, a->cam.n_rays, a->cam.n_samples, "hn_mlp_fwd")) return rc;
    if ((a->act == nullptr) != (a->masks == nullptr))
        return set_error(HN_E_BADARG, "hn_mlp_fwd: act and masks must both be given (backward) or both be NULL");
    int n_sm = 148;
    if (int rc = prepare_chain_device(&n_sm)) return rc;
    ChainArgs c{};
    c.cam = a->cam; c.bias = a->bias; c.w_density = a->w_density; c.packed = a->packed;
    c.feat = a->feat; c.sigma = a->sigma; c.delta = a->delta; c.zvals = a->zvals;
    c.act = a->act; c.masks = a->masks; c.status = a->status;
    const int64_t M = total_samples(a->cam.B, a->cam.n_rays, a->cam.n_samples);
    const int n_tiles = (int)(M / HN_TILE);
    const int tiles_per_item = (int)(((int64_t)a->cam.n_rays * a->cam.n_samples) / HN_TILE);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    mlp_chain_kernel<false><<<grid, kFwdThreads, kFwdSmem, (cudaStream_t)stream>>>(c, n_tiles, tiles_per_item);
    return check_launch("hn_mlp_fwd");
}
