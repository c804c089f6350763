// hn_nr.cu — the consumer of the composited feature map, NeuralRenderer (NetWorks/neural_renderer.py:11-91,
// NetWorks/PixelShuffleUpsample.py:8-45), forward and backward, behind ONE library call per
// direction.  Every 1x1 convolution, its LeakyReLU, the RGB heads with their skip sums and the final sigmoid run in one
// grouped tensor-core GEMM kernel (tcgen05 kind::tf32, fp32 accumulation in tensor memory) that reads and writes the NCHW
// planes directly; the memory-bound tails (leaky-relu + residual + pixel shuffle + blur, bilinear x2 + blur) are the kernels
// of hn_render2d.cu.  With n blocks the forward call issues 5 n kernels and the backward call 6 n + 1 (Reso32HR, four blocks up
// to 512 x 512: 45 launches for the whole renderer, both directions, from two C calls).
//
// One GEMM problem:  D[m, n] = sum_k A(m, k) * B(n, k), both operands fetched element-wise as base[row * rs + k * ks]:
//   pixel rows   (forward / data gradient): A = an NCHW activation (m = pixel: rs = 1, ks = plane stride), B = the weight
//                [Cout, Cin] (forward: rs = Cin, ks = 1) or its transpose (data gradient: rs = 1, ks = Cin);
//                epilogue: + bias, LeakyReLU, x LeakyReLU'(saved activation), + addend, store plane-major; optionally the RGB
//                head of the block: rgb_out = rgb_in + W_rgb . act + b_rgb (+ sigmoid), a per-pixel dot product over the
//                accumulator row the thread already holds.
//   weight gradient: A = the pre-activation gradient (m = out channel, k = pixel: rs = plane stride, ks = 1), B = the layer
//                input (n = in channel), contraction over the pixels of all items split across CTAs; partial products are
//                added to the gradient buffer with atomics; the bias gradient is the row sum of A, taken while loading it.
// Operands pass through registers on their way into the canonical SWIZZLE_128B K-major shared-memory layout (128-byte rows of
// 32 tf32): NCHW planes with rows = pixels are transposed there (float4 loads along the pixels, 4 x 4 register transpose, 16-byte
// row chunks), rounded to tf32 with one integer add; a two-stage ring feeds four K = 8 MMAs per 32-wide block.
// (kind::tf32 accepts K-major operands only - with either major bit set the accumulator comes back all zero, tools/probe_tf32.cu -
// so the row-contiguous planes cannot be handed to the tensor core as they lie in memory.)
// Measured (profiles/r02_nr_*): the whole renderer forward + backward at Reso32HR, batch 2: 1.7 ms (the module-by-module cuDNN
// path: 3.3 ms incl. its launch gaps).  What bounds it: the low-resolution layers are chains of 8-16 dependent global-memory
// round trips on 48-128 CTAs; the high-resolution layers run ~5 us per 128-pixel tile with three CTAs per SM (registers);
// a multi-tile streaming variant was measured and brought nothing once its register prefetch spilled.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "hn_api.h"
#include "hn_tc.cuh"

extern "C" int hn_upsample_tail_fwd(const float*, const float*, const float*, float*, int, int, int, int, void*);
extern "C" int hn_upsample_tail_bwd(const float*, const float*, const float*, float*, float*, int, int, int, int, void*);
extern "C" int hn_rgb_upsample_fwd(const float*, const float*, float*, int, int, int, void*);
extern "C" int hn_rgb_upsample_bwd(const float*, const float*, float*, int, int, int, void*);

namespace hn {

#ifndef HN_NR_MIN_CTAS
#define HN_NR_MIN_CTAS 2
#endif
constexpr int kNrThreads = 256;                              // loader / epilogue threads (8 warps); a ninth warp issues the MMAs
constexpr int kNrCtaThreads = kNrThreads + 32;
constexpr int kNrMaxN = 256;
constexpr uint32_t kNrStageA = 128 * 128;                    // 128 rows x 32 tf32
constexpr uint32_t kNrStageB = kNrMaxN * 128;
constexpr uint32_t kNrStage = kNrStageA + kNrStageB;         // 48 KiB
constexpr uint32_t kNrSmem = 2 * kNrStage + 1024;            // two stages + slack for the 1 KiB alignment
constexpr int kNrMaxProblems = 4;
constexpr float kSlope = 0.2f;

struct NrProb {
    const float* a; const float* b;
    float* out;
    const float* bias; const float* mask_act; const float* addend;
    const float* wrgb; const float* brgb; const float* rgb_in; float* rgb_out;
    float* dbias;
    long long a_item, b_item, out_item;
    int a_rs, a_ks, b_rs, b_ks;
    int M, N, K;                 // valid rows of A (per item), valid rows of B, contraction length (per item)
    int kind;                    // 0 = pixel rows, 1 = weight gradient
    int m_tiles, n_tiles, n_tile, k_chunk, k_chunks, out_ld;
    int lrelu, sigmoid, epi;
    int tile0, tiles;
};
struct NrLaunch { NrProb p[kNrMaxProblems]; int n; int tmem_cols; int async_loads; uint32_t stage_bytes, window_bytes; int* status; };   // tmem_cols / stage_bytes: sized for the widest tile of the launch

// fp32 -> tf32 operand bits: the tensor core reads the upper 19 bits of the word, so adding half an ulp of the 10-bit mantissa
// rounds to nearest (ties away) in ONE integer add; cvt.rna.tf32.f32 compiles to four instructions per element on sm_100a,
// which made the conversion the largest item of the loader.  Finite inputs only (+-inf would turn into NaN).
__device__ __forceinline__ uint32_t to_tf32(float v) { return __float_as_uint(v) + 0x1000u; }
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// split-descriptor form (hn_tc.cuh): high word constant, low word = (address >> 4) | (LBO >> 4) << 16; a K = 8 tf32 step is 32 bytes (+2)
__device__ __forceinline__ void umma_tf32_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi) : "memory");
}

struct Operand { const float* base; long long rs, ks; int rows_valid, rows_tile; bool vec; };

// chunk i of a [rows_tile x 32] block = 4 consecutive k of one row; k-contiguous operands put 8 lanes on one 128-byte row,
// row-contiguous operands put consecutive lanes on consecutive rows (both coalesce)
__device__ __forceinline__ void chunk_of(const Operand& o, int i, int* row, int* c) {
    if (o.ks == 1) { *row = i >> 3; *c = i & 7; }
    else { *row = i % o.rows_tile; *c = i / o.rows_tile; }
}
// Zero source for everything outside the valid rows / contraction range: the loads themselves stay unconditional (an `if`
// around a load makes the compiler wait for it before the next one is issued - 12 serialised DRAM round trips per block).
__device__ float4 g_nr_zero = {0.f, 0.f, 0.f, 0.f};

template <int NJ, bool VEC>
__device__ __forceinline__ void gload(float4 (&r)[NJ], const Operand& o, int k_valid, int tid) {
    const float* zero = reinterpret_cast<const float*>(&g_nr_zero);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int i = tid + kNrThreads * j;
        int row, c;
        chunk_of(o, i, &row, &c);
        const int k = 4 * c;
        const bool rv = i < o.rows_tile * 8 && row < o.rows_valid;
        const float* p = o.base + row * o.rs + k * o.ks;
        if (VEC) {                                            // k-contiguous, 16-byte aligned, contraction length a multiple of 4
            r[j] = __ldg(reinterpret_cast<const float4*>((rv && k < k_valid) ? p : zero));
        } else {
            const float* p0 = (rv && k < k_valid) ? p : zero;
            const float* p1 = (rv && k + 1 < k_valid) ? p + o.ks : zero;
            const float* p2 = (rv && k + 2 < k_valid) ? p + 2 * o.ks : zero;
            const float* p3 = (rv && k + 3 < k_valid) ? p + 3 * o.ks : zero;
            r[j] = make_float4(__ldg(p0), __ldg(p1), __ldg(p2), __ldg(p3));
        }
    }
}
template <int NJ>
__device__ __forceinline__ void sstore(const float4 (&r)[NJ], const Operand& o, uint32_t stage, int tid) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int i = tid + kNrThreads * j;
        if (i < o.rows_tile * 8) {
            int row, c;
            chunk_of(o, i, &row, &c);
            st_shared_v4(stage + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4),
                         to_tf32(r[j].x), to_tf32(r[j].y), to_tf32(r[j].z), to_tf32(r[j].w));
        }
    }
}

struct NrShared {
    uint64_t stage_free[4], stage_full[4], done;
    uint32_t tmem_base;
    float rgb_part[128 * 3];
    float bias_s[kNrMaxN], wrgb_s[3][kNrMaxN], brgb_s[4];               // this tile's columns of the bias / RGB-head weights (zero beyond N)
};

// ---- fast loaders: the thread -> (row, chunk) map is fixed at compile time, so a block costs a handful of address adds
// per load instead of integer divisions (the generic functions above remain the fallback for unaligned / odd geometries).
//   row-contiguous vector (RC, rs == 1, 16-byte aligned: NCHW activations with rows = pixels, transposed weights): thread =
//       a 4 x 4 micro-tile, rows 4 (tid % 32) .. + 3 (+ 128 j), k = 4 (tid / 32) .. + 3: four float4 loads along the rows
//       (512 contiguous bytes per warp), transposed in registers into four 16-byte row chunks
//   k-contiguous vector (KCV, ks == 1, 16-byte aligned: weights, NCHW planes with rows = channels): thread = rows
//       tid / 8 + 32 j, chunk tid % 8, one float4 load each
enum { kModeRC = 0, kModeKCV = 1, kModeGen = 2 };
struct LState { const float* p; long long ks, rstep; uint32_t soff; int r7, c0, jv, jt; };

template <int ROWS>
__device__ __forceinline__ LState lstate_rc(const Operand& o, int tid) {          // thread = 4 consecutive rows x 4 consecutive k
    LState s;
    const int kg = tid >> 5;
    s.p = o.base + 4 * (tid & 31) + 4 * kg * o.ks;
    s.ks = o.ks;
    s.rstep = 0;
    s.r7 = 4 * (tid & 1);
    s.c0 = kg;
    s.soff = ((tid & 31) >> 1) * 1024;
    const int lim = min(o.rows_valid, o.rows_tile), row0 = 4 * (tid & 31);
    s.jv = lim > row0 ? (lim - row0 + 127) / 128 : 0;                               // row groups row0 + 128 j inside the valid rows
    s.jt = o.rows_tile > row0 ? (o.rows_tile - row0 + 127) / 128 : 0;
    return s;
}
template <int ROWS>
__device__ __forceinline__ void gload_rc(float4 (&r)[ROWS / 32], const LState& s, int k_valid) {
    const float4* zero = &g_nr_zero;
    const int k = 4 * s.c0;
#pragma unroll
    for (int j = 0; j < ROWS / 128; ++j) {
        const float* q = s.p + 128 * j;
        const bool v = j < s.jv;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            r[4 * j + e] = __ldg((v && k + e < k_valid) ? reinterpret_cast<const float4*>(q + e * s.ks) : zero);
    }
}
template <int ROWS>
__device__ __forceinline__ void sstore_rc(const float4 (&r)[ROWS / 32], const LState& s, uint32_t stage) {
#pragma unroll
    for (int j = 0; j < ROWS / 128; ++j) {
        if (j >= s.jt) break;
        const uint32_t base = stage + s.soff + j * 16384;
        const float4 &a = r[4 * j], &b = r[4 * j + 1], &c = r[4 * j + 2], &d = r[4 * j + 3];     // a = k, b = k + 1, ...; .x = row 0 of the group
        st_shared_v4(base + (s.r7 + 0) * 128 + ((s.c0 ^ (s.r7 + 0)) << 4), to_tf32(a.x), to_tf32(b.x), to_tf32(c.x), to_tf32(d.x));
        st_shared_v4(base + (s.r7 + 1) * 128 + ((s.c0 ^ (s.r7 + 1)) << 4), to_tf32(a.y), to_tf32(b.y), to_tf32(c.y), to_tf32(d.y));
        st_shared_v4(base + (s.r7 + 2) * 128 + ((s.c0 ^ (s.r7 + 2)) << 4), to_tf32(a.z), to_tf32(b.z), to_tf32(c.z), to_tf32(d.z));
        st_shared_v4(base + (s.r7 + 3) * 128 + ((s.c0 ^ (s.r7 + 3)) << 4), to_tf32(a.w), to_tf32(b.w), to_tf32(c.w), to_tf32(d.w));
    }
}
template <int ROWS>
__device__ __forceinline__ LState lstate_kcv(const Operand& o, int tid) {
    LState s;
    const int row0 = tid >> 3, c = tid & 7;
    s.p = o.base + row0 * o.rs + 4 * c;
    s.ks = 4 * c;                                             // first k of this thread's chunk
    s.rstep = 32 * o.rs;
    s.r7 = row0 & 7;
    s.c0 = c;
    s.soff = (row0 >> 3) * 1024 + (row0 & 7) * 128 + ((c ^ (row0 & 7)) << 4);
    const int lim = min(o.rows_valid, o.rows_tile);
    s.jv = lim > row0 ? (lim - row0 + 31) / 32 : 0;
    s.jt = o.rows_tile > row0 ? (o.rows_tile - row0 + 31) / 32 : 0;
    return s;
}
template <int ROWS>
__device__ __forceinline__ void gload_kcv(float4 (&r)[ROWS / 32], const LState& s, int k_valid) {
    const float4* zero = &g_nr_zero;
    const bool kv = (int)s.ks < k_valid;
#pragma unroll
    for (int j = 0; j < ROWS / 32; ++j)
        r[j] = __ldg((kv && j < s.jv) ? reinterpret_cast<const float4*>(s.p + j * s.rstep) : zero);
}
template <int ROWS>
__device__ __forceinline__ void sstore_kcv(const float4 (&r)[ROWS / 32], const LState& s, uint32_t stage) {
#pragma unroll
    for (int j = 0; j < ROWS / 32; ++j)
        if (j < s.jt) st_shared_v4(stage + s.soff + j * 4096, to_tf32(r[j].x), to_tf32(r[j].y), to_tf32(r[j].z), to_tf32(r[j].w));
}

template <int MODE, int ROWS> struct Ld;
template <int ROWS> struct Ld<kModeRC, ROWS> {
    LState s;
    __device__ __forceinline__ Ld(const Operand& o, int tid) : s(lstate_rc<ROWS>(o, tid)) {}
    __device__ __forceinline__ void load(float4 (&r)[ROWS / 32], int k_valid, int) { gload_rc<ROWS>(r, s, k_valid); s.p += 32 * s.ks; }
    __device__ __forceinline__ void store(const float4 (&r)[ROWS / 32], uint32_t stage, int) { sstore_rc<ROWS>(r, s, stage); }
};
template <int ROWS> struct Ld<kModeKCV, ROWS> {
    LState s;
    __device__ __forceinline__ Ld(const Operand& o, int tid) : s(lstate_kcv<ROWS>(o, tid)) {}
    __device__ __forceinline__ void load(float4 (&r)[ROWS / 32], int k_valid, int) { gload_kcv<ROWS>(r, s, k_valid); s.p += 32; }
    __device__ __forceinline__ void store(const float4 (&r)[ROWS / 32], uint32_t stage, int) { sstore_kcv<ROWS>(r, s, stage); }
};
template <int ROWS> struct Ld<kModeGen, ROWS> {
    Operand o;
    __device__ __forceinline__ Ld(const Operand& o_, int) : o(o_) {}
    __device__ __forceinline__ void load(float4 (&r)[ROWS / 32], int k_valid, int tid) { gload<ROWS / 32, false>(r, o, k_valid, tid); o.base += 32 * o.ks; }
    __device__ __forceinline__ void store(const float4 (&r)[ROWS / 32], uint32_t stage, int tid) { sstore<ROWS / 32>(r, o, stage, tid); }
};

// The contraction loop of one output tile: two shared-memory stages; the loads of block kb + 1 are in flight while block kb is
// issued.  Returns false when a barrier wait timed out.
template <int AM, int BM, int BROWS>
__device__ __forceinline__ bool k_loop(const Operand& A, const Operand& B, int k_len, int nkb, uint32_t smem, uint32_t stage_bytes, NrShared* sh,
                                       uint32_t idesc, bool want_dbias, float (&bsum)[4], int tid) {
    bool ok = true;
    float4 ra[4], rb[BROWS / 32];
    Ld<AM, 128> la(A, tid);
    Ld<BM, BROWS> lb(B, tid);
    la.load(ra, k_len, tid);
    lb.load(rb, k_len, tid);
    for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t s = kb & 1;
        const uint32_t stA = smem + s * stage_bytes, stB = stA + kNrStageA;
        if (kb >= 2 && ok) ok = mbar_wait(smem_u32(&sh->stage_free[s]), ((kb >> 1) - 1) & 1);
        if (want_dbias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) bsum[j] += (ra[j].x + ra[j].y) + (ra[j].z + ra[j].w);
        }
        la.store(ra, stA, tid);
        lb.store(rb, stB, tid);
        if (kb + 1 < nkb) {                                   // next block's loads fly while this one is issued
            la.load(ra, k_len - (kb + 1) * 32, tid);
            lb.load(rb, k_len - (kb + 1) * 32, tid);
        }
        fence_async_smem();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(smem_u32(&sh->stage_full[s]));      // one arrival per loader warp; the MMA warp does the rest
    }
    return ok;
}

// The ninth warp: waits for a stage to be full, issues its four K = 8 MMAs, and hands the stage back when they retire.  With the
// issue on its own thread the loaders never wait for it (before, thread 0 did both and its 0.4 us of descriptor building + issue
// per block sat on everybody's critical path: 64-block weight-gradient CTA, measured 67 us = 18 wait-free + 14 cp.async issue +
// 6 row sums + 28 MMA issue).
__device__ __forceinline__ bool mma_loop(int nkb, uint32_t smem, int n_stages, uint32_t stage_bytes, NrShared* sh, uint32_t idesc, int lane) {
    bool ok = true;
    const uint32_t tmem_base = sh->tmem_base;
    for (int kb = 0; kb < nkb && ok; ++kb) {
        const int s = kb % n_stages;
        ok = mbar_wait(smem_u32(&sh->stage_full[s]), (kb / n_stages) & 1);
        tc_fence_after_sync();
        if (lane == 0 && ok) {
            const uint32_t a_lo = desc_lo(smem + s * stage_bytes, 16), b_lo = a_lo + (kNrStageA >> 4);
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks) umma_tf32_lohi(tmem_base, a_lo + 2 * ks, b_lo + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
            umma_commit(smem_u32(&sh->stage_free[s]));
            if (kb == nkb - 1) umma_commit(smem_u32(&sh->done));
        }
        __syncwarp();
    }
    return ok;
}

// ---- asynchronous contraction loop of the weight-gradient problems (both operands k-contiguous NCHW planes): cp.async
// (LDGSTS, 16 bytes) writes every chunk straight to its place in the K-major SWIZZLE_128B tile, no registers in between, so LA
// K blocks are in flight per CTA behind a ring of 2-4 stages (a weight-gradient CTA walks 16-64 K blocks; through registers
// every block cost a full global-memory round trip).  Values are not rounded on the way: the tensor core truncates to tf32.
// Chunk map as in the KCV loader: rows tid / 8 + 32 j, chunk tid % 8.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int bytes = valid ? 16 : 0;                           // 0 = fill the 16 bytes with zeros, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NJ>
__device__ __forceinline__ void issue_kmajor(const Operand& o, int kb, int k_len, uint32_t stage, int tid) {
    const int hi = tid >> 3, c = tid & 7;
    const bool kv = kb * 32 + 4 * c < k_len;
    const float* src = o.base + hi * o.rs + kb * 32 + 4 * c;
    const uint32_t dst = stage + (hi >> 3) * 1024 + (hi & 7) * 128 + ((c ^ (hi & 7)) << 4);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int row = hi + 32 * j;
        if (row < o.rows_tile) cp_async16(dst + j * 4096, src + j * 32 * o.rs, kv && row < o.rows_valid);
    }
}
template <int LA, int BROWS>                                    // LA = K blocks in flight beyond the one being consumed
__device__ __forceinline__ bool k_loop_async(const Operand& A, const Operand& B, int k_len, int nkb, uint32_t smem, int n_stages, uint32_t stage_bytes,
                                             NrShared* sh, uint32_t idesc, bool want_dbias, float (&bsum)[4], int tid) {
    bool ok = true;
#pragma unroll
    for (int b = 0; b < LA; ++b) {
        if (b < nkb) {
            issue_kmajor<4>(A, b, k_len, smem + b * stage_bytes, tid);
            issue_kmajor<BROWS / 32>(B, b, k_len, smem + b * stage_bytes + kNrStageA, tid);
        }
        cp_async_commit();
    }
    for (int kb = 0; kb < nkb; ++kb) {
        const int nx = kb + LA;
        if (nx < nkb) {
            const int sn = nx % n_stages;
            if (nx >= n_stages && ok) ok = mbar_wait(smem_u32(&sh->stage_free[sn]), ((nx / n_stages) - 1) & 1);   // its previous tenant's MMAs retired
            issue_kmajor<4>(A, nx, k_len, smem + sn * stage_bytes, tid);
            issue_kmajor<BROWS / 32>(B, nx, k_len, smem + sn * stage_bytes + kNrStageA, tid);
        }
        cp_async_commit();
        cp_async_wait<LA>();                                    // this thread's chunks of block kb have landed
        const int s = kb % n_stages;
        const uint32_t stA = smem + s * stage_bytes, stB = stA + kNrStageA;
        if (want_dbias) {                                       // row sums of this thread's own chunks (rows tid / 8 + 32 j)
            const int hi = tid >> 3, c = tid & 7;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint4 v = ld_shared_v4(stA + ((hi >> 3) + 4 * j) * 1024 + (hi & 7) * 128 + ((c ^ (hi & 7)) << 4));
                bsum[j] += (__uint_as_float(v.x) + __uint_as_float(v.y)) + (__uint_as_float(v.z) + __uint_as_float(v.w));
            }
        }
        fence_async_smem();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(smem_u32(&sh->stage_full[s]));
    }
    return ok;
}

// Epilogue of a pixel-row tile for one 32-column piece, specialised at compile time (a generic version with run-time null checks
// compiled into one dependent load -> use -> store chain per element, ~300 cycles each): the per-column vectors (bias, RGB-head
// weights) were staged in shared memory while the contraction ran; the per-element operand (saved activation for LeakyReLU',
// or the addend) is loaded for all 32 columns before the accumulator is waited for.
enum { kAuxNone = 0, kAuxMask = 1, kAuxAdd = 2 };
constexpr int kPiece = 16;                                      // accumulator columns per epilogue piece (a 32-column tile keeps both warp groups busy)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
template <bool BIAS, bool LRELU, int AUX, bool RGB>
__device__ __forceinline__ void epi_piece(const uint32_t (&v)[kPiece], const float (&aux)[kPiece], const NrShared& sh, float* out, int col0, int nv, bool row_ok,
                                          long long plane, float (&rgb)[3]) {
#pragma unroll
    for (int c = 0; c < kPiece; ++c) {
        float y = __uint_as_float(v[c]);
        if (BIAS) y += sh.bias_s[col0 + c];
        if (LRELU) y = fmaxf(y, y * kSlope);
        if (AUX == kAuxMask) y = aux[c] > 0.f ? y : y * kSlope;
        if (AUX == kAuxAdd) y += aux[c];
        if (row_ok && c < nv) out[c * plane] = y;
        if (RGB) {                                              // columns beyond N carry zero weights
            rgb[0] = fmaf(sh.wrgb_s[0][col0 + c], y, rgb[0]);
            rgb[1] = fmaf(sh.wrgb_s[1][col0 + c], y, rgb[1]);
            rgb[2] = fmaf(sh.wrgb_s[2][col0 + c], y, rgb[2]);
        }
    }
}
__device__ __forceinline__ void load_aux(float (&aux)[kPiece], const float* src, int nv, bool row_ok, long long plane) {
#pragma unroll
    for (int c = 0; c < kPiece; ++c) aux[c] = (row_ok && c < nv) ? __ldg(src + c * plane) : 0.f;
}

// Epilogue of one pixel-row tile (all threads of the CTA): warp w drains TMEM lanes (w % 4) * 32 .. + 31, the two warp groups take
// alternate 16-column pieces; the RGB head's two partial dot products meet in shared memory.
__device__ __forceinline__ void epilogue_rows(const NrProb& P, NrShared& sh, uint32_t tmem_acc, int item, int m0, int n0, int n_tile,
                                              const float (&rgb_prev)[3], bool ok, int warp, int lane) {
    const int q = warp & 3, half = warp >> 2;
    const int m = m0 + q * 32 + lane;
    const uint32_t lane_addr = tmem_acc + ((uint32_t)(q * 32) << 16);
    const int n_pieces = (n_tile + kPiece - 1) / kPiece;
    float rgb[3] = {0.f, 0.f, 0.f};
    const long long plane = P.M;
    const float* wrgb = P.wrgb;
    const bool row_ok = m < P.M;
    const long long idx0 = item * P.out_item + m;
    const int epi = P.epi;
    const float* aux_src = epi == 4 ? P.mask_act : (epi == 5 ? P.addend : nullptr);
    for (int pc = half; pc < n_pieces && ok; pc += 2) {
        uint32_t v[kPiece];
        float aux[kPiece];
        const int n_first = n0 + pc * kPiece, nv = P.N - n_first;
        const long long idx_first = idx0 + n_first * plane;
        tmem_ld16(lane_addr + pc * kPiece, v);
        if (aux_src) load_aux(aux, aux_src + idx_first, nv, row_ok, plane);
        tmem_ld_wait();
        float* out = P.out + idx_first;
        switch (epi) {                                        // block-uniform
            case 0: epi_piece<false, false, kAuxNone, false>(v, aux, sh, out, pc * kPiece, nv, row_ok, plane, rgb); break;
            case 1: epi_piece<true, false, kAuxNone, false>(v, aux, sh, out, pc * kPiece, nv, row_ok, plane, rgb); break;
            case 2: epi_piece<true, true, kAuxNone, false>(v, aux, sh, out, pc * kPiece, nv, row_ok, plane, rgb); break;
            case 3: epi_piece<true, true, kAuxNone, true>(v, aux, sh, out, pc * kPiece, nv, row_ok, plane, rgb); break;
            case 4: epi_piece<false, false, kAuxMask, false>(v, aux, sh, out, pc * kPiece, nv, row_ok, plane, rgb); break;
            default: epi_piece<false, false, kAuxAdd, false>(v, aux, sh, out, pc * kPiece, nv, row_ok, plane, rgb); break;
        }
    }
    if (wrgb) {                                               // block-uniform branch
        if (half == 1) { sh.rgb_part[(q * 32 + lane) * 3 + 0] = rgb[0]; sh.rgb_part[(q * 32 + lane) * 3 + 1] = rgb[1]; sh.rgb_part[(q * 32 + lane) * 3 + 2] = rgb[2]; }
        asm volatile("bar.sync 1, %0;" ::"n"(kNrThreads) : "memory");      // the eight epilogue warps only
        if (half == 0 && row_ok && ok) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const long long idx = item * 3 * plane + j * plane + m;
                float y = rgb[j] + sh.rgb_part[(q * 32 + lane) * 3 + j] + sh.brgb_s[j] + rgb_prev[j];
                if (P.sigmoid) y = 1.0f / (1.0f + __expf(-y));
                P.rgb_out[idx] = y;
            }
        }
    }
}

// per-column vectors of the epilogue: fetched at CTA start, used after the contraction
__device__ __forceinline__ void stage_columns(const NrProb& P, NrShared& sh, int n0, int tid) {
    const int n = n0 + tid;
    sh.bias_s[tid] = (P.bias && n < P.N) ? __ldg(P.bias + n) : 0.f;
    if (P.wrgb) {
#pragma unroll
        for (int j = 0; j < 3; ++j) sh.wrgb_s[j][tid] = n < P.N ? __ldg(P.wrgb + j * P.N + n) : 0.f;
        if (tid < 3) sh.brgb_s[tid] = __ldg(P.brgb + tid);
    }
}
__device__ __forceinline__ void load_rgb_prev(float (&r)[3], const NrProb& P, int item, int m, bool active) {
#pragma unroll
    for (int j = 0; j < 3; ++j) r[j] = (active && P.rgb_in && m < P.M) ? __ldg(P.rgb_in + item * 3ll * P.M + (long long)j * P.M + m) : 0.f;
}

// BROWS = rows of the B tile the loaders are compiled for: launches whose widest tile has <= 128 columns (every layer from the
// second block on) run the 128-row instantiation - 16 registers fewer per thread, three CTAs per SM instead of two (ncu: the
// kernel sits at 14-27 % of the warp slots, bound by latency, not by any pipe).
template <int BROWS, int MIN_CTAS>
__global__ void __launch_bounds__(kNrCtaThreads, MIN_CTAS) nr_gemm_kernel(const __grid_constant__ NrLaunch L) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ NrShared sh;
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int pi = 0;
#pragma unroll
    for (int i = 1; i < kNrMaxProblems; ++i)
        if (i < L.n && (int)blockIdx.x >= L.p[i].tile0) pi = i;
    const NrProb& P = L.p[pi];

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { mbar_init(smem_u32(&sh.stage_free[i]), 1); mbar_init(smem_u32(&sh.stage_full[i]), kNrThreads / 32); }
        mbar_init(smem_u32(&sh.done), 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh.tmem_base)), "r"(L.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before_sync();
    __syncthreads();                                           // barriers and the TMEM address are visible to all nine warps
    tc_fence_after_sync();
    bool ok = true;

    {
        // ---- one output tile per CTA
        int t = (int)blockIdx.x - P.tile0;
        const int mt = t % P.m_tiles; t /= P.m_tiles;
        const int nt = t % P.n_tiles; t /= P.n_tiles;
        int item, k0;
        if (P.kind == 0) { item = t; k0 = 0; }
        else { item = t / P.k_chunks; k0 = (t % P.k_chunks) * P.k_chunk; }
        const int m0 = mt * 128, n0 = nt * P.n_tile;
        const int k_len = min(P.k_chunk, P.K - k0);
        const int nkb = (k_len + 31) / 32;
        const int n_tile = min(P.n_tile, ((P.N - n0) + 15) & ~15);

        Operand A, B;
        A.base = P.a + item * P.a_item + (long long)m0 * P.a_rs + (long long)k0 * P.a_ks;
        A.rs = P.a_rs; A.ks = P.a_ks; A.rows_valid = P.M - m0; A.rows_tile = 128;
        A.vec = P.a_ks == 1 && (k_len & 3) == 0 && (P.a_rs & 3) == 0 && ((reinterpret_cast<uintptr_t>(A.base) & 15) == 0);
        B.base = P.b + item * P.b_item + (long long)n0 * P.b_rs + (long long)k0 * P.b_ks;
        B.rs = P.b_rs; B.ks = P.b_ks; B.rows_valid = P.N - n0; B.rows_tile = n_tile;
        B.vec = P.b_ks == 1 && (k_len & 3) == 0 && (P.b_rs & 3) == 0 && ((reinterpret_cast<uintptr_t>(B.base) & 15) == 0);

        const int q = warp & 3, half = (warp >> 2) & 1;
        const int m = m0 + q * 32 + lane;
        float rgb_prev[3] = {0.f, 0.f, 0.f};
        const uint32_t idesc = umma_idesc(128, (uint32_t)n_tile, 2u, 2u, 0, 0);      // 2 = tf32 operands, f32 accumulator
        const bool want_dbias = P.kind == 1 && P.dbias != nullptr && nt == 0;
        float bsum[4] = {0.f, 0.f, 0.f, 0.f};
        const bool a_rc = P.a_rs == 1 && (P.a_ks & 3) == 0 && ((reinterpret_cast<uintptr_t>(A.base) & 15) == 0) && (A.rows_valid >= 128 || (A.rows_valid & 3) == 0);
        const bool b_rc = P.b_rs == 1 && (P.b_ks & 3) == 0 && ((reinterpret_cast<uintptr_t>(B.base) & 15) == 0) && (B.rows_valid >= n_tile || (B.rows_valid & 3) == 0);
        const bool use_async = !(a_rc && (B.vec || b_rc)) && A.vec && B.vec && L.async_loads && nkb >= 4;
        const uint32_t sb_async = kNrStageA + (uint32_t)((n_tile * 128 + 1023) & ~1023);       // this tile's stage: as many as fit the launch's window
        const int n_stages = use_async ? min(4, (int)(L.window_bytes / sb_async)) : 2;
        const uint32_t stage_bytes = use_async ? sb_async : L.stage_bytes;
        if (warp == kNrThreads / 32) {
            ok = mma_loop(nkb, smem, n_stages, stage_bytes, &sh, idesc, lane);
        } else {
            if (P.kind == 0) {
                stage_columns(P, sh, n0, tid);
                if (P.wrgb) load_rgb_prev(rgb_prev, P, item, m, half == 0);
            }
            if (a_rc && B.vec) ok = k_loop<kModeRC, kModeKCV, BROWS>(A, B, k_len, nkb, smem, L.stage_bytes, &sh, idesc, want_dbias, bsum, tid);
            else if (a_rc && b_rc) ok = k_loop<kModeRC, kModeRC, BROWS>(A, B, k_len, nkb, smem, L.stage_bytes, &sh, idesc, want_dbias, bsum, tid);
            else if (use_async) {
                if (n_stages >= 4) ok = k_loop_async<2, BROWS>(A, B, k_len, nkb, smem, n_stages, sb_async, &sh, idesc, want_dbias, bsum, tid);
                else ok = k_loop_async<1, BROWS>(A, B, k_len, nkb, smem, n_stages, sb_async, &sh, idesc, want_dbias, bsum, tid);
            }
            else if (A.vec && B.vec) ok = k_loop<kModeKCV, kModeKCV, BROWS>(A, B, k_len, nkb, smem, L.stage_bytes, &sh, idesc, want_dbias, bsum, tid);
            else ok = k_loop<kModeGen, kModeGen, BROWS>(A, B, k_len, nkb, smem, L.stage_bytes, &sh, idesc, want_dbias, bsum, tid);
        }
        if (warp == kNrThreads / 32) {
            if (!ok && lane == 0) atomicCAS(L.status, 0, 803);
        } else {
        if (ok) ok = mbar_wait(smem_u32(&sh.done), 0);
        tc_fence_after_sync();
        if (!ok && tid == 0) atomicCAS(L.status, 0, 801);

        const uint32_t tmem_base = sh.tmem_base;
        if (P.kind == 0) {
            epilogue_rows(P, sh, tmem_base, item, m0, n0, n_tile, rgb_prev, ok, warp, lane);
        } else {
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
            const int n_pieces = (n_tile + 31) / 32;
            float* out = P.out + (long long)m * P.out_ld;
            const int N = P.N;
            const bool row_ok = m < P.M;
            for (int pc = half; pc < n_pieces && ok; pc += 2) {
                uint32_t v[32];
                tmem_ld32(lane_addr + pc * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int n = n0 + pc * 32 + c;
                    if (row_ok && n < N) atomicAdd(out + n, __uint_as_float(v[c]));
                }
            }
            if (want_dbias && ok) {                           // k-contiguous A: thread = (row tid/8 + 32 j, chunk tid % 8)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float sum = bsum[j];
                    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
                    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
                    const int row = (tid >> 3) + 32 * j;
                    if ((tid & 7) == 0 && m0 + row < P.M) atomicAdd(P.dbias + m0 + row, sum);
                }
            }
        }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sh.tmem_base), "r"(L.tmem_cols) : "memory");
}

// Gradient entering a block's feat_layers convolution (neural_renderer.py:83-87 backwards): the RGB head adds W_rgb^T g_rgb to
// the gradient arriving from the next block, then LeakyReLU'.  For the last block g_rgb itself comes from the image gradient
// through the sigmoid.  One thread per 4 pixels and group of 8 channels (coalesced across the warp).
constexpr int kHeadChannels = 8;
template <int V>                                                // V = pixels per thread (4: float4 along the plane, needs P % 4 == 0)
__global__ void __launch_bounds__(256) nr_head_bwd_kernel(const float* __restrict__ g_img, const float* __restrict__ img, int sigmoid,
                                                          float* __restrict__ g_rgb, const float* __restrict__ g_net, const float* __restrict__ net,
                                                          const float* __restrict__ wrgb, float* __restrict__ g_pre, int Cn, long long P, int items) {
    const long long t = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * V;
    if (t >= P * items) return;
    const int item = (int)(t / P);
    const long long p = t % P;
    float gr[3][V];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const long long idx = ((long long)item * 3 + j) * P + p;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            if (g_img) {
                float g = __ldg(g_img + idx + e);
                if (sigmoid) { const float s = __ldg(img + idx + e); g *= s * (1.0f - s); }
                if (blockIdx.y == 0) g_rgb[idx + e] = g;
                gr[j][e] = g;
            } else gr[j][e] = g_rgb[idx + e];
        }
    }
    // fixed trip count, fully unrolled: the read-only loads of all 8 channels are issued before the first store (a run-time bounded
    // loop exposed one global round trip per channel: ncu long-scoreboard stall 28 of 31 cycles per issue)
    const int c0 = blockIdx.y * kHeadChannels;
    float a[kHeadChannels][V], gn[kHeadChannels][V];
#pragma unroll
    for (int k = 0; k < kHeadChannels; ++k) {
        const int c = c0 + k;
        const long long idx = ((long long)item * Cn + (c < Cn ? c : Cn - 1)) * P + p;
        if (V == 4) {
            const float4 av = __ldg(reinterpret_cast<const float4*>(net + idx));
            a[k][0] = av.x; a[k][1] = av.y; a[k][2] = av.z; a[k][3] = av.w;
            if (g_net) { const float4 gv = __ldg(reinterpret_cast<const float4*>(g_net + idx)); gn[k][0] = gv.x; gn[k][1] = gv.y; gn[k][2] = gv.z; gn[k][3] = gv.w; }
        } else {
            a[k][0] = __ldg(net + idx);
            if (g_net) gn[k][0] = __ldg(g_net + idx);
        }
    }
#pragma unroll
    for (int k = 0; k < kHeadChannels; ++k) {
        const int c = c0 + k;
        if (c >= Cn) break;
        const long long idx = ((long long)item * Cn + c) * P + p;
        const float w0 = __ldg(wrgb + c), w1 = __ldg(wrgb + Cn + c), w2 = __ldg(wrgb + 2 * Cn + c);
        float o[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            float v = fmaf(w0, gr[0][e], fmaf(w1, gr[1][e], w2 * gr[2][e]));
            if (g_net) v += gn[k][e];
            o[e] = a[k][e] > 0.f ? v : v * kSlope;
        }
        if (V == 4) *reinterpret_cast<float4*>(g_pre + idx) = make_float4(o[0], o[1], o[2], o[3]);
        else g_pre[idx] = o[0];
    }
}

// --------------------------------------------------------------------------------------------------------- host side
struct NrDims {
    int nb;
    int C[HN_NR_MAX_BLOCKS + 1];          // channels entering block i (C[nb] = channels leaving the last one)
    long long P[HN_NR_MAX_BLOCKS + 1];    // pixels per item at level i
    int R[HN_NR_MAX_BLOCKS + 1];
};
static bool nr_dims(int n_blocks, int feat_nc, int min_feat, int fs, NrDims* d) {
    if (n_blocks < 1 || n_blocks > HN_NR_MAX_BLOCKS || feat_nc < 4 || min_feat < 4 || fs < 2) return false;
    d->nb = n_blocks;
    for (int i = 0; i <= n_blocks; ++i) {
        d->C[i] = std::max(feat_nc >> i, min_feat);
        d->R[i] = fs << i;
        d->P[i] = (long long)d->R[i] * d->R[i];
        if (d->C[i] > 256 && i > 0) return false;                   // the RGB head needs the whole row in one accumulator
        if (d->C[i] % 4) return false;
    }
    return true;
}
static long long pad64(long long n) { return (n + 63) & ~63ll; }

// saved-for-backward layout: per block h1, z2, y, net; then the RGB chain
struct NrSaved { long long h1[HN_NR_MAX_BLOCKS], z2[HN_NR_MAX_BLOCKS], y[HN_NR_MAX_BLOCKS], net[HN_NR_MAX_BLOCKS];
                 long long rgbA, rgbU[HN_NR_MAX_BLOCKS + 1], rgbS[HN_NR_MAX_BLOCKS + 1], total; };
static NrSaved nr_saved(const NrDims& d, int B) {
    NrSaved s{};
    long long o = 0;
    for (int i = 0; i < d.nb; ++i) {
        s.h1[i] = o; o += pad64((long long)B * 2 * d.C[i] * d.P[i]);
        s.z2[i] = o; o += pad64((long long)B * 4 * d.C[i] * d.P[i]);
        s.y[i] = o; o += pad64((long long)B * d.C[i] * d.P[i + 1]);
        s.net[i] = o; o += pad64((long long)B * d.C[i + 1] * d.P[i + 1]);
    }
    s.rgbA = o; o += pad64((long long)B * 3 * d.P[0]);
    for (int l = 1; l <= d.nb; ++l) { s.rgbU[l] = o; o += pad64((long long)B * 3 * d.P[l]); s.rgbS[l] = o; o += pad64((long long)B * 3 * d.P[l]); }
    s.total = o;
    return s;
}
// backward scratch: [zeroed: dxs_i (i >= 1), gR_l (l < nb)] then gR_nb, gpre_f, g_y, dz2, gpre1 (each sized for the largest block), dxs_0
struct NrScratch { long long dxs[HN_NR_MAX_BLOCKS], gR[HN_NR_MAX_BLOCKS + 1], zero_floats, gpre_f, g_y, dz2, gpre1, total; };
static NrScratch nr_scratch(const NrDims& d, int B) {
    NrScratch s{};
    long long o = 0;
    for (int i = 1; i < d.nb; ++i) { s.dxs[i] = o; o += pad64((long long)B * d.C[i] * d.P[i]); }
    for (int l = 0; l < d.nb; ++l) { s.gR[l] = o; o += pad64((long long)B * 3 * d.P[l]); }
    s.zero_floats = o;
    s.gR[d.nb] = o; o += pad64((long long)B * 3 * d.P[d.nb]);
    long long f = 0, y = 0, z = 0, g1 = 0;
    for (int i = 0; i < d.nb; ++i) {
        f = std::max(f, (long long)B * d.C[i + 1] * d.P[i + 1]);
        y = std::max(y, (long long)B * d.C[i] * d.P[i + 1]);
        z = std::max(z, (long long)B * 4 * d.C[i] * d.P[i]);
        g1 = std::max(g1, (long long)B * 2 * d.C[i] * d.P[i]);
    }
    s.gpre_f = o; o += pad64(f);
    s.g_y = o; o += pad64(y);
    s.dz2 = o; o += pad64(z);
    s.gpre1 = o; o += pad64(g1);
    s.dxs[0] = o; o += pad64((long long)B * d.C[0] * d.P[0]);
    s.total = o;
    return s;
}

// pixel-row problem: out[item][n][m] = epi(sum_k act[item][k][m] * W(n, k))
static NrProb pixel_rows(const float* act, int K, long long Ppix, int B, const float* W, int N, bool transposed_w, int w_ld, float* out) {
    NrProb p{};
    p.a = act; p.a_item = (long long)K * Ppix; p.a_rs = 1; p.a_ks = (int)Ppix;
    p.b = W; p.b_item = 0;
    if (transposed_w) { p.b_rs = 1; p.b_ks = w_ld; } else { p.b_rs = w_ld; p.b_ks = 1; }
    p.out = out; p.out_item = (long long)N * Ppix;
    p.M = (int)Ppix; p.N = N; p.K = K; p.kind = 0;
    p.m_tiles = (int)((Ppix + 127) / 128);
    p.n_tile = std::min(kNrMaxN, (N + 15) & ~15);
    p.n_tiles = (N + p.n_tile - 1) / p.n_tile;
    p.k_chunk = K; p.k_chunks = 1;
    p.tiles = p.m_tiles * p.n_tiles * B;
    return p;
}
// weight-gradient problem: dW[m][n] += sum_items sum_pixels g[item][m][pix] * x[item][n][pix],  db[m] += sum g
static NrProb weight_grad(const float* g, int M, const float* x, int N, long long Ppix, int B, float* dW, float* db) {
    NrProb p{};
    p.a = g; p.a_item = (long long)M * Ppix; p.a_rs = (int)Ppix; p.a_ks = 1;
    p.b = x; p.b_item = (long long)N * Ppix; p.b_rs = (int)Ppix; p.b_ks = 1;
    p.out = dW; p.out_ld = N; p.dbias = db;
    p.M = M; p.N = N; p.K = (int)Ppix; p.kind = 1;
    p.m_tiles = (M + 127) / 128;
    p.n_tile = std::min(kNrMaxN, (N + 15) & ~15);
    p.n_tiles = (N + p.n_tile - 1) / p.n_tile;
    // split the pixels so that the problem has a few hundred CTAs of at least four K blocks each
    const long long tiles = (long long)p.m_tiles * p.n_tiles;
    long long chunk = 128;
    while (chunk < Ppix && tiles * B * ((Ppix + chunk - 1) / chunk) > 320) chunk *= 2;
    p.k_chunk = (int)std::min(chunk, (Ppix + 31) / 32 * 32);
    p.k_chunks = (int)((Ppix + p.k_chunk - 1) / p.k_chunk);
    p.tiles = (int)(tiles * B * p.k_chunks);
    return p;
}

static bool async_enabled() {
    static const bool on = [] { const char* e = getenv("HN_NR_ASYNC"); return !(e && e[0] == '0'); }();
    return on;
}
static int launch_group(std::vector<NrProb>& ps, int* status, cudaStream_t st) {
    static const bool split = [] { const char* e = getenv("HN_NR_SPLIT"); return e && e[0] == '1'; }();      // diagnostic: one launch per problem
    if (split && ps.size() > 1) {
        std::vector<NrProb> all;
        all.swap(ps);
        for (auto& p : all) {
            std::vector<NrProb> one{p};
            if (int rc = launch_group(one, status, st)) return rc;
        }
        return 0;
    }
    NrLaunch L{};
    int total = 0;
    L.n = 0;
    for (auto& p : ps) {
        if (p.tiles <= 0) continue;
        if (L.n == kNrMaxProblems) return set_error(HN_E_UNSUPPORTED, "hn_nr: too many problems in one launch");
        p.epi = p.wrgb ? 3 : (p.bias ? (p.lrelu ? 2 : 1) : (p.mask_act ? 4 : (p.addend ? 5 : 0)));
        p.tile0 = total;
        total += p.tiles;
        L.p[L.n++] = p;
    }
    ps.clear();
    if (!L.n) return 0;
    L.status = status;
    L.async_loads = async_enabled() ? 1 : 0;
    int widest = 16;
    for (int i = 0; i < L.n; ++i) widest = std::max(widest, L.p[i].n_tile);
    L.tmem_cols = widest <= 32 ? 32 : (widest <= 64 ? 64 : (widest <= 128 ? 128 : 256));
    L.stage_bytes = kNrStageA + (uint32_t)((widest * 128 + 1023) & ~1023);
    const bool small = widest <= 128;                         // the 128-row instantiation: three CTAs per SM, so at most 64 KiB of stages each
    L.window_bytes = 2 * L.stage_bytes;                       // two stages of the widest tile; four of the widest weight-gradient tile if that fits
    for (int i = 0; i < L.n; ++i)
        if (L.p[i].kind == 1 && L.async_loads) {
            const uint32_t sb = kNrStageA + (uint32_t)((L.p[i].n_tile * 128 + 1023) & ~1023);
            L.window_bytes = std::max(L.window_bytes, std::min(4 * sb, small ? 64u * 1024u : 2 * kNrStage));
        }
    if (small) nr_gemm_kernel<128, 3><<<total, kNrCtaThreads, L.window_bytes + 1024, st>>>(L);
    else nr_gemm_kernel<256, 2><<<total, kNrCtaThreads, L.window_bytes + 1024, st>>>(L);
    return check_launch("hn_nr (grouped tf32 GEMM)");
}

static int nr_prepare() {
    int dev = 0;
    cudaGetDevice(&dev);
    static bool ready[64] = {};
    if (dev < 64 && !ready[dev]) {
        cudaError_t e = cudaFuncSetAttribute(nr_gemm_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNrSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nr_gemm_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kNrSmem);
        if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
        ready[dev] = true;
    }
    return 0;
}

static bool nr_check(const hn_nr_fwd_t* a, NrDims* d, const char** why) {
    if (!a || !a->x || !a->saved || !a->img || !a->status || a->B <= 0) { *why = "null pointer or empty batch"; return false; }
    if (!nr_dims(a->n_blocks, a->feat_nc, a->min_feat, a->featmap_size, d)) { *why = "unsupported geometry (1..4 blocks, channels a multiple of 4, <= 256 after the first block)"; return false; }
    for (int i = 0; i < d->nb; ++i)
        if (!a->w1[i] || !a->b1[i] || !a->w2[i] || !a->b2[i] || !a->wf[i] || !a->bf[i]) { *why = "null weight pointer"; return false; }
    for (int j = 0; j <= d->nb; ++j)
        if (!a->wrgb[j] || !a->brgb[j]) { *why = "null RGB-head pointer"; return false; }
    return true;
}

}  // namespace hn

extern "C" long long hn_nr_saved_floats(int B, int n_blocks, int feat_nc, int min_feat, int featmap_size) {
    hn::NrDims d;
    if (B <= 0 || !hn::nr_dims(n_blocks, feat_nc, min_feat, featmap_size, &d)) return -1;
    return hn::nr_saved(d, B).total;
}
extern "C" long long hn_nr_scratch_floats(int B, int n_blocks, int feat_nc, int min_feat, int featmap_size) {
    hn::NrDims d;
    if (B <= 0 || !hn::nr_dims(n_blocks, feat_nc, min_feat, featmap_size, &d)) return -1;
    return hn::nr_scratch(d, B).total;
}
extern "C" int hn_nr_launches(int n_blocks, int backward) {
    return backward ? 6 * n_blocks + 1 : 5 * n_blocks;      // kernels (the two memsets of the backward call are not counted)
}

extern "C" int hn_nr_fwd(const hn_nr_fwd_t* a, void* stream) {
    using namespace hn;
    NrDims d;
    const char* why = "";
    if (!nr_check(a, &d, &why)) return set_error(HN_E_BADARG, why);
    if (int rc = nr_prepare()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const NrSaved S = nr_saved(d, a->B);
    float* sv = a->saved;
    const int B = a->B, nb = d.nb;
    std::vector<NrProb> g;

    // rgb = rgb_upsample(feat_2_rgb_list[0](x))  and  layer_1 of block 0 read the same planes: one launch
    {
        NrProb p = pixel_rows(a->x, d.C[0], d.P[0], B, a->wrgb[0], 3, false, d.C[0], sv + S.rgbA);
        p.bias = a->brgb[0];
        g.push_back(p);
    }
    const float* x_in = a->x;
    for (int i = 0; i < nb; ++i) {
        const int C = d.C[i], Cn = d.C[i + 1];
        NrProb p1 = pixel_rows(x_in, C, d.P[i], B, a->w1[i], 2 * C, false, C, sv + S.h1[i]);
        p1.bias = a->b1[i]; p1.lrelu = 1;
        g.push_back(p1);
        if (int rc = launch_group(g, a->status, st)) return rc;
        if (i == 0)
            if (int rc = hn_rgb_upsample_fwd(sv + S.rgbA, a->rgb_taps, sv + S.rgbU[1], B * 3, d.R[0], d.R[0], stream)) return rc;
        NrProb p2 = pixel_rows(sv + S.h1[i], 2 * C, d.P[i], B, a->w2[i], 4 * C, false, 2 * C, sv + S.z2[i]);
        p2.bias = a->b2[i];
        g.push_back(p2);
        if (int rc = launch_group(g, a->status, st)) return rc;
        if (int rc = hn_upsample_tail_fwd(sv + S.z2[i], x_in, a->tail_taps[i], sv + S.y[i], B, C, d.R[i], d.R[i], stream)) return rc;
        const bool last = i == nb - 1;
        NrProb pf = pixel_rows(sv + S.y[i], C, d.P[i + 1], B, a->wf[i], Cn, false, C, sv + S.net[i]);
        pf.bias = a->bf[i]; pf.lrelu = 1;
        pf.wrgb = a->wrgb[i + 1]; pf.brgb = a->brgb[i + 1]; pf.rgb_in = sv + S.rgbU[i + 1];
        pf.rgb_out = last ? a->img : sv + S.rgbS[i + 1];
        pf.sigmoid = last && a->final_actvn;
        g.push_back(pf);
        if (int rc = launch_group(g, a->status, st)) return rc;
        if (!last)
            if (int rc = hn_rgb_upsample_fwd(sv + S.rgbS[i + 1], a->rgb_taps, sv + S.rgbU[i + 2], B * 3, d.R[i + 1], d.R[i + 1], stream)) return rc;
        x_in = sv + S.net[i];
    }
    return 0;
}

extern "C" int hn_nr_bwd(const hn_nr_bwd_t* b, void* stream) {
    using namespace hn;
    NrDims d;
    const char* why = "";
    if (!b || !nr_check(&b->f, &d, &why)) return set_error(HN_E_BADARG, b ? why : "null pointer");
    if (!b->g_img || !b->scratch) return set_error(HN_E_BADARG, "hn_nr_bwd: null gradient or scratch pointer");
    if (int rc = nr_prepare()) return rc;
    const hn_nr_fwd_t* a = &b->f;
    cudaStream_t st = (cudaStream_t)stream;
    const NrSaved S = nr_saved(d, a->B);
    const NrScratch T = nr_scratch(d, a->B);
    const float* sv = a->saved;
    float* sc = b->scratch;
    const int B = a->B, nb = d.nb;
    std::vector<NrProb> g;
    float* dxs0 = b->g_x ? b->g_x : sc + T.dxs[0];

    cudaError_t e = cudaMemsetAsync(sc, 0, sizeof(float) * (size_t)T.zero_floats, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(dxs0, 0, sizeof(float) * (size_t)B * d.C[0] * d.P[0], st);
    if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));

    for (int i = nb - 1; i >= 0; --i) {
        const int C = d.C[i], Cn = d.C[i + 1];
        const bool last = i == nb - 1;
        const float* x_in = i ? sv + S.net[i - 1] : a->x;
        float* dxs = i ? sc + T.dxs[i] : dxs0;
        float* gR = sc + T.gR[i + 1];
        {   // gradient entering feat_layers[i]'s pre-activation
            const long long n = d.P[i + 1] * B;
            const dim3 cg((unsigned)((Cn + kHeadChannels - 1) / kHeadChannels));
            const float* gimg = last ? b->g_img : nullptr;
            const float* gnet = last ? nullptr : sc + T.dxs[i + 1];
            const bool v4 = d.P[i + 1] % 4 == 0 && (reinterpret_cast<uintptr_t>(b->g_img) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->img) & 15) == 0;
            if (v4)
                nr_head_bwd_kernel<4><<<dim3((unsigned)((n / 4 + 255) / 256), cg.x), 256, 0, st>>>(gimg, a->img, last && a->final_actvn, gR, gnet, sv + S.net[i],
                                                                                                    a->wrgb[i + 1], sc + T.gpre_f, Cn, d.P[i + 1], B);
            else
                nr_head_bwd_kernel<1><<<dim3((unsigned)((n + 255) / 256), cg.x), 256, 0, st>>>(gimg, a->img, last && a->final_actvn, gR, gnet, sv + S.net[i],
                                                                                                a->wrgb[i + 1], sc + T.gpre_f, Cn, d.P[i + 1], B);
            if (int rc = check_launch("hn_nr_bwd (RGB head)")) return rc;
        }
        g.push_back(pixel_rows(sc + T.gpre_f, Cn, d.P[i + 1], B, a->wf[i], C, true, C, sc + T.g_y));
        if (b->dwf[i]) g.push_back(weight_grad(sc + T.gpre_f, Cn, sv + S.y[i], C, d.P[i + 1], B, b->dwf[i], b->dbf[i]));
        if (b->dwrgb[i + 1]) g.push_back(weight_grad(gR, 3, sv + S.net[i], Cn, d.P[i + 1], B, b->dwrgb[i + 1], b->dbrgb[i + 1]));
        if (int rc = launch_group(g, a->status, st)) return rc;
        if (int rc = hn_upsample_tail_bwd(sc + T.g_y, sv + S.z2[i], a->tail_taps[i], sc + T.dz2, dxs, B, C, d.R[i], d.R[i], stream)) return rc;
        {
            NrProb p = pixel_rows(sc + T.dz2, 4 * C, d.P[i], B, a->w2[i], 2 * C, true, 2 * C, sc + T.gpre1);
            p.mask_act = sv + S.h1[i];
            g.push_back(p);
        }
        if (b->dw2[i]) g.push_back(weight_grad(sc + T.dz2, 4 * C, sv + S.h1[i], 2 * C, d.P[i], B, b->dw2[i], b->db2[i]));
        if (int rc = launch_group(g, a->status, st)) return rc;
        if (i > 0 || b->g_x) {
            NrProb p = pixel_rows(sc + T.gpre1, 2 * C, d.P[i], B, a->w1[i], C, true, C, dxs);
            p.addend = dxs;
            g.push_back(p);
        }
        if (b->dw1[i]) g.push_back(weight_grad(sc + T.gpre1, 2 * C, x_in, C, d.P[i], B, b->dw1[i], b->db1[i]));
        if (int rc = launch_group(g, a->status, st)) return rc;
        if (int rc = hn_rgb_upsample_bwd(gR, a->rgb_taps, sc + T.gR[i], B * 3, d.R[i], d.R[i], stream)) return rc;
    }
    if (b->g_x) {
        NrProb p = pixel_rows(sc + T.gR[0], 3, d.P[0], B, a->wrgb[0], d.C[0], true, d.C[0], b->g_x);
        p.addend = b->g_x;
        g.push_back(p);
    }
    if (b->dwrgb[0]) g.push_back(weight_grad(sc + T.gR[0], 3, a->x, d.C[0], d.P[0], B, b->dwrgb[0], b->dbrgb[0]));
    return launch_group(g, a->status, st);
}
