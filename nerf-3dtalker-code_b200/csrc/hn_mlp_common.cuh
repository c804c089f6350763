// hn_mlp_common.cuh — pieces shared by the fused forward / data-gradient kernels: CTA role layout, bounded
// barrier waits, and the lean epilogue primitives (TMEM row slice -> fp16 operand-image row slice).
//
// CTA layout (640 threads): warp 0 weight producer, warp 1 MMA issuer, warp 2 TMEM allocator, warp 3 idle,
// warps 4..19 epilogue.  An accumulator chunk is 128 rows (TMEM lanes) x up to 128 columns; epilogue warp
// e = warp-4 owns lane quarter (e & 3) (the only quarter a warp may read with tcgen05.ld) and column group
// (e >> 2): 32 rows x 32 columns per warp per chunk, so four warps per scheduler overlap each other's
// TMEM-load / bias-load / shared-store latencies.
#pragma once
#include "hn_tc.cuh"

namespace hn {

constexpr int kCtrlWarps = 4;
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;                  // 512
constexpr int kFusedThreads = (kCtrlWarps + kEpiWarps) * 32; // 640

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

// Optional pipeline cycle counters (build with -DHN_PIPE_COUNTERS, read with tools/pipeline_counters.py): the
// forward kernel's CTA 0 writes, as 64-bit cycle counts, status[2..17] (MMA issuer: total, waits by barrier class,
// issue, commit) and status[18..29] (epilogue warp 0: total, wait, TMEM load, math+store, sync, next-tile PE).
#ifdef HN_PIPE_COUNTERS
#define HN_PC_DECL(v, n) long long v[n] = {0}; const long long v##_start = clock64(); long long v##_t0 = v##_start
#define HN_PC_T0(v) do { v##_t0 = clock64(); } while (0)
#define HN_PC_LAP(v, i) do { const long long _t = clock64(); v[i] += _t - v##_t0; v##_t0 = _t; } while (0)
#define HN_PC_FLUSH(v, n, dst, cond) do { if (cond) { v[0] = clock64() - v##_start; long long* _o = reinterpret_cast<long long*>(dst); \
                                          for (int _i = 0; _i < n; ++_i) _o[_i] = v[_i]; } } while (0)
#else
#define HN_PC_DECL(v, n) do {} while (0)
#define HN_PC_T0(v) do {} while (0)
#define HN_PC_LAP(v, i) do {} while (0)
#define HN_PC_FLUSH(v, n, dst, cond) do {} while (0)
#endif

// Hot-path wait: bounded spin on try_wait only (no clock reads, no flag polling inside the loop).  A try_wait that fails
// returns after a hardware-defined time of the order of a microsecond, so the bound is several seconds: a protocol bug
// surfaces as a status code instead of a hung GPU, and a healthy run pays two instructions per wait.
__device__ __forceinline__ bool wait_spin(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* status, int code) {
    const uint32_t b = smem_u32(bar);
#pragma unroll 1
    for (uint32_t i = 0; i < (1u << 24); ++i) {
        if (mbar_try_wait(b, parity)) return true;
        if ((i & 1023u) == 1023u && *abort_flag) return false;
    }
    *abort_flag = 1;
    atomicCAS(status, 0, code);
    return false;
}

// Wait of a role that expects to wait long (epilogue warps for an accumulator, producers for a free ring stage, the saver):
// polls with a pause.  Every mbarrier operation of the CTA goes through one synchronisation unit; a dozen warps polling
// back-to-back keep it saturated and every arrive / try_wait of the MMA issuer queues behind them.
__device__ __forceinline__ bool wait_backoff(uint64_t* bar, uint32_t parity, volatile int* abort_flag, int* status, int code) {
    const uint32_t b = smem_u32(bar);
#pragma unroll 1
    for (uint32_t i = 0; i < (1u << 22); ++i) {
        if (mbar_try_wait(b, parity)) return true;
        __nanosleep(100);
        if ((i & 255u) == 255u && *abort_flag) return false;
    }
    *abort_flag = 1;
    atomicCAS(status, 0, code);
    return false;
}

__device__ __forceinline__ void named_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one mbarrier arrival per warp: every lane's prior shared-memory / TMEM accesses are ordered before it by the
// warp barrier (each lane has already executed its own proxy / tcgen05 fence).  512 per-thread arrivals on one
// barrier word serialise (~1K cycles); 16 per-warp arrivals do not.
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// two floats -> packed f16x2 (lo = a, hi = b), saturating to +-65504; one F2FP instruction each
__device__ __forceinline__ uint32_t pack_sat(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ uint32_t pack_relu_sat(float a, float b) {     // max(x,0) fused into the conversion
    uint32_t r;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// write 32 consecutive half-precision columns [col0, col0+32) of row `row` into an operand image whose first
// block starts at `base` (blocks are 64 columns wide and kBlockBytes apart)
template <bool RELU>
__device__ __forceinline__ void store_row32(uint32_t base, int row, int col0, const float (&y)[32]) {
    const uint32_t blk = base + (col0 >> 6) * kBlockBytes + (row >> 3) * 1024 + (row & 7) * 128;
    const int ch0 = (col0 & 63) >> 3, rsw = row & 7;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t p[4];
#pragma unroll
        for (int h = 0; h < 4; ++h)
            p[h] = RELU ? pack_relu_sat(y[8 * c + 2 * h], y[8 * c + 2 * h + 1]) : pack_sat(y[8 * c + 2 * h], y[8 * c + 2 * h + 1]);
        st_shared_v4(blk + (((ch0 + c) ^ rsw) << 4), p[0], p[1], p[2], p[3]);
    }
}

// same, from 16 already packed f16 pairs (columns [col0, col0+32))
__device__ __forceinline__ void store_row_packed(uint32_t base, int row, int col0, const uint32_t (&p)[16]) {
    const uint32_t blk = base + (col0 >> 6) * kBlockBytes + (row >> 3) * 1024 + (row & 7) * 128;
    const int ch0 = (col0 & 63) >> 3, rsw = row & 7;
#pragma unroll
    for (int c = 0; c < 4; ++c) st_shared_v4(blk + (((ch0 + c) ^ rsw) << 4), p[4 * c], p[4 * c + 1], p[4 * c + 2], p[4 * c + 3]);
}

// bit i of the result = (y[i] > 0) (sign-bit funnel shifts: one instruction per element; +0.0 counts as positive).
// Four independent 8-deep chains instead of one 32-deep dependency chain.
__device__ __forceinline__ uint32_t positive_mask32(const float (&y)[32]) {
    uint32_t m[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) m[c] = __funnelshift_l(__float_as_uint(y[8 * c + i]), m[c], 1);
    }
    // m[c] holds the signs of y[8c..8c+7] in its low byte, element 8c at bit 7
    const uint32_t packed = (m[0] << 24) | ((m[1] & 0xFFu) << 16) | ((m[2] & 0xFFu) << 8) | (m[3] & 0xFFu);
    return ~__brev(packed);
}

}  // namespace hn
