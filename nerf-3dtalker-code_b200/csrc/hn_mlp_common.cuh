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

// bit i of the result = (y[i] > 0) (sign-bit funnel shifts: one instruction per element; +0.0 counts as positive)
__device__ __forceinline__ uint32_t positive_mask32(const float (&y)[32]) {
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) m = __funnelshift_l(__float_as_uint(y[i]), m, 1);
    return ~__brev(m);
}

}  // namespace hn
