// probe_epi.cu — cost of one epilogue "piece" (32 rows x 32 accumulator columns: TMEM load, +bias from shared memory,
// relu + f16 pack, optional sign masks, TMEM store) per warp, alone and with all epilogue warps busy (design aid).
//   probe_epi <warps 1..16> <masks 0|1> <mma 0|1: a TS-mode MMA stream runs beside>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../nerf-3dtalker-code_b200/csrc/hn_mlp_common.cuh"
using namespace hn;

__global__ void __launch_bounds__(640, 1) k(long long* out, uint32_t* sink, int warps, int masks, int mma, int reps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(smem_u32(&done), 1); stop = 0; mbar_fence_init(); }
    if (warp == 16) tmem_alloc<512>(smem_u32(&tmem_s));
    for (int i = tid; i < 1024; i += 640) reinterpret_cast<float*>(smem_raw)[65536 / 4 + i] = 0.001f * i;    // bias row at 64 KiB
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tm = tmem_s;
    if (warp == 17 && mma) {
        const uint32_t idesc = umma_idesc(128, 128, kF16, kF16, 0, 0);
        int r = 0;
        while (!stop) {
            const uint32_t b_lo = desc_lo(smem + (r & 3) * 16384, 16);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) umma_f16_ts_lo(tm + 384, tm + 256 + (ks & 3) * 8, b_lo + (ks & 3) * 2, idesc, 1u);
            }
            __syncwarp();
            ++r;
        }
        if (elect_one()) umma_commit(smem_u32(&done));
        __syncwarp();
        mbar_wait(smem_u32(&done), 0);
    } else if (warp < warps) {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t col = (uint32_t)(warp >> 2) * 32;
        const uint32_t bp = smem + 65536 + col * 4;
        uint32_t macc = 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            uint32_t v[32];
            tmem_ld32(tm + lane_base + col, v);
            tmem_ld_wait();
            float y[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float4 bb;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bb.x), "=f"(bb.y), "=f"(bb.z), "=f"(bb.w) : "r"(bp + i * 16));
                y[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + bb.x; y[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + bb.y;
                y[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + bb.z; y[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + bb.w;
            }
            if (masks) macc ^= positive_mask32(y);
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_relu_sat(y[2 * i], y[2 * i + 1]);
            tmem_st16(tm + lane_base + 128 + (col >> 1), pk);
            tmem_st_wait();
        }
        const long long t1 = clock64();
        if (lane == 0) { out[blockIdx.x * 32 + warp] = t1 - t0; }
        sink[blockIdx.x * 640 + tid] = macc;
        __syncwarp();
        if (warp == 0 && lane == 0) { __nanosleep(20000); stop = 1; }
    } else if (warp == 0) {
    }
    if (warps == 0 && tid == 0) stop = 1;
    tc_fence_before_sync(); __syncthreads();
    if (warp == 16) tmem_free<512>(tm);
}
int main(int argc, char** argv) {
    const int warps = argc > 1 ? atoi(argv[1]) : 16, masks = argc > 2 ? atoi(argv[2]) : 0, mma = argc > 3 ? atoi(argv[3]) : 0, reps = 2000;
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 148 * 32 * 8); cudaMemset(out, 0, 148 * 32 * 8); cudaMalloc(&sink, 148 * 640 * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70656);
    for (int it = 0; it < 2; ++it) k<<<148, 640, 70656>>>(out, sink, warps, masks, mma, reps);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(32); cudaMemcpy(h.data(), out, 32 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("warps=%2d masks=%d mma=%d: %s | %.0f cycles per piece per warp -> a 128x128 chunk by %d warps: %.0f cycles\n", warps, masks, mma, cudaGetErrorString(e),
           (double)mx / reps, warps, (double)mx / reps * 16.0 / warps);
    return e != cudaSuccess;
}
