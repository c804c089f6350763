// probe_tmem.cu — how fast can epilogue warps read / write tensor memory while the tensor core is busy?
// (design aid, not part of the product).   probe_tmem <mma 0 none|1 A=SMEM|2 A=TMEM> <epi_warps 4|8|16> <op 0 ld|1 st|2 ld+st>
// One issuing warp streams 128x128x16 MMAs; `epi_warps` warps loop over 32x32b.x32 TMEM loads (and/or x16 stores).
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>
#include <cuda_runtime.h>
#include "../nerf-3dtalker-code_b200/csrc/hn_tc.cuh"
using namespace hn;

__global__ void __launch_bounds__(640, 1) k(long long* out, int mma, int epi_warps, int op, int reps, int mma_reps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(smem_u32(&done), 1); mbar_fence_init(); }
    if (warp == 16) tmem_alloc<512>(smem_u32(&tmem_s));
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tm = tmem_s;
    if (warp == 17 && mma) {
        const uint32_t idesc = umma_idesc(128, 128, kF16, kF16, 0, 0);
        const long long t0 = clock64();
        for (int r = 0; r < mma_reps; ++r) {
            const uint32_t b_lo = desc_lo(smem + (r & 3) * 16384, 16), a_lo = desc_lo(smem + 65536 + (r & 1) * 16384, 16);
            const uint32_t d = tm + 384;                       // accumulator columns 384..511 (never touched by the readers)
            const uint32_t a_t = tm + 256 + (r & 3) * 32;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (mma == 2) umma_f16_ts_lo(d, a_t + ks * 8, b_lo + ks * 2, idesc, 1u);
                    else umma_f16_lohi(d, a_lo + ks * 2, b_lo + ks * 2, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&done));
        __syncwarp();
        mbar_wait(smem_u32(&done), 0);
        if (lane == 0) out[blockIdx.x * 4 + 0] = clock64() - t0;
    } else if (warp < epi_warps) {
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t col = (uint32_t)(warp >> 2) * 32 % 256;  // readers stay inside columns 0..255
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            uint32_t v[32];
            if (op != 1) {
                tmem_ld32(tm + lane_base + col, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += v[i];
            }
            if (op != 0) {
                uint32_t p[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) p[i] = acc + i;
                tmem_st16(tm + lane_base + col, p);
                tmem_st_wait();
            }
        }
        const long long t1 = clock64();
        if (lane == 0) { atomicMax((unsigned long long*)&out[blockIdx.x * 4 + 1], (unsigned long long)(t1 - t0)); out[blockIdx.x * 4 + 2] = acc; }
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 16) tmem_free<512>(tm);
}

int main(int argc, char** argv) {
    const int mma = argc > 1 ? atoi(argv[1]) : 0, epi_warps = argc > 2 ? atoi(argv[2]) : 16, op = argc > 3 ? atoi(argv[3]) : 0;
    const int reps = 2000, mma_reps = 4000;
    long long* out; cudaMalloc(&out, 148 * 4 * 8); cudaMemset(out, 0, 148 * 4 * 8);
    const int smem = 98304 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int it = 0; it < 2; ++it) { cudaMemset(out, 0, 148 * 4 * 8); k<<<148, 640, smem>>>(out, mma, epi_warps, op, reps, mma_reps); }
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(148 * 4);
    cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
    const double bytes = (op == 2 ? 4096.0 + 2048.0 : op == 1 ? 2048.0 : 4096.0) * epi_warps;   // per iteration, all warps
    printf("mma=%d epi_warps=%d op=%d: %s | MMA %.1f cyc each | epilogue loop %.0f cyc/iter, %.1f B/cyc/SM\n", mma, epi_warps, op, cudaGetErrorString(e),
           mma ? (double)h[0] / (mma_reps * 4.0) : 0.0, (double)h[1] / reps, bytes * reps / (double)h[1]);
    return e != cudaSuccess;
}
