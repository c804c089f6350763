// probe_mma_rate.cu — measures sustained tcgen05.mma (SS mode, M=128, K=16, fp16) issue/execute rate per SM for
// different N, with and without a concurrent stream of bulk copies into shared memory (weight-ring traffic).
//   usage: probe_mma_rate <N> <fill 0|1> <a_same 0|1>
// prints cycles per MMA (tensor floor: 128*N/256 cycles).  Design aid for DESIGN.md §5; not part of the product.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../nerf-3dtalker-code_b200/csrc/hn_tc.cuh"
using namespace hn;

__global__ void __launch_bounds__(128, 1) rate_kernel(const uint8_t* src, long long* out, int N, int fill, int reps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar_done, bar_fill[4];
    __shared__ uint32_t tmem_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(smem_u32(&bar_done), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar_fill[i]), 1); stop = 0; mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_s));
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tm = tmem_s;
    // layout: A blocks 0..5 (96 KB), B region 64 KB at 96K, fill ring 4 x 16 KB at 160K
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(128, N, kF16, kF16, 0, 0);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t a = smem + (r % 6) * 16384, b = smem + 98304 + (r & 1) * 32768;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
                umma_f16(tm + ((r * N) & 255), umma_desc_kmajor(a, ks), umma_desc_kmajor(b, ks), idesc, true);
        }
        umma_commit(smem_u32(&bar_done));
        const long long t1 = clock64();
        mbar_wait(smem_u32(&bar_done), 0);
        const long long t2 = clock64();
        stop = 1;
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else if (tid == 32 && fill) {
        uint32_t n = 0;
        while (!stop) {
            const uint32_t s = n & 3, par = (n >> 2) & 1;
            if (n >= 4) mbar_wait(smem_u32(&bar_fill[s]), par ^ 1);
            mbar_arrive_expect_tx(smem_u32(&bar_fill[s]), 16384);
            bulk_g2s(smem + 163840 + s * 16384, src + (size_t)((n * 7 + blockIdx.x) % 512) * 16384, 16384, smem_u32(&bar_fill[s]));
            ++n;
        }
        // drain
        for (uint32_t k = (n >= 4 ? n - 4 : 0); k < n; ++k) mbar_wait(smem_u32(&bar_fill[k & 3]), (k >> 2) & 1);
        if (blockIdx.x == 0) out[2] = n;
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_free<512>(tm);
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 128, fill = argc > 2 ? atoi(argv[2]) : 0;
    const int reps = 4096;
    uint8_t* src; long long* out;
    cudaMalloc(&src, 512 * 16384); cudaMemset(src, 0, 512 * 16384);
    cudaMalloc(&out, 64); cudaMemset(out, 0, 64);
    const int smem = 163840 + 65536 + 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int it = 0; it < 2; ++it) rate_kernel<<<148, 128, smem>>>(src, out, N, fill, reps);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8] = {0};
    cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
    printf("N=%d fill=%d: %s issue %.1f cyc/MMA, complete %.1f cyc/MMA (floor %d), fills %lld (%.1f B/cyc)\n", N, fill, cudaGetErrorString(e),
           (double)h[0] / (reps * 4), (double)h[1] / (reps * 4), 128 * N / 256, h[2], h[1] ? (double)h[2] * 16384 / h[1] : 0.0);
    return e != cudaSuccess;
}
