// probe_issue.cu — what does the MMA-issuing warp pay per batch of 8 MMAs for the bookkeeping around them?
// (design aid, not part of the product).   probe_issue <extras bitmask>
//   bit0: tcgen05.commit to an mbarrier per batch      bit1: tcgen05.fence::after_thread_sync per batch
//   bit2: try_wait on a completed barrier, result used next batch   bit3: indexed constant-table read (8 B) per batch
//   bit4: second commit per batch                      bit5: blocking try_wait (result used at once)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../nerf-3dtalker-code_b200/csrc/hn_tc.cuh"
using namespace hn;
__constant__ unsigned long long c_tab[256];

__global__ void __launch_bounds__(640, 1) k(long long* out, int extras, int reps, int pollers, int busy) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = smem_u32(smem_raw);
    __shared__ __align__(8) uint64_t done, sink[4], ready, never;
    __shared__ volatile int stop;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(smem_u32(&done), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&sink[i]), 1); mbar_init(smem_u32(&ready), 1); mbar_init(smem_u32(&never), 1); stop = 0; mbar_fence_init(); }
    if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_s));
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    if (tid == 0) mbar_arrive(smem_u32(&ready));              // phase 0 of `ready` is complete for the whole run
    __syncthreads();
    const uint32_t tm = tmem_s;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc(128, 128, kF16, kF16, 0, 0);
        bool pre = true;
        unsigned long long t = c_tab[0];
        uint32_t sum = 0;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            unsigned long long nxt = t;
            if (extras & 8) nxt = c_tab[(r * 7 + (int)(t & 3)) & 255];
            if (extras & 32) { while (!mbar_try_wait(smem_u32(&ready), 0)) {} }
            if ((extras & 4) && !pre) { while (!mbar_try_wait(smem_u32(&ready), 0)) {} }
            if (extras & 2) tc_fence_after_sync();
            const uint32_t b_lo = desc_lo(smem + (r & 3) * 16384, 16);
            const uint32_t d = tm + 256 + ((uint32_t)(t & 1) * 128);
            const uint32_t a_t = tm + (r & 3) * 32;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) umma_f16_ts_lo(d, a_t + (ks & 3) * 8, b_lo + (ks & 3) * 2, idesc, 1u);
                if (extras & 1) umma_commit(smem_u32(&sink[r & 3]));
                if (extras & 16) umma_commit(smem_u32(&sink[(r + 1) & 3]));
            }
            __syncwarp();
            if (extras & 4) pre = mbar_try_wait(smem_u32(&ready), 0);
            sum += (uint32_t)t;
            t = nxt;
        }
        if (elect_one()) umma_commit(smem_u32(&done));
        __syncwarp();
        mbar_wait(smem_u32(&done), 0);
        if (lane == 0) { out[blockIdx.x * 2] = clock64() - t0; out[blockIdx.x * 2 + 1] = sum; }
        stop = 1;
    } else if (warp >= 4 && warp < 4 + pollers) {
        // waiting warps as in the real kernel: spin on a barrier that does not complete
        while (!stop) { if (mbar_try_wait(smem_u32(&never), 0)) break; }
    } else if (warp >= 4 + pollers && warp < 4 + pollers + busy) {
        // busy ALU warps (epilogue-like arithmetic)
        float x = (float)tid; uint32_t m = 0;
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { x = x * 1.0001f + 0.5f; m = __funnelshift_l(__float_as_uint(x), m, 1); }
        }
        if (x == 1.2345f && m == 77) out[0] = 0;
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 1) tmem_free<512>(tm);
}
int main(int argc, char** argv) {
    const int extras = argc > 1 ? atoi(argv[1]) : 0, reps = 2000, pollers = argc > 2 ? atoi(argv[2]) : 0, busy = argc > 3 ? atoi(argv[3]) : 0;
    unsigned long long tab[256]; for (int i = 0; i < 256; ++i) tab[i] = i * 2654435761ull;
    cudaMemcpyToSymbol(c_tab, tab, sizeof(tab));
    long long* out; cudaMalloc(&out, 148 * 16); cudaMemset(out, 0, 148 * 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
    for (int it = 0; it < 2; ++it) k<<<148, 640, 66560>>>(out, extras, reps, pollers, busy);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("extras=%2d pollers=%d busy=%d: %s | %.1f cycles per batch of 8 MMAs (tensor time %d)\n", extras, pollers, busy, cudaGetErrorString(e), (double)h[0] / reps, 8 * 74);
    return e != cudaSuccess;
}
