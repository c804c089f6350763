// probe_tf32.cu — hardware probe: tcgen05.mma kind::tf32 with the A operand in MN-major (row-contiguous) SWIZZLE_128B tiles.
// One CTA computes D[128 x 64] = A[128 x 32] * B[64 x 32]^T (B K-major, the validated layout of csrc/hn_nr.cu) for several
// hypotheses about the MN-major tile layout / descriptor fields and prints the error of each against the host product.
//   hypothesis = (group_stride, katom_stride, swizzle on/off, LBO, SBO, K-step advance):
//     element (row m, k) of A lives at  (m / 32) * group_stride + (k / 8) * katom_stride + (k % 8) * 128
//                                       + ((((m % 32) / 4) ^ (swz ? k % 8 : 0)) * 16) + (m % 4) * 4
// Not part of the product; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/probe_tf32 tools/probe_tf32.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../nerf-3dtalker-code_b200/csrc/hn_tc.cuh"

using namespace hn;

struct Hyp { int a_mn, group_stride, katom_stride, swz, lbo, sbo, kstep; const char* name; int b_mn; };
struct ProbeArgs { const uint8_t* a_img; const uint8_t* b_img; float* d; int* status; Hyp h; };

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sA = smem;                 // 16 KiB
    uint8_t* sB = smem + 16384;         // 8 KiB
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<64>(smem_u32(&tmem_base_s));
    for (int i = tid * 16; i < 16384; i += 128 * 16) *reinterpret_cast<uint4*>(sA + i) = *reinterpret_cast<const uint4*>(p.a_img + i);
    for (int i = tid * 16; i < 8192; i += 128 * 16) *reinterpret_cast<uint4*>(sB + i) = *reinterpret_cast<const uint4*>(p.b_img + i);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(128, 64, 2u, 2u, (uint32_t)p.h.a_mn, (uint32_t)p.h.b_mn);
        for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = p.h.a_mn ? umma_desc(smem_u32(sA) + ks * p.h.kstep, p.h.lbo, p.h.sbo) : umma_desc_kmajor(smem_u32(sA), ks);
            const uint64_t bd = p.h.b_mn ? umma_desc(smem_u32(sB) + ks * p.h.kstep, p.h.lbo, p.h.sbo) : umma_desc_kmajor(smem_u32(sB), ks);
            umma_tf32(tmem_base, ad, bd, idesc, ks > 0);
        }
        umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    if (!mbar_wait(smem_u32(&bar), 0)) atomicExch(p.status, 2);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) p.d[(size_t)tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<64>(tmem_base);
}

int main() {
    const int M = 128, N = 64, K = 32;
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    srand(4321);
    for (auto& v : A) v = (float)(rand() % 17 - 8);              // small integers: exact in tf32
    for (auto& v : B) v = (float)(rand() % 17 - 8);
    std::vector<uint8_t> bimg(8192, 0);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        const size_t off = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
        *reinterpret_cast<float*>(&bimg[off]) = B[(size_t)n * K + k];
    }
    const Hyp hyps[] = {
        {0, 0, 0, 1, 0, 0, 0, "control: A K-major"},
        {1, 4096, 1024, 1, 4096, 1024, 1024, "groups outer (4096), k atoms 1024, swizzle, LBO 4096 SBO 1024"},
        {1, 4096, 1024, 1, 1024, 4096, 1024, "same data, LBO 1024 SBO 4096"},
        {1, 1024, 4096, 1, 1024, 4096, 4096, "k atoms outer (4096), groups 1024, swizzle, LBO 1024 SBO 4096"},
        {1, 1024, 4096, 1, 4096, 1024, 4096, "same data, LBO 4096 SBO 1024"},
        {1, 4096, 1024, 0, 4096, 1024, 1024, "groups outer, NO swizzle in the data"},
        {1, 1024, 4096, 0, 1024, 4096, 4096, "k atoms outer, NO swizzle in the data"},
        {0, 4096, 1024, 1, 4096, 1024, 1024, "B MN-major (A K-major): groups outer, LBO 4096 SBO 1024", 1},
        {0, 4096, 1024, 1, 1024, 4096, 1024, "B MN-major: same data, LBO 1024 SBO 4096", 1},
        {0, 1024, 2048, 1, 1024, 2048, 2048, "B MN-major: k atoms outer (2048), LBO 1024 SBO 2048", 1},
        {0, 1024, 2048, 1, 2048, 1024, 2048, "B MN-major: k atoms outer (2048), LBO 2048 SBO 1024", 1},
    };
    uint8_t *da, *db; float* dd; int* ds;
    cudaMalloc(&da, 16384); cudaMalloc(&db, 8192); cudaMalloc(&dd, (size_t)M * N * 4); cudaMalloc(&ds, 4);
    cudaMemcpy(db, bimg.data(), 8192, cudaMemcpyHostToDevice);
    const int smem = 16384 + 8192 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int rc = 1;
    for (const Hyp& h : hyps) {
        std::vector<uint8_t> aimg(16384, 0);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
            size_t off;
            if (!h.a_mn) off = (size_t)(m >> 3) * 1024 + (m & 7) * 128 + (((k >> 2) ^ (m & 7)) << 4) + (k & 3) * 4;
            else off = (size_t)(m / 32) * h.group_stride + (size_t)(k / 8) * h.katom_stride + (k % 8) * 128 +
                       ((((m % 32) / 4) ^ (h.swz ? k % 8 : 0)) << 4) + (m % 4) * 4;
            *reinterpret_cast<float*>(&aimg[off]) = A[(size_t)m * K + k];
        }
        std::vector<uint8_t> bi(bimg);
        if (h.b_mn) {
            std::fill(bi.begin(), bi.end(), 0);
            for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
                const size_t off = (size_t)(n / 32) * h.group_stride + (size_t)(k / 8) * h.katom_stride + (k % 8) * 128 +
                                   ((((n % 32) / 4) ^ (h.swz ? k % 8 : 0)) << 4) + (n % 4) * 4;
                *reinterpret_cast<float*>(&bi[off]) = B[(size_t)n * K + k];
            }
        }
        cudaMemcpy(db, bi.data(), 8192, cudaMemcpyHostToDevice);
        cudaMemcpy(da, aimg.data(), 16384, cudaMemcpyHostToDevice);
        cudaMemset(dd, 0xff, (size_t)M * N * 4); cudaMemset(ds, 0, 4);
        ProbeArgs p{da, db, dd, ds, h};
        probe_kernel<<<1, 128, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s (%s)\n", cudaGetErrorString(e), h.name); return 3; }
        std::vector<float> D((size_t)M * N); int st = 0;
        cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0, zeros = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
            const double err = fabs(ref - D[(size_t)m * N + n]);
            if (!(err <= maxerr)) maxerr = err;
            bad += err > 1e-3;
            zeros += D[(size_t)m * N + n] == 0.f;
        }
        printf("%-72s status=%d maxerr=%.3e wrong=%d/%d zeros=%d D[0][0]=%g D[37][5]=%g -> %s\n", h.name, st, maxerr, bad, M * N, zeros, D[0], D[37 * N + 5],
               (st == 0 && bad == 0) ? "PASS" : "FAIL");
        if (st == 0 && bad == 0 && (h.a_mn || h.b_mn)) rc = 0;
    }
    return rc;
}
