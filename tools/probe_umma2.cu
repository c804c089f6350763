// probe_umma2.cu — hardware probe for the 2-CTA (cta_group::2) tcgen05 path.
// A cluster of two CTAs computes D[256 x N] = A[256 x K] * B[N x K]^T: CTA r holds A rows [128r,128r+128) and
// B rows [N/2*r, N/2*r + N/2) (K-major operand images, same shared-memory offsets in both CTAs); the leader
// (rank 0) issues tcgen05.mma.cta_group::2 with M=256 and commits with a cluster multicast; each CTA reads its
// own 128 accumulator rows from its own TMEM.  Also exercises the remote mbarrier arrive the fused kernels need.
//   usage: probe_umma2 <N> <K>       exit code 0 = match
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "../nerf-3dtalker-code_b200/csrc/hn_tc.cuh"
using namespace hn;
namespace cg = cooperative_groups;

struct Args { const uint8_t* a_img; const uint8_t* b_img; float* d; int* status; int N, K, a_bytes, b_half_bytes; };

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void remote_arrive(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma2_f16(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, bool acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2_kernel(Args p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_ready, bar_done;
    __shared__ uint32_t tmem_s;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t rank = cluster.block_rank();
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t* sA = smem;              // up to 64 KiB : this CTA's 128 rows of A
    uint8_t* sB = smem + 65536;      // up to 64 KiB : this CTA's N/2 rows of B

    if (tid == 0) { mbar_init(smem_u32(&bar_ready), 2); mbar_init(smem_u32(&bar_done), 1); mbar_fence_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    const uint8_t* a_src = p.a_img + (size_t)rank * p.a_bytes;
    const uint8_t* b_src = p.b_img + (size_t)rank * p.b_half_bytes;
    for (int i = tid * 16; i < p.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sA + i) = *reinterpret_cast<const uint4*>(a_src + i);
    for (int i = tid * 16; i < p.b_half_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sB + i) = *reinterpret_cast<const uint4*>(b_src + i);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    cluster.sync();
    tc_fence_after_sync();
    const uint32_t tm = tmem_s;
    // both CTAs tell the leader "my operands are in shared memory" with a (possibly remote) arrive
    if (tid == 0) remote_arrive(mapa(smem_u32(&bar_ready), 0));
    if (rank == 0 && tid == 0) {
        long long t0 = clock64();
        while (!try_wait_cluster(smem_u32(&bar_ready), 0)) if (clock64() - t0 > 2000000000ll) { atomicExch(p.status, 1); break; }
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc(256, p.N, kF16, kF16, 0, 0);
        const int half_rb = (p.N / 2 + 127) / 128;       // 128-row blocks of this CTA's B half per K block
        for (int s = 0; s < p.K / 16; ++s) {
            const int k0 = s * 16;
            umma2_f16(tm, umma_desc_kmajor(smem_u32(sA) + (k0 / 64) * 16384, (k0 % 64) / 16),
                      umma_desc_kmajor(smem_u32(sB) + (k0 / 64) * half_rb * 16384, (k0 % 64) / 16), idesc, s > 0);
        }
        umma2_commit_mc(smem_u32(&bar_done), 3);
    }
    __syncwarp();
    {
        long long t0 = clock64();
        while (!mbar_try_wait(smem_u32(&bar_done), 0)) if (clock64() - t0 > 2000000000ll) { atomicExch(p.status, 2); break; }
    }
    tc_fence_after_sync();
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32 && c0 + j < p.N; ++j) p.d[((size_t)rank * 128 + tid) * p.N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster.sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

static uint16_t h16(float x) { __half h = __float2half_rn(x); return *reinterpret_cast<uint16_t*>(&h); }
static float f16(uint16_t u) { __half h = *reinterpret_cast<__half*>(&u); return __half2float(h); }

int main(int argc, char** argv) {
    Args p{};
    p.N = argc > 1 ? atoi(argv[1]) : 128; p.K = argc > 2 ? atoi(argv[2]) : 128;
    const int M = 256, N = p.N, K = p.K, nh = N / 2;
    std::vector<float> A((size_t)M * K), B((size_t)N * K);
    srand(77);
    for (auto& v : A) v = f16(h16((rand() % 2001 - 1000) / 1000.0f));
    for (auto& v : B) v = f16(h16((rand() % 2001 - 1000) / 1000.0f));
    p.a_bytes = (K / 64) * 16384;
    const int half_rb = (nh + 127) / 128;
    p.b_half_bytes = (K / 64) * half_rb * 16384;
    std::vector<uint8_t> ai(2 * (size_t)p.a_bytes, 0), bi(2 * (size_t)p.b_half_bytes, 0);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k)
        *reinterpret_cast<uint16_t*>(&ai[(size_t)(m / 128) * p.a_bytes + (size_t)(k / 64) * 16384 + image_offset(m % 128, k % 64)]) = h16(A[(size_t)m * K + k]);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        const int r = n / nh, nn = n % nh;
        *reinterpret_cast<uint16_t*>(&bi[(size_t)r * p.b_half_bytes + (size_t)(k / 64) * half_rb * 16384 + (size_t)(nn / 128) * 16384 + image_offset(nn % 128, k % 64)]) = h16(B[(size_t)n * K + k]);
    }
    uint8_t *da, *db; float* dd; int* ds;
    cudaMalloc(&da, ai.size()); cudaMalloc(&db, bi.size()); cudaMalloc(&dd, (size_t)M * N * 4); cudaMalloc(&ds, 4);
    cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice); cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, (size_t)M * N * 4); cudaMemset(ds, 0, 4);
    p.a_img = da; p.b_img = db; p.d = dd; p.status = ds;
    const int smem = 131072 + 1024;
    cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe2_kernel<<<2, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 3; }
    std::vector<float> D((size_t)M * N); int st = 0;
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
        double err = fabs(ref - D[(size_t)m * N + n]);
        if (!(err <= maxerr)) maxerr = err;
        if (fabs(ref) > maxref) maxref = fabs(ref);
    }
    const bool ok = st == 0 && maxerr < 1e-3 * (1 + maxref);
    printf("probe2 N=%d K=%d status=%d maxerr=%.3e maxref=%.3e D[0][0]=%f D[130][5]=%f -> %s\n", N, K, st, maxerr, maxref, D[0], D[(size_t)130 * N + 5], ok ? "PASS" : "FAIL");
    return ok ? 0 : 1;
}
