// probe_ts.cu — hardware probe for the second-generation fused kernels (design aid, not part of the product):
//   mode 0: correctness of tcgen05.mma with the A operand in TENSOR MEMORY (written with tcgen05.st as packed f16 pairs,
//           lane = row, 32-bit column c = elements 2c, 2c+1) against a host product.      probe_ts 0 <N>
//   mode 1: sustained MMA rate per SM (A from TMEM or SMEM) with a concurrent weight-ring fill stream, unicast or
//           multicast over a cluster of CL CTAs.                                            probe_ts 1 <N> <ts 0|1> <CL 0=no fill|1|2|4>
//   mode 2: as mode 1 without MMAs (pure L2 -> SMEM fill rate).                             probe_ts 2 <CL>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../nerf-3dtalker-code_b200/csrc/hn_tc.cuh"
using namespace hn;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}

__device__ __forceinline__ void umma_f16_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi) : "memory");
}

// ------------------------------------------------------------------ mode 0
__global__ void __launch_bounds__(128, 1) ts_check_kernel(const __half* A, const uint8_t* b_img, int b_bytes, float* D, int N, int* status) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_s));
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tm = tmem_s;
    // A row `tid` (64 halfs) -> 32 packed columns at TMEM column 256
    uint32_t v[32];
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)tid * 64);
    for (int i = 0; i < 32; ++i) v[i] = arow[i];
    tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + 256, v);
    tmem_st_wait();
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    if (tid == 0) {
        mbar_arrive_expect_tx(smem_u32(&bars[0]), b_bytes);
        bulk_g2s(smem, b_img, b_bytes, smem_u32(&bars[0]));
        if (!mbar_wait(smem_u32(&bars[0]), 0)) atomicExch(status, 1);
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc(128, N, kF16, kF16, 0, 0);
        for (int ks = 0; ks < 4; ++ks) umma_f16_ts(tm, tm + 256 + ks * 8, umma_desc_kmajor(smem, ks), idesc, ks > 0);
        umma_commit(smem_u32(&bars[1]));
    }
    __syncwarp();
    if (!mbar_wait(smem_u32(&bars[1]), 0)) atomicExch(status, 2);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t o[32];
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, o);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(o[j]);
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_free<512>(tm);
}

// ------------------------------------------------------------------ modes 1, 2
struct RateShared {
    uint64_t done, full[8], empty[8];
    uint32_t tmem_s;
    volatile int abort;
};
__device__ __forceinline__ bool wait_to(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
    const uint32_t b = smem_u32(bar);
    if (mbar_try_wait_cluster(b, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(b, parity)) {
        if (*abort_flag) return false;
        if (clock64() - t0 > 1000000000ll) { *abort_flag = 1; return false; }
    }
    return true;
}

template <int CL>
__global__ void __launch_bounds__(128, 1) rate_kernel(const uint8_t* src, long long* out, int N, int ts, int reps, int fills, int replicas, int pieces, int stages) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ RateShared sh;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int NS = 8;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);
    if (tid == 0) {
        mbar_init(smem_u32(&sh.done), 1);
        for (int i = 0; i < NS; ++i) { mbar_init(smem_u32(&sh.full[i]), 1); mbar_init(smem_u32(&sh.empty[i]), CL); }
        sh.abort = 0; mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<512>(smem_u32(&sh.tmem_s));
    tc_fence_before_sync(); __syncthreads();
    if (CL > 1) cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tm = sh.tmem_s;
    // smem: B region 64 KiB at 0, A (SS mode) 32 KiB at 64K, fill ring NS x 16 KiB at 96K
    if (warp == 0 && reps > 0) {
        // warp-uniform issue loop (descriptors stay in uniform registers); one elected lane issues
        const uint32_t idesc = umma_idesc(128, N, kF16, kF16, 0, 0);
        const uint32_t bslot = N > 128 ? 32768 : 16384, nmask = 65536 / bslot - 1;
        const long long t0 = clock64();
        uint32_t acol = 0;
        for (int r = 0; r < reps; ++r) {
            const uint32_t b_lo = desc_lo(smem + (r & nmask) * bslot, 16), a_lo = desc_lo(smem + 65536 + (r & 1) * 16384, 16);
            const uint32_t d = tm + ((r * N) & 255);
            const uint32_t a_t = tm + 256 + acol;
            acol = (acol + 32 >= 192) ? 0 : acol + 32;
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    if (ts) umma_f16_ts_lo(d, a_t + ks * 8, b_lo + ks * 2, idesc, 1u);
                    else umma_f16_lohi(d, a_lo + ks * 2, b_lo + ks * 2, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&sh.done));
        __syncwarp();
        const long long t1 = clock64();
        wait_to(&sh.done, 0, &sh.abort);
        const long long t2 = clock64();
        if (tid == 0) { out[blockIdx.x * 4 + 0] = t1 - t0; out[blockIdx.x * 4 + 1] = t2 - t0; }
    } else if (tid == 32 && fills > 0 && CL == 1) {
        // single CTA: one thread re-arms and re-issues a stage as soon as it has landed (no consumer work)
        const long long t0 = clock64();
        for (int n = 0; n < fills + stages && !sh.abort; ++n) {
            const int s = n % stages; const uint32_t par = ((n / stages) & 1) ^ 1;
            if (n >= stages && !wait_to(&sh.full[s], par, &sh.abort)) break;
            if (n >= fills) continue;
            const uint8_t* p = src + ((size_t)(blockIdx.x % replicas) * 170 + (n % 170)) * 16384;
            mbar_arrive_expect_tx(smem_u32(&sh.full[s]), 16384);
            const uint32_t pb = 16384 / pieces;
            for (int q = 0; q < pieces; ++q) bulk_g2s(smem + 98304 + s * 16384 + q * pb, p + q * pb, pb, smem_u32(&sh.full[s]));
        }
        out[blockIdx.x * 4 + 2] = clock64() - t0;
        out[blockIdx.x * 4 + 3] = sh.abort;
    } else if (tid == 32 && fills > 0) {
        // cluster: stage s is always issued by rank s % CL (stages is a multiple of CL) once every CTA has released it
        for (int n = (int)rank; n < fills && !sh.abort; n += CL) {
            const int s = n % stages; const uint32_t par = (n / stages) & 1;
            if (!wait_to(&sh.empty[s], par ^ 1, &sh.abort)) break;
            const uint8_t* p = src + ((size_t)((blockIdx.x / CL) % replicas) * 170 + (n % 170)) * 16384;
            const uint32_t pb = 16384 / pieces;
            for (int q = 0; q < pieces; ++q)
                bulk_g2s_multicast(smem + 98304 + s * 16384 + q * pb, p + q * pb, pb, smem_u32(&sh.full[s]), kMask);
        }
    } else if (tid == 64 && fills > 0 && CL > 1) {
        const long long t0 = clock64();
        for (int n = 0; n < fills && !sh.abort; ++n) {
            const int s = n % stages; const uint32_t par = (n / stages) & 1;
            mbar_arrive_expect_tx(smem_u32(&sh.full[s]), 16384);
            if (!wait_to(&sh.full[s], par, &sh.abort)) break;
            asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32(&sh.empty[s]), n % CL)) : "memory");
        }
        out[blockIdx.x * 4 + 2] = clock64() - t0;
        out[blockIdx.x * 4 + 3] = sh.abort;
    }
    tc_fence_before_sync(); __syncthreads();
    if (CL > 1) cluster_sync_all();
    if (warp == 0) tmem_free<512>(tm);
}

template <int CL>
static cudaError_t launch_rate(int grid, int smem, const uint8_t* src, long long* out, int N, int ts, int reps, int fills, int replicas, int pieces, int stages) {
    cudaFuncSetAttribute(rate_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid / CL * CL); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = CL; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, rate_kernel<CL>, src, out, N, ts, reps, fills, replicas, pieces, stages);
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    if (mode == 0) {
        const int N = argc > 2 ? atoi(argv[2]) : 128, M = 128, K = 64;
        std::vector<__half> A((size_t)M * K), B((size_t)N * K);
        std::vector<float> Af(A.size()), Bf(B.size());
        srand(7);
        for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2half_rn((rand() % 2001 - 1000) / 1000.0f); Af[i] = __half2float(A[i]); }
        for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2half_rn((rand() % 2001 - 1000) / 1000.0f); Bf[i] = __half2float(B[i]); }
        std::vector<uint8_t> bimg((size_t)N * 128, 0);
        for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) *reinterpret_cast<__half*>(&bimg[image_offset(r, k)]) = B[(size_t)r * K + k];
        __half* dA; uint8_t* dB; float* dD; int* ds;
        cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, bimg.size()); cudaMalloc(&dD, (size_t)M * N * 4); cudaMalloc(&ds, 4);
        cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, bimg.data(), bimg.size(), cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, (size_t)M * N * 4); cudaMemset(ds, 0, 4);
        cudaFuncSetAttribute(ts_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        ts_check_kernel<<<1, 128, 65536>>>(dA, dB, (int)bimg.size(), dD, N, ds);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode 0 N=%d CUDA error: %s\n", N, cudaGetErrorString(e)); return 3; }
        std::vector<float> D((size_t)M * N); int st = 0;
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)Af[(size_t)m * K + k] * Bf[(size_t)n * K + k];
            const double err = fabs(ref - D[(size_t)m * N + n]);
            if (!(err <= maxerr)) maxerr = err;
            maxref = std::max(maxref, fabs(ref));
        }
        const bool ok = st == 0 && maxerr < 1e-3 * (1 + maxref);
        printf("mode 0 (A in TMEM) N=%d status=%d maxerr=%.3e maxref=%.3e D[0][0]=%f D[5][7]=%f -> %s\n", N, st, maxerr, maxref, D[0], D[5 * N + 7], ok ? "PASS" : "FAIL");
        return ok ? 0 : 1;
    }
    int N = 128, ts = 1, CL = 1, reps = 4096, fills = 2048;
    if (mode == 1) { N = atoi(argv[2]); ts = atoi(argv[3]); CL = atoi(argv[4]); if (CL == 0) { fills = 0; CL = 1; } }
    else { CL = atoi(argv[2]); reps = 0; }
    if (argc > 5) reps = atoi(argv[5]);
    if (argc > 6) fills = atoi(argv[6]);
    const int replicas = argc > 7 ? atoi(argv[7]) : 1, pieces = argc > 8 ? atoi(argv[8]) : 1, stages = argc > 9 ? atoi(argv[9]) : 8;
    uint8_t* src; long long* out;
    cudaMalloc(&src, (size_t)replicas * 170 * 16384); cudaMemset(src, 0, (size_t)replicas * 170 * 16384);
    cudaMalloc(&out, 148 * 4 * 8); cudaMemset(out, 0, 148 * 4 * 8);
    const int smem = 98304 + 8 * 16384 + 1024;
    cudaError_t e = cudaSuccess;
    for (int it = 0; it < 2 && e == cudaSuccess; ++it) {
        if (CL == 1) e = launch_rate<1>(148, smem, src, out, N, ts, reps, fills, replicas, pieces, stages);
        else if (CL == 2) e = launch_rate<2>(148, smem, src, out, N, ts, reps, fills, replicas, pieces, stages);
        else if (CL == 3) e = launch_rate<3>(147, smem, src, out, N, ts, reps, fills, replicas, pieces, stages);
        else e = launch_rate<4>(148, smem, src, out, N, ts, reps, fills, replicas, pieces, stages);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    std::vector<long long> h(148 * 4);
    cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
    const int grid = CL == 3 ? 147 : 148;
    double mma_max = 0, fill_max = 0, fill_min = 1e30; int aborted = 0;
    for (int b = 0; b < grid; ++b) {
        mma_max = std::max(mma_max, (double)h[b * 4 + 1]);
        fill_max = std::max(fill_max, (double)h[b * 4 + 2]); fill_min = std::min(fill_min, (double)h[b * 4 + 2]);
        aborted += (int)h[b * 4 + 3];
    }
    printf("mode %d N=%d %s CL=%d rep=%d pcs=%d stages=%d: %s aborted=%d | MMA %.1f cyc (floor %d) | fill %.1f B/cyc/SM (slowest SM; fastest %.1f)\n", mode, N, ts ? "A=TMEM" : "A=SMEM",
           CL, replicas, pieces, stages, cudaGetErrorString(e), aborted, reps ? mma_max / (reps * 4.0) : 0.0, 128 * N / 256,
           fills ? fills * 16384.0 / fill_max : 0.0, fills ? fills * 16384.0 / fill_min : 0.0);
    return e != cudaSuccess;
}
