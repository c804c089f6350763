// probe_fill.cu — L2 -> shared-memory fill rate per SM with all SMs streaming the same 2.7 MB weight image
// (design aid, not part of the product).   probe_fill <mech> <CL> <stages> <box_rows>
//   mech 0: cp.async.bulk (1-D)              mech 1: cp.async.bulk.tensor.2d (tensor map, box 64 x box_rows halfs)
//   mech 2: ld.global.v4 + st.shared.v4 by 4 loader warps (CL ignored)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../nerf-3dtalker-code_b200/csrc/hn_tc.cuh"
using namespace hn;

struct Sh { uint64_t full[8], empty[8]; volatile int abort; };

__device__ int g_flavor;
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool poll(uint32_t b, uint32_t parity, int flavor) {
    return flavor == 0 ? mbar_try_wait(b, parity) : flavor == 1 ? mbar_try_wait_cluster(b, parity) : mbar_test_wait(b, parity);
}
__device__ __forceinline__ bool wait_to(uint64_t* bar, uint32_t parity, volatile int* abort_flag) {
    const uint32_t b = smem_u32(bar);
    const int flavor = g_flavor;
    if (poll(b, parity, flavor)) return true;
    const long long t0 = clock64();
    while (!poll(b, parity, flavor)) {
        if (*abort_flag) return false;
        if (clock64() - t0 > 1000000000ll) { *abort_flag = 1; return false; }
    }
    return true;
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_2d_mc(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar), "h"(mask) : "memory");
}

template <int CL>
__global__ void __launch_bounds__(256, 1) fill_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* src, long long* out,
                                                      int mech, int fills, int stages, int stage_bytes) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ Sh sh;
    const int tid = threadIdx.x;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&sh.full[i]), 1); mbar_init(smem_u32(&sh.empty[i]), CL); }
        sh.abort = 0; mbar_fence_init();
    }
    __syncthreads();
    if (CL > 1) cluster_sync_all();
    const int n_src = (170 * 16384) / stage_bytes;               // source blocks in the 2.7 MB image
    const int rows = stage_bytes / 128;
    if (mech == 2) {
        if (tid >= 128) {                                         // 4 loader warps: 128 threads x 16 B = 2 KiB per pass
            const int t = tid - 128;
            const long long t0 = clock64();
            for (int n = 0; n < fills; ++n) {
                const uint8_t* p = src + (size_t)(n % n_src) * stage_bytes;
                const uint32_t dst = smem + (n % stages) * stage_bytes;
                for (int off = 0; off < stage_bytes; off += 8 * 2048) {
                    uint4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __ldg(reinterpret_cast<const uint4*>(p + off + j * 2048 + t * 16));
#pragma unroll
                    for (int j = 0; j < 8; ++j) st_shared_v4(dst + off + j * 2048 + t * 16, v[j].x, v[j].y, v[j].z, v[j].w);
                }
            }
            if (t == 0) { out[blockIdx.x * 2] = clock64() - t0; out[blockIdx.x * 2 + 1] = 0; }
        }
    } else if (mech == 3) {
        // 1-D bulk copies issued by FOUR threads of different warps, each with its own ring of `stages` stages
        // box_rows >= 1000: the four issuers are lanes 0..3 of ONE warp instead of lane 0 of four warps
        const bool same_warp = g_flavor >= 10;
        if (same_warp ? (tid >= 32 && tid < 36) : ((tid & 31) == 0 && tid >= 32 && tid < 160)) {
            const int j = same_warp ? tid - 32 : (tid >> 5) - 1;
            const long long t0 = clock64();
            const int my = fills / 4;
            for (int n = 0; n < my + stages && !sh.abort; ++n) {
                const int s = n % stages; const uint32_t par = ((n / stages) & 1) ^ 1;
                if (n >= stages && !wait_to(&sh.full[j * 2 + s], par, &sh.abort)) break;
                if (n >= my) continue;
                mbar_arrive_expect_tx(smem_u32(&sh.full[j * 2 + s]), stage_bytes);
                bulk_g2s(smem + (j * stages + s) * stage_bytes, src + (size_t)((n * 4 + j) % n_src) * stage_bytes, stage_bytes, smem_u32(&sh.full[j * 2 + s]));
            }
            if (j == 0) { out[blockIdx.x * 2] = clock64() - t0; out[blockIdx.x * 2 + 1] = sh.abort; }
        }
    } else if (CL == 1) {
        if (tid == 32) {
            const long long t0 = clock64();
            for (int n = 0; n < fills + stages && !sh.abort; ++n) {
                const int s = n % stages; const uint32_t par = ((n / stages) & 1) ^ 1;
                if (n >= stages && !wait_to(&sh.full[s], par, &sh.abort)) break;
                if (n >= fills) continue;
                mbar_arrive_expect_tx(smem_u32(&sh.full[s]), stage_bytes);
                if (mech == 0) bulk_g2s(smem + s * stage_bytes, src + (size_t)(n % n_src) * stage_bytes, stage_bytes, smem_u32(&sh.full[s]));
                else tma_2d(smem + s * stage_bytes, &tmap, 0, (n % n_src) * rows, smem_u32(&sh.full[s]));
            }
            out[blockIdx.x * 2] = clock64() - t0; out[blockIdx.x * 2 + 1] = sh.abort;
        }
    } else {
        if (tid == 32) {
            for (int n = (int)rank; n < fills && !sh.abort; n += CL) {
                const int s = n % stages; const uint32_t par = (n / stages) & 1;
                if (!wait_to(&sh.empty[s], par ^ 1, &sh.abort)) break;
                if (mech == 0) bulk_g2s_multicast(smem + s * stage_bytes, src + (size_t)(n % n_src) * stage_bytes, stage_bytes, smem_u32(&sh.full[s]), kMask);
                else tma_2d_mc(smem + s * stage_bytes, &tmap, 0, (n % n_src) * rows, smem_u32(&sh.full[s]), kMask);
            }
        } else if (tid == 64) {
            const long long t0 = clock64();
            for (int n = 0; n < fills && !sh.abort; ++n) {
                const int s = n % stages; const uint32_t par = (n / stages) & 1;
                mbar_arrive_expect_tx(smem_u32(&sh.full[s]), stage_bytes);
                if (!wait_to(&sh.full[s], par, &sh.abort)) break;
                asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32(&sh.empty[s]), n % CL)) : "memory");
            }
            out[blockIdx.x * 2] = clock64() - t0; out[blockIdx.x * 2 + 1] = sh.abort;
        }
    }
    __syncthreads();
    if (CL > 1) cluster_sync_all();
}

template <int CL>
static cudaError_t launch(const CUtensorMap& tm, int smem, const uint8_t* src, long long* out, int mech, int fills, int stages, int stage_bytes) {
    cudaFuncSetAttribute(fill_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148 / CL * CL); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = CL; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fill_kernel<CL>, tm, src, out, mech, fills, stages, stage_bytes);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int mech = argc > 1 ? atoi(argv[1]) : 0, CL = argc > 2 ? atoi(argv[2]) : 1, stages = argc > 3 ? atoi(argv[3]) : 4;
    const int box_rows = argc > 4 ? atoi(argv[4]) : 128;
    const int stage_bytes = box_rows * 128;
    const int flavor = argc > 5 ? atoi(argv[5]) : 0;
    cudaMemcpyToSymbol(g_flavor, &flavor, 4);
    const int fills = (64 << 20) / stage_bytes;                   // 64 MiB per SM
    uint8_t* src; long long* out;
    cudaMalloc(&src, 170 * 16384); cudaMemset(src, 0, 170 * 16384);
    cudaMalloc(&out, 148 * 2 * 8); cudaMemset(out, 0, 148 * 2 * 8);
    CUtensorMap tm{};
    {
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
        const cuuint64_t dims[2] = {64, (cuuint64_t)170 * 128};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
    }
    const int smem = (mech == 3 ? 4 : 1) * stages * stage_bytes + 1024;
    cudaError_t e = cudaSuccess;
    for (int it = 0; it < 2 && e == cudaSuccess; ++it) {
        if (CL == 1) e = launch<1>(tm, smem, src, out, mech, fills, stages, stage_bytes);
        else if (CL == 2) e = launch<2>(tm, smem, src, out, mech, fills, stages, stage_bytes);
        else e = launch<4>(tm, smem, src, out, mech, fills, stages, stage_bytes);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
    }
    std::vector<long long> h(148 * 2);
    cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
    double tmax = 0, tmin = 1e30; int aborted = 0;
    for (int b = 0; b < 148; ++b) { tmax = std::max(tmax, (double)h[b * 2]); tmin = std::min(tmin, (double)h[b * 2]); aborted += (int)h[b * 2 + 1]; }
    printf("mech=%d CL=%d stages=%d wait=%d stage=%d B: %s aborted=%d | %.1f B/cyc/SM (slowest SM; fastest %.1f)\n", mech, CL, stages, flavor, stage_bytes,
           cudaGetErrorString(e), aborted, (double)fills * stage_bytes / tmax, (double)fills * stage_bytes / tmin);
    return e != cudaSuccess;
}
