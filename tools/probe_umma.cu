// probe_umma.cu — hardware probe for the primitives in csrc/hn_tc.cuh.
// One CTA computes D[128 x N] = A[128 x K] * B[N x K]^T with tcgen05.mma from operand
// images in shared memory, for every combination of operand major-ness / format the
// library relies on, and checks the result against a host double-precision product.
//   usage: probe_umma <a_mn 0|1> <b_mn 0|1> <afmt 0=f16|1=bf16> <bfmt> <N> <K>
// Exit code 0 = match. Not part of the product; built by tools/build_probe.sh.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../nerf-3dtalker-code_b200/csrc/hn_tc.cuh"

using namespace hn;

struct ProbeArgs {
    const uint8_t* a_img; const uint8_t* b_img; float* d; int* status;
    int a_mn, b_mn, afmt, bfmt, N, K, a_bytes, b_bytes;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    uint8_t* sA = smem;                 // up to 64 KiB
    uint8_t* sB = smem + 65536;         // up to 128 KiB
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<256>(smem_u32(&tmem_base_s));
    // A through the generic proxy (what an epilogue does) ...
    for (int i = tid * 16; i < p.a_bytes; i += 128 * 16)
        *reinterpret_cast<uint4*>(sA + i) = *reinterpret_cast<const uint4*>(p.a_img + i);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    // ... B through the bulk-copy engine (what the weight producer does)
    if (tid == 0) {
        mbar_arrive_expect_tx(smem_u32(&bars[0]), p.b_bytes);
        for (int off = 0; off < p.b_bytes; off += 16384)
            bulk_g2s(smem_u32(sB + off), p.b_img + off, min(16384, p.b_bytes - off), smem_u32(&bars[0]));
        bool ok = mbar_wait(smem_u32(&bars[0]), 0);
        if (!ok) atomicExch(p.status, 1);
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc(128, p.N, p.afmt, p.bfmt, p.a_mn, p.b_mn);
        const int a_rb = 1, b_rb = (p.N + 127) / 128;            // 128-row blocks (K-major)
        const int a_cb = 2, b_cb = (p.N + 63) / 64;              // 64-col blocks  (MN-major)
        const int nsteps = p.K / 16;
        for (int s = 0; s < nsteps; ++s) {
            const int k0 = s * 16;
            uint64_t ad, bd;
            if (!p.a_mn) ad = umma_desc_kmajor(smem_u32(sA) + (k0 / 64) * a_rb * 16384, (k0 % 64) / 16);
            else         ad = umma_desc_mnmajor(smem_u32(sA) + (k0 / 128) * a_cb * 16384, (k0 % 128) / 16, 16384);
            if (!p.b_mn) bd = umma_desc_kmajor(smem_u32(sB) + (k0 / 64) * b_rb * 16384, (k0 % 64) / 16);
            else         bd = umma_desc_mnmajor(smem_u32(sB) + (k0 / 128) * b_cb * 16384, (k0 % 128) / 16, 16384);
            umma_f16(tmem_base, ad, bd, idesc, s > 0);
        }
        umma_commit(smem_u32(&bars[1]));
    }
    __syncwarp();
    if (!mbar_wait(smem_u32(&bars[1]), 0)) atomicExch(p.status, 2);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32 && c0 + j < p.N; ++j) p.d[(size_t)tid * p.N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_free<256>(tmem_base);
}

static uint16_t to_fmt(float x, int fmt) {
    if (fmt == 0) { __half h = __float2half_rn(x); return *reinterpret_cast<uint16_t*>(&h); }
    __nv_bfloat16 b = __float2bfloat16_rn(x); return *reinterpret_cast<uint16_t*>(&b);
}
static float from_fmt(uint16_t u, int fmt) {
    if (fmt == 0) { __half h = *reinterpret_cast<__half*>(&u); return __half2float(h); }
    __nv_bfloat16 b = *reinterpret_cast<__nv_bfloat16*>(&u); return __bfloat162float(b);
}
// image of a logical [R][K] operand
static std::vector<uint8_t> make_image(const std::vector<float>& x, int R, int K, int mn, int fmt, std::vector<float>& rounded) {
    size_t bytes = mn ? (size_t)((K + 127) / 128) * ((R + 63) / 64) * 16384 : (size_t)((K + 63) / 64) * ((R + 127) / 128) * 16384;
    std::vector<uint8_t> img(bytes, 0);
    rounded.resize(x.size());
    for (int r = 0; r < R; ++r) for (int k = 0; k < K; ++k) {
        uint16_t u = to_fmt(x[(size_t)r * K + k], fmt);
        rounded[(size_t)r * K + k] = from_fmt(u, fmt);
        size_t off;
        if (!mn) off = (size_t)(k / 64) * ((R + 127) / 128) * 16384 + (size_t)(r / 128) * 16384 + image_offset(r % 128, k % 64);
        else     off = (size_t)(k / 128) * ((R + 63) / 64) * 16384 + (size_t)(r / 64) * 16384 + image_offset(k % 128, r % 64);
        *reinterpret_cast<uint16_t*>(&img[off]) = u;
    }
    return img;
}

int main(int argc, char** argv) {
    if (argc < 7) { printf("usage: a_mn b_mn afmt bfmt N K\n"); return 2; }
    ProbeArgs p{};
    p.a_mn = atoi(argv[1]); p.b_mn = atoi(argv[2]); p.afmt = atoi(argv[3]); p.bfmt = atoi(argv[4]);
    p.N = atoi(argv[5]); p.K = atoi(argv[6]);
    const int M = 128;
    std::vector<float> A((size_t)M * p.K), B((size_t)p.N * p.K), Ar, Br;
    srand(1234);
    for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.0f;
    for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.0f;
    auto ai = make_image(A, M, p.K, p.a_mn, p.afmt, Ar);
    auto bi = make_image(B, p.N, p.K, p.b_mn, p.bfmt, Br);
    p.a_bytes = (int)ai.size(); p.b_bytes = (int)bi.size();
    if (p.a_bytes > 65536 || p.b_bytes > 131072) { printf("too big\n"); return 2; }
    uint8_t *da, *db; float* dd; int* ds;
    cudaMalloc(&da, ai.size()); cudaMalloc(&db, bi.size()); cudaMalloc(&dd, (size_t)M * p.N * 4); cudaMalloc(&ds, 4);
    cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xff, (size_t)M * p.N * 4); cudaMemset(ds, 0, 4);
    p.a_img = da; p.b_img = db; p.d = dd; p.status = ds;
    const int smem = 65536 + 131072 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 3; }
    std::vector<float> D((size_t)M * p.N); int st = 0;
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&st, ds, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < p.N; ++n) {
        double ref = 0;
        for (int k = 0; k < p.K; ++k) ref += (double)Ar[(size_t)m * p.K + k] * Br[(size_t)n * p.K + k];
        double err = fabs(ref - D[(size_t)m * p.N + n]);
        if (!(err <= maxerr)) maxerr = err;
        if (fabs(ref) > maxref) maxref = fabs(ref);
    }
    printf("probe a_mn=%d b_mn=%d afmt=%d bfmt=%d N=%d K=%d status=%d maxerr=%.3e maxref=%.3e D[0][0]=%f D[5][7]=%f -> %s\n",
           p.a_mn, p.b_mn, p.afmt, p.bfmt, p.N, p.K, st, maxerr, maxref, D[0], D[5 * p.N + 7],
           (st == 0 && maxerr < 1e-3 * (1 + maxref)) ? "PASS" : "FAIL");
    return (st == 0 && maxerr < 1e-3 * (1 + maxref)) ? 0 : 1;
}
