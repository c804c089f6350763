// hn_fold.cu — the latent-code side of fg_CD_predictor: the reference broadcasts shape/expression(+gaze),
// audio-style and appearance codes to every sample and concatenates them to the layer inputs (NetWorks/HeadNeRFNet.py:84-89,
// 149-152; NetWorks/models.py:69,75,80).  Here their weight columns are folded into one effective bias row per batch item,
//   bias_eff[b, FeaExt_module_0] = b0 + W0[:, 63:63+S] shape_b + W0[:, 63+S:63+S+64] audio_b
//   bias_eff[b, FeaExt_module_5] = b5 + W5[:, 63:63+S] shape_b
//   bias_eff[b, RGB_layer_1]     = br1 + WR1[:, 384:384+A] appea_b              (all other layers: the plain bias)
// and the backward turns the kernels' bias-row gradient (per-item column sums of the pre-activation gradients) into the
// gradients of the codes, of the folded weight columns and of the 12 bias vectors.  Also: the power-of-two loss scale of the
// half-precision backward chain, computed on the device from max|dL/dF|.
#include "hn_api.h"
#include "hn_mlp_sched.h"

namespace hn {

struct FoldK {
    int B, S, A;
    int r0_fused;             // 1: RGB_layer_1's row also carries W_R1[:, :384] b_R0 (the fast chains skip RGB_layer_0, hn_mlp_sched.h)
    const float* w0; int ld0; const float* w5; int ld5; const float* wr1; int ldr1;
    const float* bias[12];
    const float* shape; const float* audio; const float* appea;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// layer index (header order, density = 8) and row of bias-row entry i; -1 = padding
__device__ __forceinline__ int bias_entry(int i, int* row) {
    if (i < 8 * HN_HIDDEN) { *row = i % HN_HIDDEN; return i / HN_HIDDEN; }
    if (i < HN_BIAS_OFF_R1) { *row = i - HN_BIAS_OFF_R0; return W_R0; }
    if (i < HN_BIAS_OFF_R2) { *row = i - HN_BIAS_OFF_R1; return W_R1; }
    if (i < HN_BIAS_OFF_DENSITY) { *row = i - HN_BIAS_OFF_R2; return W_R2; }
    if (i == HN_BIAS_OFF_DENSITY) { *row = 0; return W_DENSITY; }
    return -1;
}

// one warp per (item, bias-row entry)
__global__ void __launch_bounds__(256) fold_fwd_kernel(FoldK k, float* bias_eff) {
    const int wid = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (wid >= k.B * HN_BIAS_STRIDE) return;
    const int b = wid / HN_BIAS_STRIDE, i = wid % HN_BIAS_STRIDE;
    int r;
    const int l = bias_entry(i, &r);
    float acc = 0.f;
    if (l == W_L0) {
        const float* w = k.w0 + (size_t)r * k.ld0 + HN_PE;
        for (int c = lane; c < k.S; c += 32) acc = fmaf(__ldg(w + c), __ldg(k.shape + (size_t)b * k.S + c), acc);
        for (int c = lane; c < 64; c += 32) acc = fmaf(__ldg(w + k.S + c), __ldg(k.audio + (size_t)b * 64 + c), acc);
    } else if (l == W_L5) {
        const float* w = k.w5 + (size_t)r * k.ld5 + HN_PE;
        for (int c = lane; c < k.S; c += 32) acc = fmaf(__ldg(w + c), __ldg(k.shape + (size_t)b * k.S + c), acc);
    } else if (l == W_R1) {
        const float* w = k.wr1 + (size_t)r * k.ldr1 + HN_HIDDEN;
        for (int c = lane; c < k.A; c += 32) acc = fmaf(__ldg(w + c), __ldg(k.appea + (size_t)b * k.A + c), acc);
        if (k.r0_fused) {
            const float* wh = k.wr1 + (size_t)r * k.ldr1;
            for (int c = lane; c < HN_HIDDEN; c += 32) acc = fmaf(__ldg(wh + c), __ldg(k.bias[W_R0] + c), acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) bias_eff[wid] = (l >= 0) ? __ldg(k.bias[l] + r) + acc : 0.f;
}

// d(code)[b, c] = sum_n G[b, n] W[n, c]; shape | audio | appea columns.  Block = 32 consecutive code columns x 8 slices of
// the output-channel sum (coalesced weight reads, 48-96 iterations per thread), reduced through shared memory.
__global__ void __launch_bounds__(256) fold_bwd_codes_kernel(FoldK k, const float* G, float* dshape, float* daudio, float* dappea) {
    __shared__ float part[8][32];
    const int nc = k.S + 64 + k.A;
    const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), sl = threadIdx.x >> 5;
    const float* g = G + (size_t)b * HN_BIAS_STRIDE;
    float acc = 0.f;
    if (c < k.S) {
        for (int n = sl; n < HN_HIDDEN; n += 8) {
            acc = fmaf(__ldg(g + n), __ldg(k.w0 + (size_t)n * k.ld0 + HN_PE + c), acc);
            acc = fmaf(__ldg(g + 5 * HN_HIDDEN + n), __ldg(k.w5 + (size_t)n * k.ld5 + HN_PE + c), acc);
        }
    } else if (c < k.S + 64) {
        for (int n = sl; n < HN_HIDDEN; n += 8) acc = fmaf(__ldg(g + n), __ldg(k.w0 + (size_t)n * k.ld0 + HN_PE + c), acc);
    } else if (c < nc) {
        const int cc = c - k.S - 64;
        for (int n = sl; n < HN_RGB1; n += 8) acc = fmaf(__ldg(g + HN_BIAS_OFF_R1 + n), __ldg(k.wr1 + (size_t)n * k.ldr1 + HN_HIDDEN + cc), acc);
    }
    part[sl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (sl != 0 || c >= nc) return;
#pragma unroll
    for (int i = 1; i < 8; ++i) acc += part[i][threadIdx.x];
    if (c < k.S) { if (dshape) dshape[(size_t)b * k.S + c] = acc; }
    else if (c < k.S + 64) { if (daudio) daudio[(size_t)b * 64 + c - k.S] = acc; }
    else if (dappea) dappea[(size_t)b * k.A + c - k.S - 64] = acc;
}

struct FoldGradOut { float* dw0; float* dw5; float* dwr1; float* dbias[12]; };

// folded weight columns and the bias vectors: one thread per destination element, += (each element has one owner)
__global__ void __launch_bounds__(256) fold_bwd_params_kernel(FoldK k, const float* G, FoldGradOut o) {
    const int n0 = HN_HIDDEN * (k.S + 64), n5 = HN_HIDDEN * k.S, nr = HN_RGB1 * k.A;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n0) {
        if (!o.dw0) return;
        const int n = t / (k.S + 64), c = t % (k.S + 64);
        float acc = 0.f;
        for (int b = 0; b < k.B; ++b)
            acc = fmaf(__ldg(G + (size_t)b * HN_BIAS_STRIDE + n), c < k.S ? __ldg(k.shape + (size_t)b * k.S + c) : __ldg(k.audio + (size_t)b * 64 + c - k.S), acc);
        o.dw0[(size_t)n * k.ld0 + HN_PE + c] += acc;
        return;
    }
    t -= n0;
    if (t < n5) {
        if (!o.dw5) return;
        const int n = t / k.S, c = t % k.S;
        float acc = 0.f;
        for (int b = 0; b < k.B; ++b) acc = fmaf(__ldg(G + (size_t)b * HN_BIAS_STRIDE + 5 * HN_HIDDEN + n), __ldg(k.shape + (size_t)b * k.S + c), acc);
        o.dw5[(size_t)n * k.ld5 + HN_PE + c] += acc;
        return;
    }
    t -= n5;
    if (t < nr) {
        if (!o.dwr1) return;
        const int n = t / k.A, c = t % k.A;
        float acc = 0.f;
        for (int b = 0; b < k.B; ++b) acc = fmaf(__ldg(G + (size_t)b * HN_BIAS_STRIDE + HN_BIAS_OFF_R1 + n), __ldg(k.appea + (size_t)b * k.A + c), acc);
        o.dwr1[(size_t)n * k.ldr1 + HN_HIDDEN + c] += acc;
        return;
    }
    t -= nr;
    if (t <= HN_BIAS_OFF_DENSITY) {
        int r;
        const int l = bias_entry(t, &r);
        if (l < 0 || !o.dbias[l]) return;
        float acc = 0.f;
        for (int b = 0; b < k.B; ++b) acc += __ldg(G + (size_t)b * HN_BIAS_STRIDE + t);
        o.dbias[l][r] += acc;
    }
}

// scale = 2^floor(log2(target / max|g|)); scratch = two words, zero on entry and zero again on exit (running max bits, finished-block count)
__global__ void __launch_bounds__(256) loss_scale_kernel(const float* g, int64_t n, float target, float* scale_out, unsigned* scratch) {
    float m = 0.f;
    const int64_t n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(__ldg(g + i)));
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(scratch, __float_as_uint(m));          // non-negative floats order like their bit patterns
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(scratch + 1, 1u) == gridDim.x - 1) {
            __threadfence();
            const float gmax = fmaxf(__uint_as_float(atomicMax(scratch, 0u)), 1e-30f);
            *scale_out = exp2f(floorf(log2f(target / gmax)));
            scratch[0] = 0u; scratch[1] = 0u;                       // ready for the next call: the caller zero-initialises only once
        }
    }
}

static int fill(const hn_fold_t* a, FoldK* k, const char* who) {
    if (!a || !a->w0 || !a->w5 || !a->wr1 || !a->shape_code || !a->audio || !a->appea) return set_error(HN_E_BADARG, who);
    for (int i = 0; i < 12; ++i) if (!a->bias[i]) return set_error(HN_E_BADARG, who);
    if (a->B <= 0 || a->shape_dims <= 0 || a->appea_dims <= 0 || a->ld0 < HN_PE + a->shape_dims + 64 || a->ld5 < HN_PE + a->shape_dims ||
        a->ldr1 < HN_HIDDEN + a->appea_dims)
        return set_error(HN_E_BADARG, who);
    k->B = a->B; k->S = a->shape_dims; k->A = a->appea_dims; k->r0_fused = a->r0_fused;
    k->w0 = a->w0; k->ld0 = a->ld0; k->w5 = a->w5; k->ld5 = a->ld5; k->wr1 = a->wr1; k->ldr1 = a->ldr1;
    for (int i = 0; i < 12; ++i) k->bias[i] = a->bias[i];
    k->shape = a->shape_code; k->audio = a->audio; k->appea = a->appea;
    return HN_OK;
}

}  // namespace hn

extern "C" int hn_fold_bias(const hn_fold_t* a, float* bias_eff, void* stream) {
    using namespace hn;
    FoldK k;
    if (int rc = fill(a, &k, "hn_fold_bias: null pointer or inconsistent dimensions")) return rc;
    if (!bias_eff) return set_error(HN_E_BADARG, "hn_fold_bias: null output");
    const size_t threads = (size_t)k.B * HN_BIAS_STRIDE * 32;
    fold_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(k, bias_eff);
    return check_launch("hn_fold_bias");
}

extern "C" int hn_fold_bias_bwd(const hn_fold_t* a, const float* dbias_eff, const hn_fold_grads_t* g, void* stream) {
    using namespace hn;
    FoldK k;
    if (int rc = fill(a, &k, "hn_fold_bias_bwd: null pointer or inconsistent dimensions")) return rc;
    if (!dbias_eff || !g) return set_error(HN_E_BADARG, "hn_fold_bias_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (g->dshape || g->daudio || g->dappea) {
        fold_bwd_codes_kernel<<<dim3((k.S + 64 + k.A + 31) / 32, k.B), 256, 0, st>>>(k, dbias_eff, g->dshape, g->daudio, g->dappea);
    }
    FoldGradOut o;
    o.dw0 = g->dw0; o.dw5 = g->dw5; o.dwr1 = g->dwr1;
    bool any = o.dw0 || o.dw5 || o.dwr1;
    for (int i = 0; i < 12; ++i) { o.dbias[i] = g->dbias[i]; any = any || o.dbias[i]; }
    if (any) {
        const int n = HN_HIDDEN * (k.S + 64) + HN_HIDDEN * k.S + HN_RGB1 * k.A + HN_BIAS_OFF_DENSITY + 1;
        fold_bwd_params_kernel<<<(n + 255) / 256, 256, 0, st>>>(k, dbias_eff, o);
    }
    return check_launch("hn_fold_bias_bwd");
}

namespace hn {
// Backward of the RGB_layer_0 fold (hn_mlp_sched.h): with W_f = W_R1a W_R0 and RGB_layer_1's bias row b_R1 + W_R1a b_R0 + ...,
//   dW_R1a = dW_f W_R0^T + g b_R0^T,   dW_R0 = W_R1a^T dW_f,   db_R0 = W_R1a^T g      (g = sum over items of dL/d(bias row of RGB_layer_1))
// db_R0 is delivered through RGB_layer_0's (otherwise unused) entries of item 0's bias-row gradient, from where hn_fold_bias_bwd
// picks it up like every other bias gradient.  dW_R1a contracts over the CONTIGUOUS index of both operands: one warp per element,
// lanes along k (coalesced rows, shuffle reduction); dW_R0 and db_R0: one thread per element, consecutive threads on consecutive k.
// Both products and the bias vector in ONE launch, the products as 32 x 32 output tiles (blocks [0, 72): dW_R1 = dW_f W_R0^T + g b_R0^T, 6 x 12 tiles, K = 384;
// blocks [72, 216): dW_R0 = W_R1a^T dW_f, 12 x 12 tiles, K = 192): 32-wide K chunks through shared memory, 4 outputs per thread.
// (Round 2's first version - a warp per output element straight from L2 - took 75 us for these 85 MFLOP.)
constexpr int kUnfuseTilesR1 = (HN_RGB1 / 32) * (HN_HIDDEN / 32), kUnfuseTilesR0 = (HN_HIDDEN / 32) * (HN_HIDDEN / 32);
__global__ void __launch_bounds__(256) unfuse_gemm_kernel(const hn_unfuse_t a) {
    __shared__ float As[32][33], Bs[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // ty 0..7
    if ((int)blockIdx.x == kUnfuseTilesR1 + kUnfuseTilesR0) {        // last block: db_R0 = W_R1a^T g, g = the items' summed RGB_layer_1 bias-row gradients
        float* g = &As[0][0];
        for (int t = threadIdx.x; t < HN_RGB1; t += 256) {
            float sum = 0.f;
            for (int b = 0; b < a.B; ++b) sum += a.dbias_eff[(size_t)b * HN_BIAS_STRIDE + HN_BIAS_OFF_R1 + t];
            g[t] = sum;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < HN_HIDDEN; t += 256) {
            float acc = 0.f;
#pragma unroll 16                                                   // keep 16 independent loads in flight
            for (int n = 0; n < HN_RGB1; ++n) acc = fmaf(__ldg(a.wr1 + (size_t)n * a.ldr1 + t), g[n], acc);
            a.dbias_eff[HN_BIAS_OFF_R0 + t] = acc;
            for (int b = 1; b < a.B; ++b) a.dbias_eff[(size_t)b * HN_BIAS_STRIDE + HN_BIAS_OFF_R0 + t] = 0.f;
        }
        return;
    }
    const bool r1 = (int)blockIdx.x < kUnfuseTilesR1;
    if (r1 ? !a.dwr1 : !a.dwr0) return;
    const int t = r1 ? (int)blockIdx.x : (int)blockIdx.x - kUnfuseTilesR1;
    const int m0 = (t / (HN_HIDDEN / 32)) * 32, n0 = (t % (HN_HIDDEN / 32)) * 32;   // output rows m0.., columns n0..
    const int K = r1 ? HN_HIDDEN : HN_RGB1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};                              // outputs (m0 + ty + 8 r, n0 + tx)
    for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = ty + 8 * r;
            if (r1) {       // A[m, k] = dwf[m, k];  B[n, k] = wr0[n, k]      (both k-contiguous)
                As[row][tx] = __ldg(a.dwf + (size_t)(m0 + row) * HN_HIDDEN + k0 + tx);
                Bs[row][tx] = __ldg(a.wr0 + (size_t)(n0 + row) * a.ldr0 + k0 + tx);
            } else {        // A[m, k] = wr1[k, m] (m-contiguous: As[k][m]);  B[n, k] = dwf[k, n] (n-contiguous: Bs[k][n])
                As[row][tx] = __ldg(a.wr1 + (size_t)(k0 + row) * a.ldr1 + m0 + tx);
                Bs[row][tx] = __ldg(a.dwf + (size_t)(k0 + row) * HN_HIDDEN + n0 + tx);
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const float bv = r1 ? Bs[tx][k] : Bs[k][tx];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = fmaf(r1 ? As[ty + 8 * r][k] : As[k][ty + 8 * r], bv, acc[r]);
        }
        __syncthreads();
    }
    if (r1) {
        float br = __ldg(a.b_r0 + n0 + tx);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = m0 + ty + 8 * r;
            float g = 0.f;                                            // the items' bias-row gradients of RGB_layer_1 output n
            for (int b = 0; b < a.B; ++b) g += a.dbias_eff[(size_t)b * HN_BIAS_STRIDE + HN_BIAS_OFF_R1 + n];
            a.dwr1[(size_t)n * a.ldr1 + n0 + tx] += fmaf(g, br, acc[r]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) a.dwr0[(size_t)(m0 + ty + 8 * r) * a.ldr0 + n0 + tx] += acc[r];
    }
}

}  // namespace hn

extern "C" int hn_unfuse_r0r1(const hn_unfuse_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->wr0 || !a->wr1 || !a->b_r0 || !a->dwf || !a->dbias_eff || a->B <= 0 || a->ldr0 < HN_HIDDEN || a->ldr1 < HN_HIDDEN)
        return set_error(HN_E_BADARG, "hn_unfuse_r0r1: null pointer or inconsistent dimensions");
    unfuse_gemm_kernel<<<kUnfuseTilesR1 + kUnfuseTilesR0 + 1, 256, 0, (cudaStream_t)stream>>>(*a);
    return check_launch("hn_unfuse_r0r1");
}

extern "C" int hn_loss_scale(const float* g, int64_t n, float target, float* scale_out, void* scratch8, void* stream) {
    using namespace hn;
    if (!g || n <= 0 || !scale_out || !scratch8 || !(target > 0.f)) return set_error(HN_E_BADARG, "hn_loss_scale: bad argument");
    int blocks = (int)((n / 4 + 255) / 256);
    blocks = blocks < 1 ? 1 : (blocks > 592 ? 592 : blocks);
    loss_scale_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, n, target, scale_out, (unsigned*)scratch8);
    return check_launch("hn_loss_scale");
}
