// hn_train.cu — the two steps either side of the rendering path in a training iteration:
//   * the photometric loss of the reference (Utils/HeadNeRFLossUtils.py:125-140 calc_data_loss, :196-236 calc_total_loss):
//       bg_loss      = mean((bg_img - bg_value)^2)
//       head_loss    = mean over {mask >= 0.5} x 3 channels of (merge_img - gt)^2        (F.mse_loss on the masked selection)
//       nonhead_loss = mean over {mask <  0.5} x 3 channels of (merge_img - bg_value)^2
//       total        = bg_loss + head_loss + nonhead_loss          with merge_img = nan_to_num(merge_img, nan=0)
//     as ONE reduction kernel (deterministic: per-block partials, the last block folds them in a fixed order) and ONE gradient
//     kernel, instead of ~25 masked-select / elementwise launches;
//   * Adam over one flat fp32 buffer (talker_trainer.py:722-723 torch.optim.Adam(model.parameters(), lr)), the arithmetic of
//     torch.optim.Adam's single-tensor path op for op, with the data-parallel 1/world averaging of the all-reduced gradient
//     folded in, instead of one launch group per parameter tensor.
// Both are HBM-streaming kernels: float4 accesses, grid = a multiple of the SM count.
#include "hn_api.h"

namespace hn {

constexpr int kLossThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

// nan_to_num(x, nan=0): NaN -> 0, +-inf -> +-FLT_MAX (torch defaults)
__device__ __forceinline__ float nan_to_num0(float x) {
    if (x != x) return 0.f;
    return fminf(fmaxf(x, -3.402823466e38f), 3.402823466e38f);
}

// partial sums per block: [0] sum head, [1] count head pixels, [2] sum nonhead, [3] count nonhead pixels, [4] sum bg
__global__ void __launch_bounds__(kLossThreads) photo_loss_reduce_kernel(const hn_photo_loss_t a, const int n_img_px, const int n_bg_el) {
    __shared__ float red[kLossThreads / 32][5];
    __shared__ bool last;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    const int hw = a.HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_img_px; i += gridDim.x * blockDim.x) {
        const int b = i / hw, p = i - b * hw;
        const float m = __ldg(a.mask + i);
        const bool head = m >= 0.5f;                       // (mask_tensor >= 0.5) / (mask_tensor < 0.5): NaN masks belong to neither
        const bool non = m < 0.5f;
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const size_t e = ((size_t)b * 3 + c) * hw + p;
            const float v = nan_to_num0(__ldg(a.img + e));
            const float d = v - (head ? __ldg(a.gt + e) : a.bg_value);
            s += d * d;
        }
        if (head) { acc[0] += s; acc[1] += 1.f; }
        else if (non) { acc[2] += s; acc[3] += 1.f; }
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_bg_el; i += gridDim.x * blockDim.x) {
        const float d = __ldg(a.bg_img + i) - a.bg_value;
        acc[4] += d * d;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        float v = 0.f;
        for (int w = 0; w < kLossThreads / 32; ++w) v += red[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last block folds the partials in a fixed order (double accumulators): the result does not depend on block scheduling
    if (threadIdx.x < 32) {
        double s[5] = {0, 0, 0, 0, 0};
        for (unsigned blk = threadIdx.x; blk < gridDim.x; blk += 32)
#pragma unroll
            for (int k = 0; k < 5; ++k) s[k] += (double)a.partials[(size_t)blk * 8 + k];
#pragma unroll
        for (int k = 0; k < 5; ++k)
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], sft);
        if (threadIdx.x == 0) {
            const double n_head = 3.0 * s[1], n_non = 3.0 * s[3];
            const float bg = (float)(s[4] / (double)n_bg_el);
            const float head = (float)(s[0] / n_head);          // empty selection: 0/0 = NaN, as F.mse_loss on an empty tensor
            const float non = (float)(s[2] / n_non);
            a.out[0] = bg; a.out[1] = head; a.out[2] = non; a.out[3] = bg + head + non;
            a.out[4] = (float)n_head; a.out[5] = (float)n_non; a.out[6] = (float)n_bg_el; a.out[7] = 0.f;
            *a.ticket = 0u;                                     // the scratch word is left zero for the next call
        }
    }
}

__global__ void __launch_bounds__(kLossThreads) photo_loss_grad_kernel(const hn_photo_loss_t a, const float* gout, float* d_img, float* d_bg,
                                                                       const int n_img_px, const int n_bg_el) {
    // gout = dL/d(bg_loss, head_loss, nonhead_loss, total): every term is differentiable, total = their sum
    const float g3 = gout ? __ldg(gout + 3) : 1.f;
    const float g_bg = g3 + (gout ? __ldg(gout + 0) : 0.f), g_head = g3 + (gout ? __ldg(gout + 1) : 0.f), g_non = g3 + (gout ? __ldg(gout + 2) : 0.f);
    const float k_head = 2.f * g_head / a.out[4], k_non = 2.f * g_non / a.out[5], k_bg = 2.f * g_bg / a.out[6];
    const int hw = a.HW;
    if (d_img) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_img_px; i += gridDim.x * blockDim.x) {
            const int b = i / hw, p = i - b * hw;
            const float m = __ldg(a.mask + i);
            const bool head = m >= 0.5f, non = m < 0.5f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const size_t e = ((size_t)b * 3 + c) * hw + p;
                const float raw = __ldg(a.img + e);
                const bool finite = (raw - raw) == 0.f;                      // d nan_to_num / dx = isfinite(x)
                const float v = nan_to_num0(raw);
                float d = 0.f;
                if (head) d = k_head * (v - __ldg(a.gt + e));
                else if (non) d = k_non * (v - a.bg_value);
                d_img[e] = finite ? d : 0.f;
            }
        }
    }
    if (d_bg) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_bg_el; i += gridDim.x * blockDim.x)
            d_bg[i] = k_bg * (__ldg(a.bg_img + i) - a.bg_value);
    }
}

// torch.optim.Adam (single-tensor path, amsgrad = False, maximize = False), one thread per 4 elements:
//   g      = grad * grad_scale (+ weight_decay * p)
//   m      = m + (g - m) * (1 - beta1)                      (lerp)
//   v      = v * beta2 + g * g * (1 - beta2)                (mul, addcmul)
//   denom  = sqrt(v) / sqrt(1 - beta2^t) + eps
//   p      = p - (lr / (1 - beta1^t)) * (m / denom)         (addcdiv)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   const int64_t n, const float grad_scale, const float weight_decay, const float one_minus_b1,
                                                   const float b2, const float one_minus_b2, const float bc2_sqrt, const float eps, const float step_size) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto upd = [&](float& pp, const float gg_in, float& mm, float& vv) {
        float gg = gg_in * grad_scale;
        if (weight_decay != 0.f) gg = fmaf(weight_decay, pp, gg);
        mm = mm + (gg - mm) * one_minus_b1;
        vv = vv * b2 + gg * gg * one_minus_b2;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pp = pp - step_size * (mm / denom);
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 P = reinterpret_cast<float4*>(p)[i], M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
        const float4 G = __ldg(reinterpret_cast<const float4*>(g) + i);
        upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
        reinterpret_cast<float4*>(p)[i] = P; reinterpret_cast<float4*>(m)[i] = M; reinterpret_cast<float4*>(v)[i] = V;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
}

static int sm_count() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

}  // namespace hn

extern "C" size_t hn_photo_loss_workspace_bytes(void) { return (size_t)(4 * 148 * 2) * 8 * sizeof(float) + 16; }

static int photo_grid(const hn_photo_loss_t* a) {
    const int cap = (int)((hn_photo_loss_workspace_bytes() - 16) / (8 * sizeof(float)));
    int grid = 4 * hn::sm_count();
    const int64_t work = (int64_t)a->B * a->HW;
    const int need = (int)((work + hn::kLossThreads - 1) / hn::kLossThreads);
    grid = grid < need ? grid : need;
    grid = grid < 1 ? 1 : grid;
    return grid < cap ? grid : cap;
}

extern "C" int hn_photo_loss_fwd(const hn_photo_loss_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->img || !a->bg_img || !a->gt || !a->mask || !a->partials || !a->ticket || !a->out)
        return set_error(HN_E_BADARG, "hn_photo_loss_fwd: null pointer");
    if (a->B <= 0 || a->B_bg <= 0 || a->HW <= 0) return set_error(HN_E_BADARG, "hn_photo_loss_fwd: empty image");
    if ((int64_t)a->B * a->HW * 3 > 0x7fffffff) return set_error(HN_E_UNSUPPORTED, "hn_photo_loss_fwd: more than 2^31 image elements");
    photo_loss_reduce_kernel<<<photo_grid(a), kLossThreads, 0, (cudaStream_t)stream>>>(*a, a->B * a->HW, a->B_bg * 3 * a->HW);
    return check_launch("hn_photo_loss_fwd");
}

extern "C" int hn_photo_loss_bwd(const hn_photo_loss_t* a, const float* gout, float* d_img, float* d_bg_img, void* stream) {
    using namespace hn;
    if (!a || !a->img || !a->bg_img || !a->gt || !a->mask || !a->out) return set_error(HN_E_BADARG, "hn_photo_loss_bwd: null pointer");
    if (!d_img && !d_bg_img) return HN_OK;
    photo_loss_grad_kernel<<<photo_grid(a), kLossThreads, 0, (cudaStream_t)stream>>>(*a, gout, d_img, d_bg_img, a->B * a->HW, a->B_bg * 3 * a->HW);
    return check_launch("hn_photo_loss_bwd");
}

extern "C" int hn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const hn_adam_t* h, void* stream) {
    using namespace hn;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !h) return set_error(HN_E_BADARG, "hn_adam_step: null pointer");
    if (n <= 0) return HN_OK;
    if (h->step < 1) return set_error(HN_E_BADARG, "hn_adam_step: step counts from 1");
    if ((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) != 0)
        return set_error(HN_E_BADARG, "hn_adam_step: buffers must be 16-byte aligned");
    // bias corrections on the host in double, rounded once, as torch does with Python floats
    const double bc1 = 1.0 - pow((double)h->beta1, (double)h->step);
    const double bc2 = 1.0 - pow((double)h->beta2, (double)h->step);
    const float step_size = (float)((double)h->lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    int grid = 8 * sm_count();
    const int64_t need = ((n >> 2) + 255) / 256 + 1;
    if (need < grid) grid = (int)need;
    adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, h->grad_scale, h->weight_decay,
                                                        (float)(1.0 - (double)h->beta1), h->beta2, (float)(1.0 - (double)h->beta2), bc2_sqrt, h->eps, step_size);
    return check_launch("hn_adam_step");
}
