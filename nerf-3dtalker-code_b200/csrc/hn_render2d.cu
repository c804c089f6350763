// hn_render2d.cu — first pieces of the consumer: the memory-bound tails of NeuralRenderer's
// up-sampling blocks as single kernels, forward and backward, NCHW fp32 like the reference modules.
//   upsample tail (NetWorks/PixelShuffleUpsample.py:36-45): y = blur(pixel_shuffle(leaky_relu(z2, 0.2) + repeat(x, 4), 2))
//       z2 = layer_2's pre-activation [B,4C,H,W], x = the block input [B,C,H,W], y [B,C,2H,2W]
//       (PyTorch: leaky_relu + repeat + add + pixel_shuffle + reflection pad + depthwise conv = 6 kernels and 5 intermediates)
//   rgb up-sampling (NetWorks/neural_renderer.py:47-50): y = blur(bilinear x2, align_corners=False)
// blur = kornia filter2d(normalized=True, border 'reflect') with the separable 3-tap kernel f (x) f / sum.
// The 1x1 convolutions around them stay library GEMMs.
#include "hn_api.h"

namespace hn {

__device__ __forceinline__ int reflect_idx(int t, int n) { return t < 0 ? -t : (t >= n ? 2 * n - 2 - t : t); }
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

struct Taps3 { float w[3]; };

// element (b, c, Y, X) of pixel_shuffle(leaky_relu(z2) + repeat(x, 4)) read through the un-shuffled tensors
__device__ __forceinline__ float shuffled(const float* __restrict__ z2, const float* __restrict__ x, int b, int c, int Y, int X,
                                          int C, int H, int W, float slope) {
    const int i = Y >> 1, j = X >> 1, k = 4 * c + 2 * (Y & 1) + (X & 1);
    const size_t pix = (size_t)i * W + j;
    return lrelu(__ldg(z2 + ((size_t)b * 4 * C + k) * H * W + pix), slope) + __ldg(x + ((size_t)b * C + (k % C)) * H * W + pix);
}

__global__ void __launch_bounds__(256) upsample_tail_fwd_kernel(const float* __restrict__ z2, const float* __restrict__ x, Taps3 f,
                                                                float* __restrict__ y, int B, int C, int H, int W, float slope) {
    const int H2 = 2 * H, W2 = 2 * W;
    const size_t n = (size_t)B * C * H2 * W2;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(t % W2), Y = (int)((t / W2) % H2), c = (int)((t / ((size_t)W2 * H2)) % C), b = (int)(t / ((size_t)W2 * H2 * C));
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int yy = reflect_idx(Y + a - 1, H2);
#pragma unroll
            for (int e = 0; e < 3; ++e) acc = fmaf(f.w[a] * f.w[e], shuffled(z2, x, b, c, yy, reflect_idx(X + e - 1, W2), C, H, W, slope), acc);
        }
        y[t] = acc;
    }
}

// adjoint of the reflect-padded 3-tap filter along one axis: positions P and weights with reflect(P + a) == Y
struct AdjTaps { int p[4]; float w[4]; int n; };
__device__ __forceinline__ AdjTaps adjoint_taps(int Y, int n, const Taps3& f) {
    AdjTaps t; t.n = 0;
    if (Y + 1 < n) { t.p[t.n] = Y + 1; t.w[t.n++] = f.w[0]; }     // a = -1
    t.p[t.n] = Y; t.w[t.n++] = f.w[1];                            // a = 0
    if (Y - 1 >= 0) { t.p[t.n] = Y - 1; t.w[t.n++] = f.w[2]; }    // a = +1
    if (Y == 1) { t.p[t.n] = 0; t.w[t.n++] = f.w[0]; }            // P + a = -1 reflects onto 1
    if (Y == n - 2) { t.p[t.n] = n - 1; t.w[t.n++] = f.w[2]; }    // P + a = n reflects onto n - 2
    return t;
}
__device__ __forceinline__ float blur_adjoint_at(const float* __restrict__ dy_plane, int Y, int X, int H2, int W2, const Taps3& f) {
    const AdjTaps ty = adjoint_taps(Y, H2, f), tx = adjoint_taps(X, W2, f);
    float acc = 0.f;
    for (int a = 0; a < ty.n; ++a)
        for (int e = 0; e < tx.n; ++e) acc = fmaf(ty.w[a] * tx.w[e], __ldg(dy_plane + (size_t)ty.p[a] * W2 + tx.p[e]), acc);
    return acc;
}

// one thread per element (b, m, i, j) of x: the four shuffled positions that received x[m] (k = m, m+C, m+2C, m+3C)
__global__ void __launch_bounds__(256) upsample_tail_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z2, Taps3 f,
                                                                float* __restrict__ dz2, float* __restrict__ dx, int B, int C, int H, int W, float slope) {
    const int H2 = 2 * H, W2 = 2 * W;
    const size_t n = (size_t)B * C * H * W;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % W), i = (int)((t / W) % H), m = (int)((t / ((size_t)W * H)) % C), b = (int)(t / ((size_t)W * H * C));
        float gx = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int k = m + r * C, c = k >> 2;
            const float ds = blur_adjoint_at(dy + ((size_t)b * C + c) * H2 * W2, 2 * i + ((k >> 1) & 1), 2 * j + (k & 1), H2, W2, f);
            const size_t zi = (((size_t)b * 4 * C + k) * H + i) * W + j;
            if (dz2) dz2[zi] = __ldg(z2 + zi) > 0.f ? ds : ds * slope;
            gx += ds;
        }
        if (dx) dx[t] = gx;
    }
}

// bilinear x2 (align_corners = False) source rows / weight of destination index Y
__device__ __forceinline__ void bilinear_src(int Y, int n_in, int* i0, int* i1, float* l) {
    float s = (Y + 0.5f) * 0.5f - 0.5f;
    s = s < 0.f ? 0.f : s;
    *i0 = (int)s;
    *i1 = *i0 + 1 < n_in ? *i0 + 1 : n_in - 1;
    *l = s - (float)*i0;
}
__device__ __forceinline__ float upsampled(const float* __restrict__ plane, int Y, int X, int H, int W) {
    int y0, y1, x0, x1; float ly, lx;
    bilinear_src(Y, H, &y0, &y1, &ly);
    bilinear_src(X, W, &x0, &x1, &lx);
    const float top = (1.f - lx) * __ldg(plane + (size_t)y0 * W + x0) + lx * __ldg(plane + (size_t)y0 * W + x1);
    const float bot = (1.f - lx) * __ldg(plane + (size_t)y1 * W + x0) + lx * __ldg(plane + (size_t)y1 * W + x1);
    return (1.f - ly) * top + ly * bot;
}

__global__ void __launch_bounds__(256) rgb_upsample_fwd_kernel(const float* __restrict__ x, Taps3 f, float* __restrict__ y, int planes, int H, int W) {
    const int H2 = 2 * H, W2 = 2 * W;
    const size_t n = (size_t)planes * H2 * W2;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(t % W2), Y = (int)((t / W2) % H2), p = (int)(t / ((size_t)W2 * H2));
        const float* plane = x + (size_t)p * H * W;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int yy = reflect_idx(Y + a - 1, H2);
#pragma unroll
            for (int e = 0; e < 3; ++e) acc = fmaf(f.w[a] * f.w[e], upsampled(plane, yy, reflect_idx(X + e - 1, W2), H, W), acc);
        }
        y[t] = acc;
    }
}

// one thread per up-sampled pixel: gradient through the blur (gather), then scattered onto its four bilinear sources
__global__ void __launch_bounds__(256) rgb_upsample_bwd_kernel(const float* __restrict__ dy, Taps3 f, float* __restrict__ dx, int planes, int H, int W) {
    const int H2 = 2 * H, W2 = 2 * W;
    const size_t n = (size_t)planes * H2 * W2;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(t % W2), Y = (int)((t / W2) % H2), p = (int)(t / ((size_t)W2 * H2));
        const float g = blur_adjoint_at(dy + (size_t)p * H2 * W2, Y, X, H2, W2, f);
        int y0, y1, x0, x1; float ly, lx;
        bilinear_src(Y, H, &y0, &y1, &ly);
        bilinear_src(X, W, &x0, &x1, &lx);
        float* plane = dx + (size_t)p * H * W;
        atomicAdd(plane + (size_t)y0 * W + x0, g * (1.f - ly) * (1.f - lx));
        atomicAdd(plane + (size_t)y0 * W + x1, g * (1.f - ly) * lx);
        atomicAdd(plane + (size_t)y1 * W + x0, g * ly * (1.f - lx));
        atomicAdd(plane + (size_t)y1 * W + x1, g * ly * lx);
    }
}

static int grid_for(size_t n) {
    size_t g = (n + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));       // a multiple of the SM count once the problem is large
}
static bool taps_from(const float* f3_host, Taps3* t) {
    const float s = f3_host[0] + f3_host[1] + f3_host[2];
    if (!(s != 0.f)) return false;
    for (int i = 0; i < 3; ++i) t->w[i] = f3_host[i] / s;
    return true;
}

}  // namespace hn

extern "C" int hn_upsample_tail_fwd(const float* z2, const float* x, const float* f3_host, float* y, int B, int C, int H, int W, void* stream) {
    using namespace hn;
    Taps3 t;
    if (!z2 || !x || !y || !f3_host || B <= 0 || C <= 0 || H < 2 || W < 2 || !taps_from(f3_host, &t)) return set_error(HN_E_BADARG, "hn_upsample_tail_fwd: bad argument");
    upsample_tail_fwd_kernel<<<grid_for((size_t)B * C * 4 * H * W), 256, 0, (cudaStream_t)stream>>>(z2, x, t, y, B, C, H, W, 0.2f);
    return check_launch("hn_upsample_tail_fwd");
}
extern "C" int hn_upsample_tail_bwd(const float* dy, const float* z2, const float* f3_host, float* dz2, float* dx, int B, int C, int H, int W, void* stream) {
    using namespace hn;
    Taps3 t;
    if (!dy || !z2 || !f3_host || B <= 0 || C <= 0 || H < 2 || W < 2 || !taps_from(f3_host, &t)) return set_error(HN_E_BADARG, "hn_upsample_tail_bwd: bad argument");
    upsample_tail_bwd_kernel<<<grid_for((size_t)B * C * H * W), 256, 0, (cudaStream_t)stream>>>(dy, z2, t, dz2, dx, B, C, H, W, 0.2f);
    return check_launch("hn_upsample_tail_bwd");
}
extern "C" int hn_rgb_upsample_fwd(const float* x, const float* f3_host, float* y, int planes, int H, int W, void* stream) {
    using namespace hn;
    Taps3 t;
    if (!x || !y || !f3_host || planes <= 0 || H < 2 || W < 2 || !taps_from(f3_host, &t)) return set_error(HN_E_BADARG, "hn_rgb_upsample_fwd: bad argument");
    rgb_upsample_fwd_kernel<<<grid_for((size_t)planes * 4 * H * W), 256, 0, (cudaStream_t)stream>>>(x, t, y, planes, H, W);
    return check_launch("hn_rgb_upsample_fwd");
}
extern "C" int hn_rgb_upsample_bwd(const float* dy, const float* f3_host, float* dx_zeroed, int planes, int H, int W, void* stream) {
    using namespace hn;
    Taps3 t;
    if (!dy || !dx_zeroed || !f3_host || planes <= 0 || H < 2 || W < 2 || !taps_from(f3_host, &t)) return set_error(HN_E_BADARG, "hn_rgb_upsample_bwd: bad argument");
    rgb_upsample_bwd_kernel<<<grid_for((size_t)planes * 4 * H * W), 256, 0, (cudaStream_t)stream>>>(dy, t, dx_zeroed, planes, H, W);
    return check_launch("hn_rgb_upsample_bwd");
}
