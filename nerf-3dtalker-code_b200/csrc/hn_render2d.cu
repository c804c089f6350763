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

// Block = one (item, channel) plane x one tile of 32 x 32 output pixels (16 x 16 input pixels).  The shuffled tensor
// s = pixel_shuffle(leaky_relu(z2) + repeat(x, 4)) of the tile plus a one-pixel halo (reflected at the image border) is
// built ONCE in shared memory - 2.3 global loads per output instead of 18 - and the 3 x 3 filter runs from there.
constexpr int kT = 32;                                            // output tile edge
__global__ void __launch_bounds__(256) upsample_tail_fwd_kernel(const float* __restrict__ z2, const float* __restrict__ x, Taps3 f,
                                                                float* __restrict__ y, int B, int C, int H, int W, float slope) {
    __shared__ float s[kT + 2][kT + 3];
    const int H2 = 2 * H, W2 = 2 * W;
    const int tiles_x = (W2 + kT - 1) / kT;
    const int X0 = (blockIdx.x % tiles_x) * kT, Y0 = (blockIdx.x / tiles_x) * kT;
    const int c = blockIdx.y, b = blockIdx.z;
    for (int e = threadIdx.x; e < (kT + 2) * (kT + 2); e += 256) {
        const int sy = e / (kT + 2), sx = e % (kT + 2);
        const int Y = Y0 - 1 + sy, X = X0 - 1 + sx;
        float v = 0.f;
        if (Y <= H2 && X <= W2) v = shuffled(z2, x, b, c, reflect_idx(Y, H2), reflect_idx(X, W2), C, H, W, slope);   // (<=: row n - 1 needs the reflected row n)
        s[sy][sx] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty0 = threadIdx.x >> 5;
    float* plane = y + ((size_t)b * C + c) * H2 * W2;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int ty = ty0 + 8 * r, Y = Y0 + ty, X = X0 + tx;
        if (Y >= H2 || X >= W2) continue;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int e = 0; e < 3; ++e) acc = fmaf(f.w[a] * f.w[e], s[ty + a][tx + e], acc);
        plane[(size_t)Y * W2 + X] = acc;
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

// Backward: block = one (item, channel) plane x 64 x 16 output pixels (32 x 8 input pixels, one per thread, a full warp along
// the row).  The dy tile (+ halo, zero outside the image) goes to shared memory.  The adjoint of the reflect-padded filter is
// again a 3-tap filter whose outer weights pick up the reflected tap next to the border (position 1 also receives what row -1
// read, position n-2 what row n read), so ds needs no index lists: 9 shared-memory taps per output position.
// dz2 = ds * leaky_relu'(z2) for the pixel's four sub-channels; ds is added to the x channel it came from (x[m] feeds the four
// shuffled channels m, m+C, m+2C, m+3C: atomics, dx zero-initialised).
constexpr int kBX = 64, kBY = 16;
__device__ __forceinline__ void adjoint_weights(int Y, int n, const Taps3& f, float* cm, float* c0, float* cp) {
    *cm = f.w[2] + (Y == 1 ? f.w[0] : 0.f);          // coefficient of dy[Y - 1]
    *c0 = f.w[1];
    *cp = f.w[0] + (Y == n - 2 ? f.w[2] : 0.f);      // coefficient of dy[Y + 1]
}
__global__ void __launch_bounds__(256) upsample_tail_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z2, Taps3 f,
                                                                float* __restrict__ dz2, float* __restrict__ dx, int B, int C, int H, int W, float slope) {
    __shared__ float g[kBY + 2][kBX + 3];
    const int H2 = 2 * H, W2 = 2 * W;
    const int tiles_x = (W2 + kBX - 1) / kBX;
    const int X0 = (blockIdx.x % tiles_x) * kBX, Y0 = (blockIdx.x / tiles_x) * kBY;
    const int c = blockIdx.y, b = blockIdx.z;
    const float* plane = dy + ((size_t)b * C + c) * H2 * W2;
    for (int e = threadIdx.x; e < (kBY + 2) * (kBX + 2); e += 256) {
        const int sy = e / (kBX + 2), sx = e % (kBX + 2);
        const int Y = Y0 - 1 + sy, X = X0 - 1 + sx;
        g[sy][sx] = (Y >= 0 && Y < H2 && X >= 0 && X < W2) ? __ldg(plane + (size_t)Y * W2 + X) : 0.f;
    }
    __syncthreads();
    const int lj = threadIdx.x & 31, li = threadIdx.x >> 5;
    const int j = (X0 >> 1) + lj, i = (Y0 >> 1) + li;
    if (i >= H || j >= W) return;
    const size_t pix = (size_t)i * W + j;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ly = 2 * li + (q >> 1), lx = 2 * lj + (q & 1);     // position inside the tile; shared-memory index = +1
        float ym, y0, yp, xm, x0, xp;
        adjoint_weights(Y0 + ly, H2, f, &ym, &y0, &yp);
        adjoint_weights(X0 + lx, W2, f, &xm, &x0, &xp);
        const float r0 = xm * g[ly][lx] + x0 * g[ly][lx + 1] + xp * g[ly][lx + 2];
        const float r1 = xm * g[ly + 1][lx] + x0 * g[ly + 1][lx + 1] + xp * g[ly + 1][lx + 2];
        const float r2 = xm * g[ly + 2][lx] + x0 * g[ly + 2][lx + 1] + xp * g[ly + 2][lx + 2];
        const float ds = ym * r0 + y0 * r1 + yp * r2;
        const int k = 4 * c + q;
        if (dz2) {
            const size_t zi = ((size_t)b * 4 * C + k) * H * W + pix;
            dz2[zi] = __ldg(z2 + zi) > 0.f ? ds : ds * slope;
        }
        if (dx) atomicAdd(dx + ((size_t)b * C + (k % C)) * H * W + pix, ds);
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
    if (C > 65535 || B > 65535) return set_error(HN_E_UNSUPPORTED, "hn_upsample_tail_fwd: more than 65535 channels or items");
    const dim3 grid((unsigned)(((2 * W + kT - 1) / kT) * ((2 * H + kT - 1) / kT)), (unsigned)C, (unsigned)B);
    upsample_tail_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z2, x, t, y, B, C, H, W, 0.2f);
    return check_launch("hn_upsample_tail_fwd");
}
extern "C" int hn_upsample_tail_bwd(const float* dy, const float* z2, const float* f3_host, float* dz2, float* dx, int B, int C, int H, int W, void* stream) {
    using namespace hn;
    Taps3 t;
    if (!dy || !z2 || !f3_host || B <= 0 || C <= 0 || H < 2 || W < 2 || !taps_from(f3_host, &t)) return set_error(HN_E_BADARG, "hn_upsample_tail_bwd: bad argument");
    if (C > 65535 || B > 65535) return set_error(HN_E_UNSUPPORTED, "hn_upsample_tail_bwd: more than 65535 channels or items");
    const dim3 grid((unsigned)(((2 * W + kBX - 1) / kBX) * ((2 * H + kBY - 1) / kBY)), (unsigned)C, (unsigned)B);
    upsample_tail_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, z2, t, dz2, dx, B, C, H, W, 0.2f);
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

// ---------------------------------------------------------------------------------------------------------------------
// Merge (NetWorks/HeadNeRFNet.py:103-113): the composited ray-major features F [B, N_r, C] and bg_alpha [B, N_r]
// become the channel-major map the renderer consumes, merge[b, c, r] = F[b, r, c] + bg_alpha[b, r] * bg_featmap[c, r]
// (ray r <-> pixel (r / fs, r % fs)): one transposing kernel instead of permute-copy + multiply + add, and its adjoint.
namespace hn {

__global__ void __launch_bounds__(256) merge_fwd_kernel(const float* __restrict__ F, const float* __restrict__ bg, const float* __restrict__ bgfeat,
                                                        float* __restrict__ out, int n_rays, int C) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        tile[ty + 8 * k][tx] = (r < n_rays && c < C) ? __ldg(F + ((size_t)b * n_rays + r) * C + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        if (r < n_rays && c < C)
            out[((size_t)b * C + c) * n_rays + r] = fmaf(__ldg(bg + (size_t)b * n_rays + r), __ldg(bgfeat + (size_t)c * n_rays + r), tile[tx][ty + 8 * k]);
    }
}

// g [B, C, N_r] -> gF [B, N_r, C] (transpose), g_bg[b, r] += sum_c g * bg_featmap, g_bgfeat[c, r] += sum_b g * bg_alpha
__global__ void __launch_bounds__(256) merge_bwd_kernel(const float* __restrict__ g, const float* __restrict__ bg, const float* __restrict__ bgfeat,
                                                        float* __restrict__ gF, float* __restrict__ g_bg, float* __restrict__ g_bgfeat, int n_rays, int C) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    float part = 0.f;                                             // this thread's share of sum_c g * bg_featmap for ray r0 + tx
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;
        float v = 0.f;
        if (r < n_rays && c < C) {
            v = __ldg(g + ((size_t)b * C + c) * n_rays + r);
            part = fmaf(v, __ldg(bgfeat + (size_t)c * n_rays + r), part);
            if (g_bgfeat) atomicAdd(g_bgfeat + (size_t)c * n_rays + r, v * __ldg(bg + (size_t)b * n_rays + r));
        }
        tile[ty + 8 * k][tx] = v;                                 // [c][r]
    }
    __syncthreads();
    if (gF) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = r0 + ty + 8 * k, c = c0 + tx;
            if (r < n_rays && c < C) gF[((size_t)b * n_rays + r) * C + c] = tile[tx][ty + 8 * k];
        }
    }
    if (g_bg) {
        __syncthreads();
        tile[ty][tx] = part;                                      // 8 partial sums per ray
        __syncthreads();
        if (ty == 0 && r0 + tx < n_rays) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += tile[k][tx];
            atomicAdd(g_bg + (size_t)b * n_rays + r0 + tx, s);
        }
    }
}

}  // namespace hn

extern "C" int hn_merge_fwd(const float* F, const float* bg_alpha, const float* bg_featmap, float* merge, int B, int n_rays, int C, void* stream) {
    using namespace hn;
    if (!F || !bg_alpha || !bg_featmap || !merge || B <= 0 || n_rays <= 0 || C <= 0 || B > 65535) return set_error(HN_E_BADARG, "hn_merge_fwd: bad argument");
    merge_fwd_kernel<<<dim3((n_rays + 31) / 32, (C + 31) / 32, B), 256, 0, (cudaStream_t)stream>>>(F, bg_alpha, bg_featmap, merge, n_rays, C);
    return check_launch("hn_merge_fwd");
}
extern "C" int hn_merge_bwd(const float* g_merge, const float* bg_alpha, const float* bg_featmap, float* gF, float* g_bg_zeroed, float* g_bgfeat_zeroed,
                            int B, int n_rays, int C, void* stream) {
    using namespace hn;
    if (!g_merge || !bg_alpha || !bg_featmap || B <= 0 || n_rays <= 0 || C <= 0 || B > 65535) return set_error(HN_E_BADARG, "hn_merge_bwd: bad argument");
    merge_bwd_kernel<<<dim3((n_rays + 31) / 32, (C + 31) / 32, B), 256, 0, (cudaStream_t)stream>>>(g_merge, bg_alpha, bg_featmap, gF, g_bg_zeroed, g_bgfeat_zeroed, n_rays, C);
    return check_launch("hn_merge_bwd");
}
