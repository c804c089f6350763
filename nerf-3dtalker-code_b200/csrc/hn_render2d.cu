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

// Forward tail, one CTA = one (item, channel) plane x a tile of 32 x 8 INPUT pixels (64 x 16 outputs), one thread per input pixel.
// An input pixel owns a 2 x 2 quad of the shuffled tensor s = pixel_shuffle(leaky_relu(z2) + repeat(x, 4)) (its four sub-channels
// 4c .. 4c+3): phase A builds the quads of the tile plus one quad of margin in shared memory (8 coalesced plane loads per quad,
// no per-element index arithmetic), phase B mirrors the rows / columns next to the image border (reflect padding: s[-1] = s[1],
// s[n] = s[n-2]), then every thread filters its own quad from a 4 x 4 window: 16 shared loads and 36 FMAs for four outputs,
// stored as two float2.
constexpr int kQW = 32, kQH = 8;                                   // input-pixel tile
constexpr int kSW = 2 * kQW + 4, kSH = 2 * kQH + 4;                // shared tile: outputs Y0 - 2 .. Y0 + 2 kQH + 1
__global__ void __launch_bounds__(256) upsample_tail_fwd_kernel(const float* __restrict__ z2, const float* __restrict__ x, Taps3 f,
                                                                float* __restrict__ y, int B, int C, int H, int W, float slope) {
    __shared__ float s[kSH][kSW + 1];
    const int H2 = 2 * H, W2 = 2 * W;
    const int tiles_x = (W + kQW - 1) / kQW;
    const int j0 = (blockIdx.x % tiles_x) * kQW, i0 = (blockIdx.x / tiles_x) * kQH;
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t HW = (size_t)H * W;
    const float* zq = z2 + ((size_t)b * 4 * C + 4 * c) * HW;       // sub-channel q at zq + q * HW
    const float* xb = x + (size_t)b * C * HW;
    int xm[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) xm[q] = (4 * c + q) % C;
    // phase A: quads (i0 - 1 + qi, j0 - 1 + qj), qi < kQH + 2, qj < kQW + 2
    for (int e = threadIdx.x; e < (kQH + 2) * (kQW + 2); e += 256) {
        const int qi = e / (kQW + 2), qj = e - qi * (kQW + 2);
        const int i = i0 - 1 + qi, j = j0 - 1 + qj;
        if (i < 0 || i >= H || j < 0 || j >= W) continue;
        const size_t pix = (size_t)i * W + j;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = lrelu(__ldg(zq + q * HW + pix), slope) + __ldg(xb + xm[q] * HW + pix);
        s[2 * qi][2 * qj] = v[0]; s[2 * qi][2 * qj + 1] = v[1];
        s[2 * qi + 1][2 * qj] = v[2]; s[2 * qi + 1][2 * qj + 1] = v[3];
    }
    __syncthreads();
    // phase B: reflect.  Shared row r holds output row Y0 - 2 + r (Y0 = 2 i0); rows first (all columns), then columns (all rows)
    const int Y0 = 2 * i0, X0 = 2 * j0;
    if (Y0 == 0)
        for (int t = threadIdx.x; t < kSW; t += 256) s[1][t] = s[3][t];                               // Y = -1 <- Y = 1
    if (H2 - Y0 + 2 < kSH && H2 - Y0 + 2 >= 2)
        for (int t = threadIdx.x; t < kSW; t += 256) s[H2 - Y0 + 2][t] = s[H2 - Y0][t];               // Y = H2 <- Y = H2 - 2
    __syncthreads();
    if (X0 == 0)
        for (int t = threadIdx.x; t < kSH; t += 256) s[t][1] = s[t][3];
    if (W2 - X0 + 2 < kSW && W2 - X0 + 2 >= 2)
        for (int t = threadIdx.x; t < kSH; t += 256) s[t][W2 - X0 + 2] = s[t][W2 - X0];
    __syncthreads();
    const int tj = threadIdx.x & 31, ti = threadIdx.x >> 5;
    const int i = i0 + ti, j = j0 + tj;
    if (i >= H || j >= W) return;
    float w[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int t = 0; t < 4; ++t) w[r][t] = s[2 * ti + 1 + r][2 * tj + 1 + t];
    float* plane = y + ((size_t)b * C + c) * H2 * W2;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        float o[2];
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            float acc = 0.f;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int e = 0; e < 3; ++e) acc = fmaf(f.w[a] * f.w[e], w[dy + a][dx + e], acc);
            o[dx] = acc;
        }
        *reinterpret_cast<float2*>(plane + (size_t)(2 * i + dy) * W2 + 2 * j) = make_float2(o[0], o[1]);
    }
}

// Backward: block = one (item, channel) plane x 64 x 16 output pixels (32 x 8 input pixels, one per thread, a full warp along
// the row).  The dy tile (+ halo, zero outside the image) goes to shared memory.  The adjoint of the reflect-padded filter is
// again a 3-tap filter whose outer weights pick up the reflected tap next to the border (position 1 also receives what row -1
// read, position n-2 what row n read), so ds needs no index lists: 9 shared-memory taps per output position.
// dz2 = ds * leaky_relu'(z2) for the pixel's four sub-channels; ds is added to the x channel it came from (x[m] feeds the four
// shuffled channels m, m+C, m+2C, m+3C; dx accumulates).
constexpr int kBX = 64, kBY = 16;
__device__ __forceinline__ void adjoint_weights(int Y, int n, const Taps3& f, float* cm, float* c0, float* cp) {
    *cm = f.w[2] + (Y == 1 ? f.w[0] : 0.f);          // coefficient of dy[Y - 1]
    *c0 = f.w[1];
    *cp = f.w[0] + (Y == n - 2 ? f.w[2] : 0.f);      // coefficient of dy[Y + 1]
}
// GROUPED = true (C % 4 == 0): blockIdx.y = c0 < C / 4 and the CTA walks the four planes c0 + r C / 4, whose sub-channels
// 4 c + q = 4 c0 + q + r C are exactly the four shuffled channels fed by x channel 4 c0 + q - so dx is owned by one thread and
// needs no atomics.  GROUPED = false: one plane per CTA, dx with atomics.
template <bool GROUPED>
__global__ void __launch_bounds__(256) upsample_tail_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z2, Taps3 f,
                                                                float* __restrict__ dz2, float* __restrict__ dx, int B, int C, int H, int W, float slope) {
    __shared__ float g[kBY + 2][kBX + 3];
    const int H2 = 2 * H, W2 = 2 * W;
    const int tiles_x = (W2 + kBX - 1) / kBX;
    const int X0 = (blockIdx.x % tiles_x) * kBX, Y0 = (blockIdx.x / tiles_x) * kBY;
    const int b = blockIdx.z;
    const int lj = threadIdx.x & 31, li = threadIdx.x >> 5;
    const int j = (X0 >> 1) + lj, i = (Y0 >> 1) + li;
    const bool inside = i < H && j < W;
    const size_t pix = (size_t)i * W + j;
    float wy[2][3], wx[2][3];                                 // adjoint weights of this thread's two output rows / columns
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        adjoint_weights(Y0 + 2 * li + d, H2, f, &wy[d][0], &wy[d][1], &wy[d][2]);
        adjoint_weights(X0 + 2 * lj + d, W2, f, &wx[d][0], &wx[d][1], &wx[d][2]);
    }
    float dxacc[4] = {0.f, 0.f, 0.f, 0.f};
    const int n_planes = GROUPED ? 4 : 1;
    const bool vec_rows = (W2 & 3) == 0 && X0 + kBX <= W2 && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
    for (int r = 0; r < n_planes; ++r) {
        const int c = GROUPED ? (int)blockIdx.y + r * (C >> 2) : (int)blockIdx.y;
        const float* plane = dy + ((size_t)b * C + c) * H2 * W2;
        if (r) __syncthreads();
        if (vec_rows) {                                       // 64 interior columns as float4 (one row = 16 lanes), the two halo columns apart
            for (int e = threadIdx.x; e < (kBY + 2) * (kBX / 4); e += 256) {
                const int sy = e / (kBX / 4), c4 = e % (kBX / 4);
                const int Y = Y0 - 1 + sy;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (Y >= 0 && Y < H2) v = __ldg(reinterpret_cast<const float4*>(plane + (size_t)Y * W2 + X0) + c4);
                g[sy][1 + 4 * c4] = v.x; g[sy][2 + 4 * c4] = v.y; g[sy][3 + 4 * c4] = v.z; g[sy][4 + 4 * c4] = v.w;
            }
            for (int e = threadIdx.x; e < 2 * (kBY + 2); e += 256) {
                const int sy = e >> 1, sx = (e & 1) ? kBX + 1 : 0;
                const int Y = Y0 - 1 + sy, X = X0 - 1 + sx;
                g[sy][sx] = (Y >= 0 && Y < H2 && X >= 0 && X < W2) ? __ldg(plane + (size_t)Y * W2 + X) : 0.f;
            }
        } else {
            for (int e = threadIdx.x; e < (kBY + 2) * (kBX + 2); e += 256) {
                const int sy = e / (kBX + 2), sx = e % (kBX + 2);
                const int Y = Y0 - 1 + sy, X = X0 - 1 + sx;
                g[sy][sx] = (Y >= 0 && Y < H2 && X >= 0 && X < W2) ? __ldg(plane + (size_t)Y * W2 + X) : 0.f;
            }
        }
        __syncthreads();
        if (!inside) continue;
        float win[4][4];                                      // dy at rows Y - 1 .. Y + 2, columns X - 1 .. X + 2 of this thread's quad origin
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int e = 0; e < 4; ++e) win[a][e] = g[2 * li + a][2 * lj + e];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int dyq = q >> 1, dxq = q & 1;
            const float r0 = wx[dxq][0] * win[dyq][dxq] + wx[dxq][1] * win[dyq][dxq + 1] + wx[dxq][2] * win[dyq][dxq + 2];
            const float r1 = wx[dxq][0] * win[dyq + 1][dxq] + wx[dxq][1] * win[dyq + 1][dxq + 1] + wx[dxq][2] * win[dyq + 1][dxq + 2];
            const float r2 = wx[dxq][0] * win[dyq + 2][dxq] + wx[dxq][1] * win[dyq + 2][dxq + 1] + wx[dxq][2] * win[dyq + 2][dxq + 2];
            const float ds = wy[dyq][0] * r0 + wy[dyq][1] * r1 + wy[dyq][2] * r2;
            const int k = 4 * c + q;
            if (dz2) {
                const size_t zi = ((size_t)b * 4 * C + k) * H * W + pix;
                dz2[zi] = __ldg(z2 + zi) > 0.f ? ds : ds * slope;
            }
            if (GROUPED) dxacc[q] += ds;
            else if (dx) atomicAdd(dx + ((size_t)b * C + (k % C)) * H * W + pix, ds);
        }
    }
    if (GROUPED && dx && inside) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float* d = dx + ((size_t)b * C + 4 * blockIdx.y + q) * H * W + pix;
            *d += dxacc[q];                                   // accumulate semantics without atomics: this thread owns the element
        }
    }
}

// Software-pipelined form of the grouped kernel for full-width, 16-byte aligned tiles (W2 % 64 == 0): while plane r is filtered,
// the dy tile and the z2 values of plane r + 1 are already in flight in registers (double-buffered shared tile, one barrier
// per plane).  The plain kernel above exposed two global round trips per plane - ncu: 22 of 33 stall cycles on the long scoreboard.
__global__ void __launch_bounds__(256) upsample_tail_bwd_pipe_kernel(const float* __restrict__ dy, const float* __restrict__ z2, Taps3 f,
                                                                     float* __restrict__ dz2, float* __restrict__ dx, int B, int C, int H, int W, float slope) {
    __shared__ float g[2][kBY + 2][kBX + 4];
    const int H2 = 2 * H, W2 = 2 * W;
    const int tiles_x = W2 / kBX;
    const int X0 = (blockIdx.x % tiles_x) * kBX, Y0 = (blockIdx.x / tiles_x) * kBY;
    const int b = blockIdx.z, tid = threadIdx.x;
    const int lj = tid & 31, li = tid >> 5;
    const int j = (X0 >> 1) + lj, i = (Y0 >> 1) + li;
    const bool inside = i < H && j < W;
    const size_t pix = (size_t)i * W + j, HW = (size_t)H * W;
    float wy[2][3], wx[2][3];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        adjoint_weights(Y0 + 2 * li + d, H2, f, &wy[d][0], &wy[d][1], &wy[d][2]);
        adjoint_weights(X0 + 2 * lj + d, W2, f, &wx[d][0], &wx[d][1], &wx[d][2]);
    }
    // this thread's share of a tile fill: float4 (row tid / 16, columns 4 (tid % 16) ..), for tid < 32 also rows 16 / 17, for tid < 36 a halo element
    const int sy0 = tid >> 4, c4 = tid & 15, sy1 = 16 + (tid >> 4), syh = tid >> 1, sxh = (tid & 1) ? kBX + 1 : 0;
    float4 f0, f1 = make_float4(0.f, 0.f, 0.f, 0.f);
    float fh = 0.f, zv[4] = {1.f, 1.f, 1.f, 1.f};
    auto load_plane = [&](int r) {
        const int c = (int)blockIdx.y + r * (C >> 2);
        const float* plane = dy + ((size_t)b * C + c) * H2 * W2;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        int Y = Y0 - 1 + sy0;
        f0 = (Y >= 0 && Y < H2) ? __ldg(reinterpret_cast<const float4*>(plane + (size_t)Y * W2 + X0) + c4) : zero;
        if (tid < 32) {
            Y = Y0 - 1 + sy1;
            f1 = (Y >= 0 && Y < H2) ? __ldg(reinterpret_cast<const float4*>(plane + (size_t)Y * W2 + X0) + c4) : zero;
        }
        if (tid < 2 * (kBY + 2)) {
            Y = Y0 - 1 + syh;
            const int X = X0 - 1 + sxh;
            fh = (Y >= 0 && Y < H2 && X >= 0 && X < W2) ? __ldg(plane + (size_t)Y * W2 + X) : 0.f;
        }
        if (inside && dz2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) zv[q] = __ldg(z2 + ((size_t)b * 4 * C + 4 * c + q) * HW + pix);
        }
    };
    load_plane(0);
    float dxacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float (*gb)[kBX + 4] = g[r & 1];
        gb[sy0][1 + 4 * c4] = f0.x; gb[sy0][2 + 4 * c4] = f0.y; gb[sy0][3 + 4 * c4] = f0.z; gb[sy0][4 + 4 * c4] = f0.w;
        if (tid < 32) { gb[sy1][1 + 4 * c4] = f1.x; gb[sy1][2 + 4 * c4] = f1.y; gb[sy1][3 + 4 * c4] = f1.z; gb[sy1][4 + 4 * c4] = f1.w; }
        if (tid < 2 * (kBY + 2)) gb[syh][sxh] = fh;
        float zc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) zc[q] = zv[q];
        const int c = (int)blockIdx.y + r * (C >> 2);
        if (r + 1 < 4) load_plane(r + 1);                     // travels while this plane is filtered
        __syncthreads();
        if (!inside) continue;
        float win[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int e = 0; e < 4; ++e) win[a][e] = gb[2 * li + a][2 * lj + e];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int dyq = q >> 1, dxq = q & 1;
            const float r0 = wx[dxq][0] * win[dyq][dxq] + wx[dxq][1] * win[dyq][dxq + 1] + wx[dxq][2] * win[dyq][dxq + 2];
            const float r1 = wx[dxq][0] * win[dyq + 1][dxq] + wx[dxq][1] * win[dyq + 1][dxq + 1] + wx[dxq][2] * win[dyq + 1][dxq + 2];
            const float r2 = wx[dxq][0] * win[dyq + 2][dxq] + wx[dxq][1] * win[dyq + 2][dxq + 1] + wx[dxq][2] * win[dyq + 2][dxq + 2];
            const float ds = wy[dyq][0] * r0 + wy[dyq][1] * r1 + wy[dyq][2] * r2;
            if (dz2) dz2[((size_t)b * 4 * C + 4 * c + q) * HW + pix] = zc[q] > 0.f ? ds : ds * slope;
            dxacc[q] += ds;
        }
    }
    if (dx && inside) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float* d = dx + ((size_t)b * C + 4 * blockIdx.y + q) * HW + pix;
            *d += dxacc[q];
        }
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
// blur(bilinear x2) along one axis is a fixed linear map with at most three input taps per output: output Y (input index
// i = Y / 2) reads inputs i - 1, i, i + 1.  w[t] = weight of input i - 1 + t: the three blur taps at the reflected positions
// Y - 1, Y, Y + 1, each spread over its two (clamped) bilinear sources.
__device__ __forceinline__ void up_weights(int Y, int n_in, const Taps3& f, float (&w)[3]) {
    const int i = Y >> 1;
    w[0] = w[1] = w[2] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        int i0, i1; float l;
        bilinear_src(reflect_idx(Y + a - 1, 2 * n_in), n_in, &i0, &i1, &l);
        const int t0 = i0 - (i - 1), t1 = i1 - (i - 1);
        const float a0 = f.w[a] * (1.f - l), a1 = f.w[a] * l;
#pragma unroll
        for (int t = 0; t < 3; ++t) w[t] += (t == t0 ? a0 : 0.f) + (t == t1 ? a1 : 0.f);
    }
}
__device__ __forceinline__ int clampi(int v, int n) { return v < 0 ? 0 : (v >= n ? n - 1 : v); }

// one thread per INPUT pixel: its 3 x 3 neighbourhood gives the 2 x 2 outputs above it
__global__ void __launch_bounds__(256) rgb_upsample_fwd_kernel(const float* __restrict__ x, Taps3 f, float* __restrict__ y, int planes, int H, int W) {
    const int W2 = 2 * W;
    const size_t n = (size_t)planes * H * W;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % W), i = (int)((t / W) % H), p = (int)(t / ((size_t)W * H));
        const float* plane = x + (size_t)p * H * W;
        float v[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int e = 0; e < 3; ++e) v[a][e] = __ldg(plane + (size_t)clampi(i - 1 + a, H) * W + clampi(j - 1 + e, W));
        float wy[2][3], wx[2][3];
        up_weights(2 * i, H, f, wy[0]); up_weights(2 * i + 1, H, f, wy[1]);
        up_weights(2 * j, W, f, wx[0]); up_weights(2 * j + 1, W, f, wx[1]);
        float* out = y + (size_t)p * 4 * H * W;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            float o[2];
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float acc = 0.f;
#pragma unroll
                for (int a = 0; a < 3; ++a) acc = fmaf(wy[dy][a], wx[dx][0] * v[a][0] + wx[dx][1] * v[a][1] + wx[dx][2] * v[a][2], acc);
                o[dx] = acc;
            }
            *reinterpret_cast<float2*>(out + (size_t)(2 * i + dy) * W2 + 2 * j) = make_float2(o[0], o[1]);
        }
    }
}

// adjoint as a gather, one thread per INPUT pixel (p, q): the outputs that read input row p are Y = 2 p - 2 .. 2 p + 3 (their
// input index Y / 2 is p - 1, p or p + 1), with weight up_weights(Y)[p - (Y / 2 - 1)]; same along the columns.  No atomics.
__device__ __forceinline__ void adjoint_up_weights(int p, int n_in, const Taps3& f, float (&c)[6]) {
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        const int Y = 2 * p - 2 + t;
        float w[3] = {0.f, 0.f, 0.f};
        if (Y >= 0 && Y < 2 * n_in) up_weights(Y, n_in, f, w);
        const int slot = p - ((Y >> 1) - 1);                   // 2, 2, 1, 1, 0, 0 for t = 0 .. 5
        c[t] = slot == 0 ? w[0] : (slot == 1 ? w[1] : w[2]);
    }
}
__global__ void __launch_bounds__(256) rgb_upsample_bwd_kernel(const float* __restrict__ dy, Taps3 f, float* __restrict__ dx, int planes, int H, int W) {
    const int H2 = 2 * H, W2 = 2 * W;
    const size_t n = (size_t)planes * H * W;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int q = (int)(t % W), p = (int)((t / W) % H), pl = (int)(t / ((size_t)W * H));
        float cy[6], cx[6];
        adjoint_up_weights(p, H, f, cy);
        adjoint_up_weights(q, W, f, cx);
        const float* g = dy + (size_t)pl * H2 * W2;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const int Y = 2 * p - 2 + a;
            if (Y < 0 || Y >= H2) continue;
            float row = 0.f;
#pragma unroll
            for (int e = 0; e < 6; ++e) {
                const int X = 2 * q - 2 + e;
                if (X >= 0 && X < W2) row = fmaf(cx[e], __ldg(g + (size_t)Y * W2 + X), row);
            }
            acc = fmaf(cy[a], row, acc);
        }
        dx[t] += acc;                                          // accumulate semantics; this thread owns the element
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
    const dim3 grid((unsigned)(((W + kQW - 1) / kQW) * ((H + kQH - 1) / kQH)), (unsigned)C, (unsigned)B);
    upsample_tail_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z2, x, t, y, B, C, H, W, 0.2f);
    return check_launch("hn_upsample_tail_fwd");
}
extern "C" int hn_upsample_tail_bwd(const float* dy, const float* z2, const float* f3_host, float* dz2, float* dx, int B, int C, int H, int W, void* stream) {
    using namespace hn;
    Taps3 t;
    if (!dy || !z2 || !f3_host || B <= 0 || C <= 0 || H < 2 || W < 2 || !taps_from(f3_host, &t)) return set_error(HN_E_BADARG, "hn_upsample_tail_bwd: bad argument");
    if (C > 65535 || B > 65535) return set_error(HN_E_UNSUPPORTED, "hn_upsample_tail_bwd: more than 65535 channels or items");
    const unsigned tiles = (unsigned)(((2 * W + kBX - 1) / kBX) * ((2 * H + kBY - 1) / kBY));
    if (C % 4 == 0 && (2 * W) % kBX == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0)
        upsample_tail_bwd_pipe_kernel<<<dim3(tiles, (unsigned)(C / 4), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(dy, z2, t, dz2, dx, B, C, H, W, 0.2f);
    else if (C % 4 == 0) upsample_tail_bwd_kernel<true><<<dim3(tiles, (unsigned)(C / 4), (unsigned)B), 256, 0, (cudaStream_t)stream>>>(dy, z2, t, dz2, dx, B, C, H, W, 0.2f);
    else upsample_tail_bwd_kernel<false><<<dim3(tiles, (unsigned)C, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(dy, z2, t, dz2, dx, B, C, H, W, 0.2f);
    return check_launch("hn_upsample_tail_bwd");
}
extern "C" int hn_rgb_upsample_fwd(const float* x, const float* f3_host, float* y, int planes, int H, int W, void* stream) {
    using namespace hn;
    Taps3 t;
    if (!x || !y || !f3_host || planes <= 0 || H < 2 || W < 2 || !taps_from(f3_host, &t)) return set_error(HN_E_BADARG, "hn_rgb_upsample_fwd: bad argument");
    rgb_upsample_fwd_kernel<<<grid_for((size_t)planes * H * W), 256, 0, (cudaStream_t)stream>>>(x, t, y, planes, H, W);
    return check_launch("hn_rgb_upsample_fwd");
}
extern "C" int hn_rgb_upsample_bwd(const float* dy, const float* f3_host, float* dx_zeroed, int planes, int H, int W, void* stream) {
    using namespace hn;
    Taps3 t;
    if (!dy || !dx_zeroed || !f3_host || planes <= 0 || H < 2 || W < 2 || !taps_from(f3_host, &t)) return set_error(HN_E_BADARG, "hn_rgb_upsample_bwd: bad argument");
    rgb_upsample_bwd_kernel<<<grid_for((size_t)planes * H * W), 256, 0, (cudaStream_t)stream>>>(dy, t, dx_zeroed, planes, H, W);
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
