// hn_fine.cu — hierarchical ("fine") resampling of a ray from its coarse compositing weights: NetWorks/utils.py:164-265
// (FineSample.forward + _calc_sample_points_by_zvals).  One warp per ray:
//   pdf / cdf of the interior coarse weights w[1 .. N_c-2]  (warp reduction + warp scan, two values per lane at N_c = 64)
//   inverse-CDF lookup of N_f + 1 uniforms  (searchsorted right=True by binary search in shared memory, linear interpolation
//   between the coarse bin centres, the reference's 1e-5 guards)
//   merge with the N_c coarse depths by a bitonic sort of the 256-slot padded key array (torch.sort of the concatenation)
//   depths -> z_dists (x ray_l) and sample points o + d * l * z, sample-major.
// part of HeadNeRF's API surface, dead in this reference (hier_sampling=False everywhere, and its
// fine pass is called with the wrong argument count, HeadNeRFNet.py:182-185) - so this is the standalone operator with parity
// against the reference's own FineSample module, not a second render pass.
#include "hn_api.h"
#include "hn_sample.cuh"

namespace hn {

constexpr int kFineSlots = 256;            // >= N_c + N_f + 1 (64 + 129 = 193)
constexpr int kFineWarps = 4;

__global__ void __launch_bounds__(kFineWarps * 32) fine_sample_kernel(const hn_fine_sample_t a) {
    __shared__ float s_cdf[kFineWarps][128];
    __shared__ float s_bins[kFineWarps][128];
    __shared__ float s_keys[kFineWarps][kFineSlots];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * kFineWarps + warp;
    if (ray >= a.n_rays_total) return;
    const int nc = a.n_coarse, nf1 = a.n_fine + 1, ni = nc - 2;           // ni interior weights, ni + 1 cdf entries / bins
    const float* w = a.weights + ray * nc;
    const float* z = a.zvals + ray * nc;
    float* cdf = s_cdf[warp];
    float* bins = s_bins[warp];
    float* keys = s_keys[warp];
    // ---- pdf = w / sum(w + 1e-5) over the interior samples; cdf = [0, cumsum(pdf)]   (utils.py:225-230)
    float part = 0.f;
    for (int i = lane; i < ni; i += 32) part += __ldg(w + 1 + i) + 1e-5f;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s);
    const float total = part;
    // inclusive scan in index order: lane l owns the contiguous run [l * per, (l + 1) * per)
    const int per = (ni + 31) / 32;
    float run = 0.f;
    float local[4];                                                         // per <= 4 (n_coarse <= 130)
    for (int k = 0; k < per; ++k) {
        const int i = lane * per + k;
        const float p = i < ni ? __ldg(w + 1 + i) / total : 0.f;
        run += p;
        local[k] = run;
    }
    float incl = run;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += up;
    }
    const float base = incl - run;
    if (lane == 0) cdf[0] = 0.f;
    for (int k = 0; k < per; ++k) {
        const int i = lane * per + k;
        if (i < ni) cdf[1 + i] = base + local[k];
    }
    for (int i = lane; i < ni + 1; i += 32) bins[i] = 0.5f * (__ldg(z + i + 1) + __ldg(z + i));      // utils.py:246
    for (int i = lane; i < nc; i += 32) keys[i] = __ldg(z + i);
    __syncwarp();
    // ---- inverse-CDF samples   (utils.py:232-254)
    for (int j = lane; j < nf1; j += 32) {
        const float u = a.uniform ? __ldg(a.uniform + ray * nf1 + j) : (float)j / (float)(nf1 - 1);     // linspace(0, 1, N_f + 1)
        int lo = 0, hi = ni + 1;                                            // searchsorted(right=True): number of cdf entries <= u
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
        }
        const int below = max(0, lo - 1), above = min(ni, lo);
        const float c0 = cdf[below], c1 = cdf[above];
        float denom = c1 - c0;
        if (denom < 1e-5f) denom = 1.f;
        const float t = (u - c0) / denom;
        const float b0 = bins[below], b1 = bins[above];
        keys[nc + j] = b0 + t * (b1 - b0);
    }
    for (int i = nc + nf1 + lane; i < kFineSlots; i += 32) keys[i] = __int_as_float(0x7f800000);          // +inf padding sorts last
    __syncwarp();
    // ---- sort(cat(coarse, fine))   (utils.py:256): bitonic network over 256 slots, 4 compare-exchanges per lane and step
    for (int k = 2; k <= kFineSlots; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < kFineSlots / 2; t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));       // lower index of the pair
                const int l = i | j;
                const bool up = (i & k) == 0;
                const float x = keys[i], y = keys[l];
                if ((x > y) == up) { keys[i] = y; keys[l] = x; }
            }
            __syncwarp();
        }
    }
    // ---- depths -> outputs   (utils.py:172-211): the last sorted depth only closes the last interval
    const int np = nc + nf1 - 1;
    const int64_t b = ray / a.n_rays, r = ray % a.n_rays;
    const float ox = __ldg(a.ray_o + (b * 3 + 0) * a.n_rays + r), oy = __ldg(a.ray_o + (b * 3 + 1) * a.n_rays + r), oz = __ldg(a.ray_o + (b * 3 + 2) * a.n_rays + r);
    const float dx = __ldg(a.ray_d + (b * 3 + 0) * a.n_rays + r), dy = __ldg(a.ray_d + (b * 3 + 1) * a.n_rays + r), dz = __ldg(a.ray_d + (b * 3 + 2) * a.n_rays + r);
    const float l = __ldg(a.ray_l + ray);
    for (int i = lane; i < np; i += 32) {
        const float zi = keys[i], zn = keys[i + 1];
        a.out_zvals[ray * np + i] = zi;
        a.out_zdists[ray * np + i] = (zn - zi) * l;
        if (a.out_pts) {
            float* p = a.out_pts + (ray * np + i) * 3;
            p[0] = ox + dx * l * zi; p[1] = oy + dy * l * zi; p[2] = oz + dz * l * zi;
        }
    }
}

}  // namespace hn

extern "C" int hn_fine_sample(const hn_fine_sample_t* a, void* stream) {
    using namespace hn;
    if (!a || !a->weights || !a->zvals || !a->ray_o || !a->ray_d || !a->ray_l || !a->out_zvals || !a->out_zdists)
        return set_error(HN_E_BADARG, "hn_fine_sample: null pointer");
    if (a->n_rays <= 0 || a->n_rays_total <= 0 || a->n_rays_total % a->n_rays != 0) return set_error(HN_E_BADARG, "hn_fine_sample: ray counts");
    if (a->n_coarse < 4 || a->n_coarse > 130 || a->n_fine < 1 || a->n_coarse + a->n_fine + 1 > kFineSlots)
        return set_error(HN_E_UNSUPPORTED, "hn_fine_sample: 4 <= n_coarse <= 130 and n_coarse + n_fine + 1 <= 256");
    const unsigned grid = (unsigned)((a->n_rays_total + kFineWarps - 1) / kFineWarps);
    fine_sample_kernel<<<grid, kFineWarps * 32, 0, (cudaStream_t)stream>>>(*a);
    return check_launch("hn_fine_sample");
}
