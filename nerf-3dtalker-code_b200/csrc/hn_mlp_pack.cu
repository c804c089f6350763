// hn_mlp_pack.cu — hn_pack_weights: fp32 state-dict weights ([out,in] row-major, NetWorks/models.py:32-59)
// -> the half-precision weight-unit stream the fused kernels consume.  One unit = one 16 KiB operand image
// (<=128 rows x 64 cols, SWIZZLE_128B), units stored in exactly the order the kernels' MMA tables read them:
//   [ forward units (kFwdUnits) | data-gradient units (W^T, bwd.n_units) | tensor-memory data-gradient chain units (kBwdTUnits) ].
#include <mutex>
#include "hn_api.h"
#include "hn_mlp_sched.h"
#include "hn_tc.cuh"

namespace hn {

__constant__ PackOp c_pack[kFwdUnits + kBwdUnitsMax + kBwdTUnits];

struct PackArgs { const float* w[12]; int ld[12]; int l5_hidden_col; int n_units; };

// one CTA per unit; thread t handles 16-byte chunks (8 halves) of the image
__global__ void __launch_bounds__(256) pack_kernel(PackArgs a, uint8_t* packed) {
    const int u = blockIdx.x;
    const PackOp op = c_pack[u];
    const float* W = a.w[op.w_idx];
    const int ld = a.ld[op.w_idx];
    const int col0 = op.col0 + (op.l5_hidden ? a.l5_hidden_col : 0);
    uint8_t* dst = packed + (size_t)u * kUnitBytes;
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int r = i >> 3, c8 = (i & 7) * 8;
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c8 + 2 * h + e;
                float x = 0.f;
                if (r < op.valid_r && c < op.valid_c)
                    x = op.transposed ? __ldg(W + (size_t)(op.row0 + c) * ld + col0 + r)
                                      : __ldg(W + (size_t)(op.row0 + r) * ld + col0 + c);
                v[e] = fminf(fmaxf(x, -65504.f), 65504.f);
            }
            pk[h] = pack_h2(v[0], v[1]);
        }
        *reinterpret_cast<uint4*>(dst + image_offset(r, c8)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

// W_f = W_R1[:, :384] W_R0  (192 x 384, fp32): RGB_layer_0 has no activation, so the two layers are one linear map per sample
// (hn_mlp_sched.h).  One thread per element, 384-long dot product; runs once per weight version.
__global__ void __launch_bounds__(256) fuse_r0r1_kernel(const float* wr1, int ldr1, const float* wr0, int ldr0, float* wf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= HN_RGB1 * HN_HIDDEN) return;
    const int n = t / HN_HIDDEN, k = t % HN_HIDDEN;
    float acc = 0.f;
    for (int c = 0; c < HN_HIDDEN; ++c) acc = fmaf(__ldg(wr1 + (size_t)n * ldr1 + c), __ldg(wr0 + (size_t)c * ldr0 + k), acc);
    wf[t] = acc;
}

static std::mutex g_mu;
static bool g_uploaded[64] = {};

}  // namespace hn

static size_t packed_units_bytes() { return (size_t)(hn::kFwdUnits + hn::host_schedules().n_bwd_pack_units + hn::kBwdTUnits) * hn::kUnitBytes; }

extern "C" size_t hn_packed_weights_bytes(void) {
    return packed_units_bytes() + (size_t)HN_RGB1 * HN_HIDDEN * sizeof(float);      // + the fp32 scratch of the fused RGB_layer_1 x RGB_layer_0 matrix
}

extern "C" int hn_pack_weights(const hn_weights_t* w, void* packed, void* stream) {
    using namespace hn;
    if (!w || !packed) return set_error(HN_E_BADARG, "hn_pack_weights: null pointer");
    for (int i = 0; i < 12; ++i)
        if (!w->w[i] || w->ld[i] <= 0) return set_error(HN_E_BADARG, "hn_pack_weights: null weight pointer or bad leading dimension");
    if (w->ld[W_L0] < HN_PE + 1 || w->l5_hidden_col < HN_PE || w->l5_hidden_col + HN_HIDDEN > w->ld[W_L5] ||
        w->ld[W_R1] < HN_HIDDEN || w->ld[W_R0] < HN_HIDDEN || w->ld[W_R2] != HN_RGB1)
        return set_error(HN_E_UNSUPPORTED, "hn_pack_weights: layer shapes do not match fg_CD_predictor (hidden 384, feat 256)");
    const HostSchedules& hs = host_schedules();
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (dev < 64 && !g_uploaded[dev]) {
            cudaError_t e = cudaMemcpyToSymbol(c_pack, hs.fwd_pack, sizeof(PackOp) * kFwdUnits, 0);
            if (e == cudaSuccess)
                e = cudaMemcpyToSymbol(c_pack, hs.bwd_pack, sizeof(PackOp) * hs.n_bwd_pack_units, sizeof(PackOp) * kFwdUnits);
            if (e == cudaSuccess)
                e = cudaMemcpyToSymbol(c_pack, hs.bwdt_pack, sizeof(PackOp) * kBwdTUnits, sizeof(PackOp) * (kFwdUnits + hs.n_bwd_pack_units));
            if (e != cudaSuccess) return set_error((int)e, cudaGetErrorString(e));
            g_uploaded[dev] = true;
        }
    }
    float* wf = reinterpret_cast<float*>((uint8_t*)packed + packed_units_bytes());
    fuse_r0r1_kernel<<<(HN_RGB1 * HN_HIDDEN + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w->w[W_R1], w->ld[W_R1], w->w[W_R0], w->ld[W_R0], wf);
    if (int rc = check_launch("hn_pack_weights (RGB_layer_0 fold)")) return rc;
    PackArgs a;
    for (int i = 0; i < 12; ++i) { a.w[i] = w->w[i]; a.ld[i] = w->ld[i]; }
    a.w[W_R1] = wf; a.ld[W_R1] = HN_HIDDEN;                           // every RGB_layer_1 unit is cut from the fused matrix
    a.l5_hidden_col = w->l5_hidden_col;
    a.n_units = kFwdUnits + hs.n_bwd_pack_units + kBwdTUnits;
    pack_kernel<<<a.n_units, 256, 0, (cudaStream_t)stream>>>(a, (uint8_t*)packed);
    return check_launch("hn_pack_weights");
}
