/* headnerf_b200.h — C ABI of libheadnerf_b200.so: the sm_100a implementation of the HeadNeRF rendering
 * hot path (ray sampling -> positional encoding -> fg_CD_predictor MLP -> alpha compositing).
 *
 * The reference (NeRF-3DTalker) is pure Python/PyTorch and has no FFI of its own; each entry point
 * below replaces the reference code cited beside it (paths relative to the reference root) and is
 * bound from Python with ctypes (see INTEGRATION.md and nerf-3dtalker-code_b200/_lib.py).
 *
 * Conventions
 *  - plain C types only; every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the caller owns all memory (inputs, outputs, saved-for-backward, workspace); the library never
 *    synchronises and launches on the stream passed as `stream` (a cudaStream_t).  Its only own device
 *    resources, created once per device on first use and never freed: immutable schedule tables (constant
 *    memory, plus one < 64 KiB table for hn_pack_weights_precise) and, for hn_mlp_bwd_weights, one
 *    non-blocking side stream with two events that is forked from and joined back into `stream`
 *    (thread safety: calls on one device are serialised by a library mutex across that fork / launch / join
 *    sequence, so host threads may call concurrently on different streams);
 *  - every function returns 0 on success, a negative HN_E_* code on a bad argument, or a positive
 *    cudaError_t value; hn_last_error() returns a thread-local human-readable message;
 *  - samples are indexed m = (b*n_rays + r)*n_samples + s; M = B*n_rays*n_samples must be a multiple
 *    of 128 and n_samples one of 32/64/128 (whole rays per 128-sample tile, tiles never straddle items);
 *  - "image" buffers hold half-precision operand blocks of 128 rows x 64 columns = 16 KiB in the
 *    tensor-core SWIZZLE_128B layout (csrc/hn_tc.cuh); block (slot, tile, kb) of a saved tensor lives at
 *    byte ((slot_base + kb) * n_tiles + tile) * 16384.
 */
#ifndef HEADNERF_B200_H_
#define HEADNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HN_ABI_VERSION 4

/* error codes (negative = argument / configuration errors) */
#define HN_OK 0
#define HN_E_BADARG (-1)      /* null pointer, non-positive size, misaligned buffer                */
#define HN_E_UNSUPPORTED (-2) /* dimensions outside what the sm_100a kernels are specialised for    */
#define HN_E_PROTOCOL (-3)    /* a kernel's internal pipeline timed out (reported via status word)  */

/* Fixed architecture of fg_CD_predictor (NetWorks/models.py:29-59, HeadNeRFOptions.py:20-29). */
#define HN_HIDDEN 384     /* mlp_hidden_nchannels                              */
#define HN_FEAT 256       /* featmap_nc                                        */
#define HN_RGB1 192       /* HIDDEN/2, output of RGB_layer_1                   */
#define HN_PE 63          /* 3 + 2*3*10 positional-encoding channels           */
#define HN_PE_PAD 64
#define HN_TILE 128       /* samples per tile                                  */

/* Per-batch-item effective biases (latent codes folded in): one row of
 * HN_BIAS_STRIDE floats per batch item, laid out as
 *   [0,3072)    FeaExt_module_0..7   (8 x 384)
 *   [3072,3456) RGB_layer_0          (384)
 *   [3456,3648) RGB_layer_1          (192)
 *   [3648,3904) RGB_layer_2          (256)
 *   [3904]      density_module bias  (1)            rest: padding                                    */
#define HN_BIAS_OFF_R0 3072
#define HN_BIAS_OFF_R1 3456
#define HN_BIAS_OFF_R2 3648
#define HN_BIAS_OFF_DENSITY 3904
#define HN_BIAS_STRIDE 3920

/* Saved-activation slots (forward -> backward), in units of 64-column blocks. */
#define HN_SLOT_PE 0      /* 1 block : positional encoding (63 ch + zero pad)                        */
#define HN_SLOT_H0 1      /* 8 x 6 blocks : outputs of FeaExt_module_0..7 (post-ReLU)                */
#define HN_SLOT_R0 49     /* 6 blocks : output of RGB_layer_0 (no activation)                        */
#define HN_SLOT_X 55      /* 3 blocks : output of RGB_layer_1 (post-ReLU)                            */
#define HN_ACT_BLOCKS 58
/* Gradient slots written by hn_mlp_bwd_data, read by hn_mlp_bwd_weights (pre-activation gradients). */
#define HN_GSLOT_Z0 0     /* 8 x 6 blocks : dL/d(pre-act) of FeaExt_module_0..7                      */
#define HN_GSLOT_R0 48    /* 6 blocks : RGB_layer_0                                                  */
#define HN_GSLOT_R1 54    /* 3 blocks : RGB_layer_1                                                  */
#define HN_GSLOT_DENS 57  /* 1 block  : density head, column 0 = dL/d(pre-ReLU density), rest zero    */
#define HN_GRAD_BLOCKS 59 /* + 1 never-written pad block (rows of it only reach ignored outputs)      */
/* ReLU sign masks: 9 masked layers (FeaExt 0..7: 12 words/sample, RGB_layer_1: 6 words/sample). */
#define HN_MASK_WORDS 104 /* 8*12 + 6 uint32 per sample, padded so a row is a multiple of 16 bytes      */

int hn_abi_version(void);
const char* hn_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Weight packing.  Replaces nothing in the reference (cuDNN consumes the fp32 [out,in,1,1] tensors
 * directly, NetWorks/models.py:32-59); produces the half-precision operand images the kernels stream.
 * `w` holds 12 device pointers to the fp32 state-dict tensors in this order:
 *   FeaExt_module_0..7.weight, density_module.weight, RGB_layer_0.weight, RGB_layer_1.weight,
 *   RGB_layer_2.weight,   with row strides (= in_channels) in `ld`.  Column blocks that multiply
 * per-item constants (shape/audio/appearance codes) are NOT packed: they are folded into the bias.
 * `pe_col`/`h_col` give the column offset of the PE block and of the hidden block inside layers 0/5
 * and RGB_layer_1 (reference concat orders).                                           */
typedef struct {
    const float* w[12];
    int ld[12];
    int l5_hidden_col;   /* column of the h block in FeaExt_module_5.weight (242 without gaze)        */
} hn_weights_t;

size_t hn_packed_weights_bytes(void);
int hn_pack_weights(const hn_weights_t* w, void* packed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ray generation + stratified sampling (standalone form).   NetWorks/utils.py:64-161
 * Outputs are per sample, sample-major.  Any output pointer may be NULL.                            */
typedef struct {
    int B, n_rays, n_samples;
    float world_z1, world_z2;
    const float* xy;        /* [B,2,n_rays]  (reference layout)                                      */
    const float* Rmats;     /* [B,3,3] camera-to-world rotation                                      */
    const float* Tvecs;     /* [B,3]   camera-to-world translation                                   */
    const float* inv_inmats;/* [B,3,3]                                                               */
    const float* t_rand;    /* [B,n_rays,n_samples+1] uniform draws, or NULL for mode "test"          */
} hn_camera_t;

int hn_sample_rays(const hn_camera_t* cam, float* pts /*[M,3]*/, float* zvals /*[M]*/, float* z_dists /*[M]*/,
                   float* ray_d /*[B*n_rays,3]*/, float* ray_l /*[B*n_rays]*/, void* stream);

/* Backward of the ray set-up (autograd of NetWorks/utils.py:147-158): the per-ray gradients produced by hn_mlp_bwd_data /
 * hn_mlp_bwd_data_precise (w.r.t. ray origin, ray_d * ray_l and ray_l) -> dL/dRmats [B,3,3], dL/dTvecs [B,3],
 * dL/dinv_inmats [B,3,3], accumulated (+=, caller zero-initialises); any output may be NULL.              */
int hn_camera_bwd(const hn_camera_t* cam, const float* g_ray_o /*[B*n_rays,3]*/, const float* g_ray_v /*[B*n_rays,3]*/,
                  const float* g_ray_l /*[B*n_rays]*/, float* dR, float* dT, float* dKinv, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused sampling + positional encoding + MLP forward.
 * NetWorks/utils.py:43-51,147-161 ; NetWorks/HeadNeRFNet.py:139-152,84-95 ; NetWorks/models.py:62-87  */
typedef struct {
    hn_camera_t cam;
    const float* bias;        /* [B, HN_BIAS_STRIDE] effective biases                                 */
    const float* w_density;   /* [384] fp32 density_module.weight                                     */
    const void* packed;       /* from hn_pack_weights                                                 */
    float* feat;              /* [M,256] out (NULL allowed when `xbar` is set)                        */
    float* sigma;             /* [M] out: relu(density)                                               */
    float* delta;             /* [M] out: z_dists                                                     */
    float* zvals;             /* [M] out or NULL                                                      */
    void* act;                /* [HN_ACT_BLOCKS, n_tiles, 16 KiB] out or NULL (no backward)            */
    uint32_t* masks;          /* [M, HN_MASK_WORDS] out or NULL                                       */
    int* status;              /* device int, zero-initialised by caller; nonzero = pipeline fault     */
} hn_mlp_fwd_t;

int hn_mlp_fwd(const hn_mlp_fwd_t* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Alpha compositing, forward and backward.                  NetWorks/utils.py:273-309
 * feat is sample-major [M,C]; one warp per ray.  C must be a multiple of 128 (<= 256).               */
typedef struct {
    int n_rays_total, n_samples, C;
    const float* feat;   /* [M,C] */
    const float* sigma;  /* [M]   */
    const float* delta;  /* [M]   */
    const float* zvals;  /* [M] or NULL (depth not computed) */
    float* F;            /* [n_rays_total, C] out */
    float* bg_alpha;     /* [n_rays_total]    out */
    float* depth;        /* [n_rays_total]    out or NULL */
    float* weights;      /* [M]               out or NULL */
} hn_composite_fwd_t;

int hn_composite_fwd(const hn_composite_fwd_t* a, void* stream);

typedef struct {
    int n_rays_total, n_samples, C;
    const float* feat; const float* sigma; const float* delta; const float* zvals;
    const float* gF;        /* [n_rays_total, C] */
    const float* g_bg;      /* [n_rays_total]    */
    const float* g_depth;   /* [n_rays_total] or NULL */
    float* dfeat;           /* [M,C] fp32 out, or NULL                                               */
    void* dfeat_image;      /* half-precision image out [C/64, n_tiles, 16 KiB], scaled by *grad_scale, or NULL */
    const float* grad_scale;/* device scalar (power of two) used for dfeat_image                      */
    float* dsigma;          /* [M] out  (w.r.t. the post-ReLU density)                                */
    float* ddelta;          /* [M] out or NULL                                                        */
} hn_composite_bwd_t;

int hn_composite_bwd(const hn_composite_bwd_t* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MLP backward, data path: dL/dfeat, dL/dsigma -> pre-activation gradients of every layer (saved as
 * images for the weight pass), dL/d(positional encoding) reduced to per-ray camera-side gradients.
 * Autograd of NetWorks/models.py:62-87 and NetWorks/utils.py:43-51,80-86.                            */
typedef struct {
    hn_camera_t cam;
    const void* packed;
    const float* w_density;
    const void* dfeat_image;  /* from hn_composite_bwd                                                */
    const float* dsigma;      /* [M] fp32, unscaled                                                   */
    const float* ddelta;      /* [M] fp32, unscaled, or NULL                                          */
    const float* sigma;       /* [M] saved forward output (ReLU mask of the density head)             */
    const float* grad_scale;  /* device scalar                                                        */
    const uint32_t* masks;    /* saved by hn_mlp_fwd                                                  */
    const void* act;          /* saved by hn_mlp_fwd (PE block is re-used for the sampling gradient)   */
    void* grads;              /* [HN_GRAD_BLOCKS, n_tiles, 16 KiB] out                                 */
    float* g_ray_o;           /* [B*n_rays,3] out or NULL: dL/d ray origin                             */
    float* g_ray_v;           /* [B*n_rays,3] out or NULL: dL/d (ray_d * ray_l)                        */
    float* g_ray_l;           /* [B*n_rays]   out or NULL: dL/d ray_l through z_dists                  */
    int* status;
} hn_mlp_bwd_data_t;

/* (SURVEY.md section 8b lists a separate hn_sample_pe_bwd: it is folded in here - with g_ray_* given this kernel also runs the
 * positional-encoding backward and the per-ray sampling reductions, and hn_camera_bwd finishes the chain to R / T / K^-1.)      */
int hn_mlp_bwd_data(const hn_mlp_bwd_data_t* a, void* stream);

/* MLP backward, weight path: contraction over all samples of (pre-activation gradient)^T x (layer input).
 * Accumulates (atomically, caller zero-initialises) into fp32 gradient tensors shaped like the
 * reference state dict, and into d(bias) [B, HN_BIAS_STRIDE].                                         */
typedef struct {
    int B, n_rays, n_samples;
    const void* act;          /* saved by hn_mlp_fwd                                                   */
    const void* grads;        /* saved by hn_mlp_bwd_data                                              */
    const void* dfeat_image;  /* from hn_composite_bwd (pre-activation gradient of RGB_layer_2)        */
    const float* grad_scale;
    float* dw[12];            /* same order / leading dimensions as hn_weights_t; NULL entries skipped; */
    int ld[12];               /* all NULL = only the bias gradients of the latent-folded layers         */
    int l5_hidden_col;
    float* dbias;             /* [B, HN_BIAS_STRIDE]                                                  */
    void* items_workspace;    /* device scratch for the work-item table, hn_wgrad_workspace_bytes(B)   */
    size_t items_workspace_bytes;
    int* status;
    int want_all_bias;        /* with every dw NULL: 0 = bias gradients of the latent-folded layers only (FeaExt_module_0,
                               * _5, RGB_layer_1: what the code gradients need), 1 = of all 12 layers (bias-only fine-tuning) */
    int r0_fused;             /* 1: `act` / `grads` come from the fast chains (no RGB_layer_0: RGB_layer_1's layer input is
                               * FeaExt_module_7's output and its weight gradient is dL/dW_f -> `dwf`; dw[9] is not written,
                               * dw[10] only marks that the gradient is wanted); 0: from hn_mlp_bwd_data_precise (12 layers)     */
    float* dwf;               /* r0_fused: [192, 384] zero-initialised, accumulated; NULL = RGB_layer_1 / _0 gradients not wanted */
    void* det_workspace;      /* NULL: sample-range partial sums meet in dw / dbias through atomic adds (summation order, and the
                               * last bits, vary from run to run).  Non-NULL (hn_wgrad_det_workspace_bytes(B) bytes): every work
                               * item writes a private slice and a second kernel adds the slices in a fixed order - run-to-run
                               * bit-identical gradients (what the reference asks of cuDNN with cudnn.deterministic, train.py:26-29) */
    size_t det_workspace_bytes;
} hn_mlp_bwd_weights_t;

size_t hn_wgrad_workspace_bytes(int B);
size_t hn_wgrad_det_workspace_bytes(int B);
int hn_mlp_bwd_weights(const hn_mlp_bwd_weights_t* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Latent-code folding.  Replaces the broadcast + concat of the codes to every sample
 * (NetWorks/HeadNeRFNet.py:84-89,149-152; NetWorks/models.py:69,75,80): the codes' weight columns become one
 * effective bias row per batch item (layout above); the backward maps the bias-row gradient produced by
 * hn_mlp_bwd_weights / hn_mlp_bwd_data_precise to the gradients of the codes, of the folded weight columns
 * (accumulated, +=, into full-size weight-gradient tensors) and of the 12 bias vectors (+=).              */
typedef struct {
    int B, shape_dims, appea_dims;    /* audio-style code is 64 wide (NetWorks/models.py:32)                */
    const float* w0;  int ld0;        /* FeaExt_module_0.weight [384, 63 + shape_dims + 64]                 */
    const float* w5;  int ld5;        /* FeaExt_module_5.weight [384, 63 + shape_dims + 384]                */
    const float* wr1; int ldr1;       /* RGB_layer_1.weight     [192, 384 + appea_dims]                     */
    const float* bias[12];            /* bias vectors in hn_weights_t order ([8] = density_module.bias)     */
    const float* shape_code;          /* [B, shape_dims]                                                    */
    const float* audio;               /* [B, 64]                                                            */
    const float* appea;               /* [B, appea_dims]                                                    */
    int r0_fused;                     /* 1 (the fast kernels): RGB_layer_1's row also carries W_R1[:, :384] b_R0 - those chains
                                       * skip RGB_layer_0, which has no activation and is multiplied into RGB_layer_1 by
                                       * hn_pack_weights; 0 (hn_mlp_fwd_precise, which runs the twelve layers one by one)       */
    const float* wr0; int ldr0;       /* RGB_layer_0.weight [384, 384]: read by hn_render_bwd only (may be NULL elsewhere)       */
} hn_fold_t;

typedef struct {
    float* dshape; float* daudio; float* dappea;   /* written (=); NULL = not needed                       */
    float* dw0; float* dw5; float* dwr1;           /* latent columns accumulated (+=); NULL = not needed    */
    float* dbias[12];                              /* accumulated (+=); NULL = not needed                   */
} hn_fold_grads_t;

int hn_fold_bias(const hn_fold_t* a, float* bias_eff /*[B, HN_BIAS_STRIDE]*/, void* stream);

/* Backward of the RGB_layer_0 fold (NetWorks/models.py:79-80 has no activation between RGB_layer_0 and RGB_layer_1, so the fast
 * kernels run them as one matrix W_f = W_R1[:, :384] W_R0): hn_mlp_bwd_weights (r0_fused = 1) delivers dL/dW_f [192,384] and the
 * bias-row gradient; this turns them into dL/dW_R1[:, :384] (+=), dL/dW_R0 (+=) and - written into RGB_layer_0's entries of item
 * 0's bias-row gradient, where hn_fold_bias_bwd finds it - dL/db_R0.  Call between hn_mlp_bwd_weights and hn_fold_bias_bwd.   */
typedef struct {
    int B;
    const float* wr0; int ldr0;       /* RGB_layer_0.weight [384, 384]                                       */
    const float* wr1; int ldr1;       /* RGB_layer_1.weight [192, 384 + appea_dims]                          */
    const float* b_r0;                /* RGB_layer_0.bias [384]                                              */
    const float* dwf;                 /* [192, 384] from hn_mlp_bwd_weights                                  */
    float* dbias_eff;                 /* [B, HN_BIAS_STRIDE] in / out                                        */
    float* dwr0;                      /* += [384, ldr0], or NULL                                             */
    float* dwr1;                      /* += hidden columns of [192, ldr1], or NULL                           */
} hn_unfuse_t;
int hn_unfuse_r0r1(const hn_unfuse_t* a, void* stream);
int hn_fold_bias_bwd(const hn_fold_t* a, const float* dbias_eff /*[B, HN_BIAS_STRIDE]*/, const hn_fold_grads_t* g, void* stream);

/* Power-of-two loss scale of the half-precision backward chain: *scale_out = 2^floor(log2(target / max|g|)),
 * computed on the device (no host sync).  scratch8 = 8 bytes, zero before the first call (the kernel leaves
 * them zero again, so one buffer per stream can be reused without re-clearing).                           */
int hn_loss_scale(const float* g, int64_t n, float target, float* scale_out, void* scratch8, void* stream);

/* ------------------------------------------------------------------------------------------------
 * High-precision mode ("fp32-accumulate mode" with split operands): the same chain, layer by layer, with
 * fp32 activations in HBM and every tensor-core product computed as hi*hi + lo*hi + hi*lo over
 * half-precision hi/lo splits of BOTH operands (~22 significant bits; csrc/hn_precise.cu).
 * Replaces the same reference code as hn_mlp_fwd / hn_mlp_bwd_data (NetWorks/models.py:62-87,
 * NetWorks/utils.py:43-51,147-161 and their autograd); used when the feature-map tolerance must hold at
 * any activation scale and for the ill-conditioned camera gradients.
 * Workspaces are fp32, sample-major: `acts` = [PE 64 | FeaExt_module_0..7 8x384 | RGB_layer_0 384 |
 * RGB_layer_1 192] floats per sample, stored buffer after buffer (buffer k at acts + M*offset_k, row
 * stride = its width); `gz` = [dZ FeaExt_module_0..7 8x384 | dZ RGB_layer_0 384 | dZ RGB_layer_1 192 |
 * dL/dPE 64], loss-scaled by *grad_scale.  Both have hn_precise_workspace_floats(M) elements.          */
size_t hn_precise_packed_bytes(void);
size_t hn_precise_workspace_floats(int64_t M);
int hn_pack_weights_precise(const hn_weights_t* w, void* packed_hl, void* stream);

typedef struct {
    hn_camera_t cam;
    const float* bias;        /* [B, HN_BIAS_STRIDE] effective biases                                 */
    const float* w_density;   /* [384]                                                                */
    const void* packed_hl;    /* from hn_pack_weights_precise                                         */
    float* feat;              /* [M,256] out                                                          */
    float* sigma;             /* [M] out                                                              */
    float* delta;             /* [M] out                                                              */
    float* zvals;             /* [M] out or NULL                                                      */
    float* acts;              /* workspace / saved activations (see above)                            */
    int* status;
} hn_mlp_fwd_precise_t;

int hn_mlp_fwd_precise(const hn_mlp_fwd_precise_t* a, void* stream);

typedef struct {
    hn_camera_t cam;
    const void* packed_hl;
    const float* w_density;
    const float* dfeat;       /* [M,256] fp32, unscaled (hn_composite_bwd)                            */
    const float* dsigma;      /* [M] unscaled                                                         */
    const float* ddelta;      /* [M] unscaled or NULL                                                 */
    const float* sigma;       /* [M] saved forward output                                             */
    const float* grad_scale;  /* device scalar (power of two)                                         */
    const float* acts;        /* saved by hn_mlp_fwd_precise                                          */
    float* gz;                /* out: pre-activation gradients (see above)                            */
    float* g_ray_o;           /* [B*n_rays,3] accumulated, or NULL (no camera gradients)              */
    float* g_ray_v;
    float* g_ray_l;
    float* dbias;             /* [B, HN_BIAS_STRIDE] accumulated (caller zero-initialises), or NULL   */
    void* act_image;          /* optional: half-precision operand images of acts / gz / dfeat in the  */
    void* grads_image;        /* layouts hn_mlp_bwd_weights reads (all three or none)                 */
    void* dfeat_image;
    int* status;
} hn_mlp_bwd_data_precise_t;

int hn_mlp_bwd_data_precise(const hn_mlp_bwd_data_precise_t* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Consumer side (SURVEY.md section 8f, row 1 - first pieces): the memory-bound tails of NeuralRenderer's up-sampling
 * blocks as single kernels, NCHW fp32 like the reference modules; the 1x1 convolutions around them stay library GEMMs.
 * `f3_host` = the three taps of the blur buffer (`blur_layer.f` / `rgb_upsample.1.f`, normally [1,2,1]) on the HOST.
 *   hn_upsample_tail_*  : y = blur(pixel_shuffle(leaky_relu(z2, 0.2) + repeat(x, 4), 2))    NetWorks/PixelShuffleUpsample.py:36-45
 *                         z2 [B,4C,H,W] (layer_2 pre-activation), x [B,C,H,W], y / dy [B,C,2H,2W]; dz2 / dx may be NULL;
 *                         dx is accumulated (+=, caller zero-initialises)
 *   hn_rgb_upsample_*   : y = blur(bilinear x2, align_corners=False)                         NetWorks/neural_renderer.py:47-50
 *                         x [planes,H,W], y / dy [planes,2H,2W]; dx is accumulated (+=, caller zero-initialises)            */
int hn_upsample_tail_fwd(const float* z2, const float* x, const float* f3_host, float* y, int B, int C, int H, int W, void* stream);
int hn_upsample_tail_bwd(const float* dy, const float* z2, const float* f3_host, float* dz2, float* dx, int B, int C, int H, int W, void* stream);
int hn_rgb_upsample_fwd(const float* x, const float* f3_host, float* y, int planes, int H, int W, void* stream);
int hn_rgb_upsample_bwd(const float* dy, const float* f3_host, float* dx_zeroed, int planes, int H, int W, void* stream);

/* Merge (NetWorks/HeadNeRFNet.py:103-113): merge[b,c,r] = F[b,r,c] + bg_alpha[b,r] * bg_featmap[c,r], ray r <-> pixel (r / fs, r % fs).
 * F [B,n_rays,C] and bg_alpha [B,n_rays] are the outputs of hn_composite_fwd; merge / g_merge are channel-major [B,C,n_rays]
 * (= [B,C,fs,fs]); bg_featmap [C,n_rays].  Backward: gF written, g_bg and g_bgfeat accumulated (+=, caller zero-initialises);
 * any of the three may be NULL.                                                                                          */
int hn_merge_fwd(const float* F, const float* bg_alpha, const float* bg_featmap, float* merge, int B, int n_rays, int C, void* stream);
int hn_merge_bwd(const float* g_merge, const float* bg_alpha, const float* bg_featmap, float* gF, float* g_bg_zeroed, float* g_bgfeat_zeroed,
                 int B, int n_rays, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The whole hot path behind one call per direction (fast kernels): NetWorks/HeadNeRFNet.py:123-160 up to the composited
 * features, and its autograd.  All buffers caller-owned; "zeroed" = the caller zero-initialises (the kernels accumulate).   */
typedef struct {
    hn_camera_t cam;
    hn_fold_t fold;           /* latent codes, folded weights, bias vectors (fold.B == cam.B)                     */
    const float* w_density;   /* [384]                                                                            */
    const void* packed;       /* from hn_pack_weights                                                             */
    float* bias_eff;          /* [B, HN_BIAS_STRIDE] workspace (kept for backward)                                */
    float* feat;              /* [M,256] workspace (kept for backward)                                            */
    float* sigma;             /* [M]                                                                              */
    float* delta;             /* [M]                                                                              */
    void* act;                /* hn_act_bytes(M), or NULL when no backward follows                                */
    uint32_t* masks;          /* hn_mask_bytes(M), or NULL (together with act)                                    */
    float* F;                 /* [B*n_rays, 256] out                                                              */
    float* bg_alpha;          /* [B*n_rays] out                                                                   */
    int* status;              /* zeroed device int[64]                                                            */
} hn_render_fwd_t;

int hn_render_fwd(const hn_render_fwd_t* a, void* stream);

typedef struct {
    hn_camera_t cam;
    hn_fold_t fold;
    const float* w_density;
    const void* packed;
    const float* feat; const float* sigma; const float* delta; const void* act; const uint32_t* masks;   /* from hn_render_fwd */
    const float* gF;          /* [B*n_rays, 256] upstream gradient                                                 */
    const float* g_bg;        /* [B*n_rays]                                                                        */
    float grad_target;        /* target magnitude of the scaled gradients (<= 0: 64)                               */
    void* dfeat_image;        /* hn_dfeat_image_bytes(M) workspace                                                 */
    float* dsigma;            /* [M] workspace                                                                     */
    float* ddelta;            /* [M] workspace, camera gradients only                                              */
    void* grads;              /* hn_grads_bytes(M) workspace (parameter / code gradients)                          */
    float* scale;             /* device float workspace                                                            */
    void* scale_scratch8;     /* 8 zero bytes (see hn_loss_scale)                                                  */
    float* dbias_eff;         /* [B, HN_BIAS_STRIDE] zeroed workspace                                              */
    void* items_workspace; size_t items_workspace_bytes;   /* hn_wgrad_workspace_bytes(B)                          */
    float* g_ray_o; float* g_ray_v; float* g_ray_l;        /* zeroed workspaces, camera gradients only             */
    float* dw[12]; int ld[12]; int l5_hidden_col;          /* zeroed weight-gradient outputs (NULL entries skipped) */
    float* dwf;                                            /* zeroed [192,384] workspace (dw[9] / dw[10] / fold_grads.dbias[9]) */
    hn_fold_grads_t fold_grads;                            /* code gradients (=), folded columns / biases (+=)      */
    float* dR; float* dT; float* dKinv;                    /* zeroed [B,3,3] / [B,3] / [B,3,3], or NULL             */
    int* status;
} hn_render_bwd_t;

int hn_render_bwd(const hn_render_bwd_t* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step around the path.
 * Photometric loss of the reference, Utils/HeadNeRFLossUtils.py:125-140 (calc_data_loss) summed as in :196-236
 * (calc_total_loss): bg_loss = mean((bg_img - bg_value)^2); head_loss = mean over {mask >= 0.5} of (img - gt)^2;
 * nonhead_loss = mean over {mask < 0.5} of (img - bg_value)^2, img = nan_to_num(merge_img, nan = 0).  NCHW fp32.
 * hn_photo_loss_fwd: one deterministic reduction kernel -> out[0..3] = bg, head, nonhead, total; out[4..6] = the three
 * element counts (kept for the backward).  hn_photo_loss_bwd: one kernel, d_img / d_bg_img written (=), either may be
 * NULL; gout = 4 device floats dL/d(bg, head, nonhead, total) or NULL (= 0, 0, 0, 1).                               */
typedef struct {
    int B;                    /* items of merge_img / gt / mask                                                    */
    int B_bg;                 /* items of bg_img (1 in the reference: neural_render(bg_featmap))                   */
    int HW;                   /* pixels per image plane                                                            */
    float bg_value;           /* 1 = white, 0 = black (HeadNeRFLossUtils.py:72-75)                                 */
    const float* img;         /* [B,3,HW] merge_img                                                                */
    const float* bg_img;      /* [B_bg,3,HW]                                                                       */
    const float* gt;          /* [B,3,HW]                                                                          */
    const float* mask;        /* [B,1,HW] float; head = mask >= 0.5                                                */
    float* partials;          /* hn_photo_loss_workspace_bytes() scratch (first bytes), ...                        */
    unsigned int* ticket;     /* ... one zeroed 32-bit word inside it (left zero by the kernel)                    */
    float* out;               /* [8] device floats                                                                 */
} hn_photo_loss_t;

size_t hn_photo_loss_workspace_bytes(void);
int hn_photo_loss_fwd(const hn_photo_loss_t* a, void* stream);
int hn_photo_loss_bwd(const hn_photo_loss_t* a, const float* gout, float* d_img, float* d_bg_img, void* stream);

/* Adam over one flat fp32 buffer (talker_trainer.py:722-723,1062-1067: torch.optim.Adam(model.parameters(), lr) .step()),
 * arithmetic of torch.optim.Adam's single-tensor path (amsgrad off): one launch for every parameter of the model when
 * parameters, gradients (dist.GradBucket.flat) and both moments are flat buffers.  grad_scale multiplies the gradient
 * first (1/world_size after an all-reduce SUM: the averaging costs no extra pass).  step counts from 1.                */
typedef struct {
    float lr, beta1, beta2, eps, weight_decay, grad_scale;
    int64_t step;
} hn_adam_t;

int hn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const hn_adam_t* h, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Hierarchical resampling: NetWorks/utils.py:164-265 FineSample.forward - per ray, the pdf of
 * the interior coarse compositing weights, n_fine + 1 inverse-CDF depths (uniform = NULL: linspace(0,1), the reference's
 * test mode; else the caller's uniforms [n_rays_total, n_fine + 1], its rand() of train mode), merged and sorted with the
 * coarse depths; outputs n_coarse + n_fine samples per ray, sample-major: zvals, z_dists (x ray_l) and points
 * o + d * l * z (out_pts may be NULL).  ray_o / ray_d are channel-major [B,3,n_rays] like the reference's tensors.        */
typedef struct {
    int64_t n_rays_total;     /* B * n_rays                                                                        */
    int n_rays;               /* rays per item                                                                     */
    int n_coarse, n_fine;     /* num_sample_coarse (64), num_sample_fine (128)                                     */
    const float* weights;     /* [n_rays_total, n_coarse] compositing weights (hn_composite_fwd)                   */
    const float* zvals;       /* [n_rays_total, n_coarse] coarse depths                                            */
    const float* uniform;     /* [n_rays_total, n_fine + 1] or NULL                                                */
    const float* ray_o;       /* [B,3,n_rays]                                                                      */
    const float* ray_d;       /* [B,3,n_rays]                                                                      */
    const float* ray_l;       /* [n_rays_total]                                                                    */
    float* out_zvals;         /* [n_rays_total, n_coarse + n_fine]                                                 */
    float* out_zdists;        /* [n_rays_total, n_coarse + n_fine]                                                 */
    float* out_pts;           /* [n_rays_total, n_coarse + n_fine, 3] or NULL                                      */
} hn_fine_sample_t;

int hn_fine_sample(const hn_fine_sample_t* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * The consumer: NeuralRenderer.forward (NetWorks/neural_renderer.py:72-91) with its
 * PixelShuffleUpsample blocks (NetWorks/PixelShuffleUpsample.py:36-45), forward and backward, one call per direction.
 * Every 1x1 convolution, LeakyReLU(0.2), RGB head + skip sum and the final sigmoid run in one grouped tensor-core GEMM
 * kernel (tcgen05 kind::tf32 - the precision cuDNN's default gives the reference on this GPU - fp32 accumulation) over the
 * NCHW planes; the tails are hn_upsample_tail_* / hn_rgb_upsample_*.  Weights are the state-dict tensors as they are:
 * Conv2d weight [out, in, 1, 1] = row-major [out, in].  channels(i) = max(feat_nc >> i, min_feat), resolution(i) =
 * featmap_size << i, out_dim = 3.  Supported: 1..4 blocks, channels a multiple of 4 and <= 256 after the first block.   */
#define HN_NR_MAX_BLOCKS 4
typedef struct {
    int B, n_blocks, feat_nc, min_feat, featmap_size, final_actvn;
    const float* x;                                   /* [B, feat_nc, fs, fs]                                       */
    const float* w1[HN_NR_MAX_BLOCKS];                /* feat_upsample_list.i.layer_1.weight [2C, C]                */
    const float* b1[HN_NR_MAX_BLOCKS];
    const float* w2[HN_NR_MAX_BLOCKS];                /* feat_upsample_list.i.layer_2.weight [4C, 2C]               */
    const float* b2[HN_NR_MAX_BLOCKS];
    const float* wf[HN_NR_MAX_BLOCKS];                /* feat_layers.i.weight [C', C]                               */
    const float* bf[HN_NR_MAX_BLOCKS];
    const float* wrgb[HN_NR_MAX_BLOCKS + 1];          /* feat_2_rgb_list.j.weight [3, channels(j)]                  */
    const float* brgb[HN_NR_MAX_BLOCKS + 1];
    float tail_taps[HN_NR_MAX_BLOCKS][3];             /* feat_upsample_list.i.blur_layer.f, HOST values             */
    float rgb_taps[3];                                /* rgb_upsample.1.f, HOST values                              */
    float* saved;                                     /* hn_nr_saved_floats() floats: activations kept for backward */
    float* img;                                       /* [B, 3, fs << n_blocks, fs << n_blocks] out                 */
    int* status;                                      /* zeroed device int[64]                                      */
} hn_nr_fwd_t;

typedef struct {
    hn_nr_fwd_t f;                                    /* the forward call's arguments (saved / img as it left them) */
    const float* g_img;                               /* dL/dimg [B, 3, S, S]                                       */
    float* scratch;                                   /* hn_nr_scratch_floats() floats                              */
    float* g_x;                                       /* dL/dx [B, feat_nc, fs, fs] out (overwritten), or NULL      */
    /* weight / bias gradients, ACCUMULATED (+=, atomics; caller zero-initialises or passes .grad); NULL weight = skip  */
    float* dw1[HN_NR_MAX_BLOCKS]; float* db1[HN_NR_MAX_BLOCKS];
    float* dw2[HN_NR_MAX_BLOCKS]; float* db2[HN_NR_MAX_BLOCKS];
    float* dwf[HN_NR_MAX_BLOCKS]; float* dbf[HN_NR_MAX_BLOCKS];
    float* dwrgb[HN_NR_MAX_BLOCKS + 1]; float* dbrgb[HN_NR_MAX_BLOCKS + 1];
} hn_nr_bwd_t;

long long hn_nr_saved_floats(int B, int n_blocks, int feat_nc, int min_feat, int featmap_size);     /* -1: unsupported geometry */
long long hn_nr_scratch_floats(int B, int n_blocks, int feat_nc, int min_feat, int featmap_size);
int hn_nr_launches(int n_blocks, int backward);                                                     /* kernels per call         */
int hn_nr_fwd(const hn_nr_fwd_t* a, void* stream);
int hn_nr_bwd(const hn_nr_bwd_t* b, void* stream);

/* Bytes of the saved-for-backward buffers for M samples. */
size_t hn_act_bytes(int64_t M);
size_t hn_grads_bytes(int64_t M);
size_t hn_mask_bytes(int64_t M);
size_t hn_dfeat_image_bytes(int64_t M);

#ifdef __cplusplus
}
#endif
#endif /* HEADNERF_B200_H_ */
