"""ctypes binding of libheadnerf_b200.so — the C ABI declared in include/headnerf_b200.h.

There is deliberately no fallback: if the shared library is missing or a call fails, this raises.
The structures below mirror the header field for field."""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# HN_LIB_PATH: load an experimental build of the SAME library (profiling variants under build/); never a fallback
LIB_PATH = os.environ.get("HN_LIB_PATH") or os.path.join(_PKG, "libheadnerf_b200.so")

HIDDEN, FEAT, RGB1, PE, TILE = 384, 256, 192, 63, 128
BIAS_OFF_R0, BIAS_OFF_R1, BIAS_OFF_R2, BIAS_OFF_DENSITY, BIAS_STRIDE = 3072, 3456, 3648, 3904, 3920
ACT_BLOCKS, GRAD_BLOCKS, MASK_WORDS = 58, 59, 104

_p = C.c_void_p


class Camera(C.Structure):
    _fields_ = [("B", C.c_int), ("n_rays", C.c_int), ("n_samples", C.c_int),
                ("world_z1", C.c_float), ("world_z2", C.c_float),
                ("xy", _p), ("Rmats", _p), ("Tvecs", _p), ("inv_inmats", _p), ("t_rand", _p)]


class Weights(C.Structure):
    _fields_ = [("w", _p * 12), ("ld", C.c_int * 12), ("l5_hidden_col", C.c_int)]


class MlpFwd(C.Structure):
    _fields_ = [("cam", Camera), ("bias", _p), ("w_density", _p), ("packed", _p), ("feat", _p), ("sigma", _p),
                ("delta", _p), ("zvals", _p), ("act", _p), ("masks", _p), ("status", _p)]


class CompositeFwd(C.Structure):
    _fields_ = [("n_rays_total", C.c_int), ("n_samples", C.c_int), ("C", C.c_int),
                ("feat", _p), ("sigma", _p), ("delta", _p), ("zvals", _p),
                ("F", _p), ("bg_alpha", _p), ("depth", _p), ("weights", _p)]


class CompositeBwd(C.Structure):
    _fields_ = [("n_rays_total", C.c_int), ("n_samples", C.c_int), ("C", C.c_int),
                ("feat", _p), ("sigma", _p), ("delta", _p), ("zvals", _p),
                ("gF", _p), ("g_bg", _p), ("g_depth", _p),
                ("dfeat", _p), ("dfeat_image", _p), ("grad_scale", _p), ("dsigma", _p), ("ddelta", _p)]


class MlpBwdData(C.Structure):
    _fields_ = [("cam", Camera), ("packed", _p), ("w_density", _p), ("dfeat_image", _p), ("dsigma", _p),
                ("ddelta", _p), ("sigma", _p), ("grad_scale", _p), ("masks", _p), ("act", _p), ("grads", _p),
                ("g_ray_o", _p), ("g_ray_v", _p), ("g_ray_l", _p), ("status", _p)]


class MlpBwdWeights(C.Structure):
    _fields_ = [("B", C.c_int), ("n_rays", C.c_int), ("n_samples", C.c_int),
                ("act", _p), ("grads", _p), ("dfeat_image", _p), ("grad_scale", _p),
                ("dw", _p * 12), ("ld", C.c_int * 12), ("l5_hidden_col", C.c_int), ("dbias", _p),
                ("items_workspace", _p), ("items_workspace_bytes", C.c_size_t), ("status", _p), ("want_all_bias", C.c_int),
                ("r0_fused", C.c_int), ("dwf", _p), ("det_workspace", _p), ("det_workspace_bytes", C.c_size_t)]


class MlpFwdPrecise(C.Structure):
    _fields_ = [("cam", Camera), ("bias", _p), ("w_density", _p), ("packed_hl", _p), ("feat", _p), ("sigma", _p),
                ("delta", _p), ("zvals", _p), ("acts", _p), ("status", _p)]


class MlpBwdDataPrecise(C.Structure):
    _fields_ = [("cam", Camera), ("packed_hl", _p), ("w_density", _p), ("dfeat", _p), ("dsigma", _p), ("ddelta", _p),
                ("sigma", _p), ("grad_scale", _p), ("acts", _p), ("gz", _p), ("g_ray_o", _p), ("g_ray_v", _p),
                ("g_ray_l", _p), ("dbias", _p), ("act_image", _p), ("grads_image", _p), ("dfeat_image", _p), ("status", _p)]


class Fold(C.Structure):
    _fields_ = [("B", C.c_int), ("shape_dims", C.c_int), ("appea_dims", C.c_int), ("w0", _p), ("ld0", C.c_int), ("w5", _p), ("ld5", C.c_int),
                ("wr1", _p), ("ldr1", C.c_int), ("bias", _p * 12), ("shape_code", _p), ("audio", _p), ("appea", _p),
                ("r0_fused", C.c_int), ("wr0", _p), ("ldr0", C.c_int)]


class Unfuse(C.Structure):
    _fields_ = [("B", C.c_int), ("wr0", _p), ("ldr0", C.c_int), ("wr1", _p), ("ldr1", C.c_int), ("b_r0", _p), ("dwf", _p),
                ("dbias_eff", _p), ("dwr0", _p), ("dwr1", _p)]


class FoldGrads(C.Structure):
    _fields_ = [("dshape", _p), ("daudio", _p), ("dappea", _p), ("dw0", _p), ("dw5", _p), ("dwr1", _p), ("dbias", _p * 12)]


class RenderFwd(C.Structure):
    _fields_ = [("cam", Camera), ("fold", Fold), ("w_density", _p), ("packed", _p), ("bias_eff", _p), ("feat", _p), ("sigma", _p),
                ("delta", _p), ("act", _p), ("masks", _p), ("F", _p), ("bg_alpha", _p), ("status", _p)]


class RenderBwd(C.Structure):
    _fields_ = [("cam", Camera), ("fold", Fold), ("w_density", _p), ("packed", _p), ("feat", _p), ("sigma", _p), ("delta", _p),
                ("act", _p), ("masks", _p), ("gF", _p), ("g_bg", _p), ("grad_target", C.c_float), ("dfeat_image", _p), ("dsigma", _p),
                ("ddelta", _p), ("grads", _p), ("scale", _p), ("scale_scratch8", _p), ("dbias_eff", _p), ("items_workspace", _p),
                ("items_workspace_bytes", C.c_size_t), ("g_ray_o", _p), ("g_ray_v", _p), ("g_ray_l", _p), ("dw", _p * 12),
                ("ld", C.c_int * 12), ("l5_hidden_col", C.c_int), ("dwf", _p), ("fold_grads", FoldGrads), ("dR", _p), ("dT", _p), ("dKinv", _p),
                ("status", _p)]


class PhotoLoss(C.Structure):
    _fields_ = [("B", C.c_int), ("B_bg", C.c_int), ("HW", C.c_int), ("bg_value", C.c_float), ("img", _p), ("bg_img", _p), ("gt", _p),
                ("mask", _p), ("partials", _p), ("ticket", _p), ("out", _p)]


class FineSample(C.Structure):
    _fields_ = [("n_rays_total", C.c_int64), ("n_rays", C.c_int), ("n_coarse", C.c_int), ("n_fine", C.c_int), ("weights", _p), ("zvals", _p),
                ("uniform", _p), ("ray_o", _p), ("ray_d", _p), ("ray_l", _p), ("out_zvals", _p), ("out_zdists", _p), ("out_pts", _p)]


class Adam(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("grad_scale", C.c_float), ("step", C.c_int64)]


NR_MAX_BLOCKS = 4


class NrFwd(C.Structure):
    _fields_ = [("B", C.c_int), ("n_blocks", C.c_int), ("feat_nc", C.c_int), ("min_feat", C.c_int), ("featmap_size", C.c_int),
                ("final_actvn", C.c_int), ("x", _p),
                ("w1", _p * NR_MAX_BLOCKS), ("b1", _p * NR_MAX_BLOCKS), ("w2", _p * NR_MAX_BLOCKS), ("b2", _p * NR_MAX_BLOCKS),
                ("wf", _p * NR_MAX_BLOCKS), ("bf", _p * NR_MAX_BLOCKS), ("wrgb", _p * (NR_MAX_BLOCKS + 1)), ("brgb", _p * (NR_MAX_BLOCKS + 1)),
                ("tail_taps", (C.c_float * 3) * NR_MAX_BLOCKS), ("rgb_taps", C.c_float * 3),
                ("saved", _p), ("img", _p), ("status", _p)]


class NrBwd(C.Structure):
    _fields_ = [("f", NrFwd), ("g_img", _p), ("scratch", _p), ("g_x", _p),
                ("dw1", _p * NR_MAX_BLOCKS), ("db1", _p * NR_MAX_BLOCKS), ("dw2", _p * NR_MAX_BLOCKS), ("db2", _p * NR_MAX_BLOCKS),
                ("dwf", _p * NR_MAX_BLOCKS), ("dbf", _p * NR_MAX_BLOCKS), ("dwrgb", _p * (NR_MAX_BLOCKS + 1)), ("dbrgb", _p * (NR_MAX_BLOCKS + 1))]


EXPORTS = ["hn_abi_version", "hn_last_error", "hn_packed_weights_bytes", "hn_pack_weights", "hn_sample_rays",
           "hn_mlp_fwd", "hn_composite_fwd", "hn_composite_bwd", "hn_mlp_bwd_data", "hn_mlp_bwd_weights",
           "hn_act_bytes", "hn_grads_bytes", "hn_mask_bytes", "hn_dfeat_image_bytes", "hn_wgrad_workspace_bytes", "hn_wgrad_det_workspace_bytes",
           "hn_precise_packed_bytes", "hn_precise_workspace_floats", "hn_pack_weights_precise", "hn_mlp_fwd_precise",
           "hn_mlp_bwd_data_precise", "hn_fold_bias", "hn_fold_bias_bwd", "hn_loss_scale", "hn_camera_bwd",
           "hn_upsample_tail_fwd", "hn_upsample_tail_bwd", "hn_rgb_upsample_fwd", "hn_rgb_upsample_bwd", "hn_merge_fwd", "hn_merge_bwd", "hn_render_fwd", "hn_render_bwd",
           "hn_photo_loss_workspace_bytes", "hn_photo_loss_fwd", "hn_photo_loss_bwd", "hn_adam_step", "hn_fine_sample", "hn_unfuse_r0r1",
           "hn_nr_saved_floats", "hn_nr_scratch_floats", "hn_nr_launches", "hn_nr_fwd", "hn_nr_bwd"]

_lib = None


class HeadNeRFLibraryError(RuntimeError):
    pass


def load():
    """dlopen the library once.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise HeadNeRFLibraryError(
            f"{LIB_PATH} not found: the CUDA library must be built first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.hn_abi_version.restype = C.c_int
    lib.hn_last_error.restype = C.c_char_p
    lib.hn_packed_weights_bytes.restype = C.c_size_t
    for name in ("hn_act_bytes", "hn_grads_bytes", "hn_mask_bytes", "hn_dfeat_image_bytes"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.c_int64]
    lib.hn_wgrad_workspace_bytes.restype = C.c_size_t
    lib.hn_wgrad_workspace_bytes.argtypes = [C.c_int]
    lib.hn_wgrad_det_workspace_bytes.restype = C.c_size_t
    lib.hn_wgrad_det_workspace_bytes.argtypes = [C.c_int]
    lib.hn_pack_weights.argtypes = [C.POINTER(Weights), _p, _p]
    lib.hn_precise_packed_bytes.restype = C.c_size_t
    lib.hn_precise_workspace_floats.restype = C.c_size_t
    lib.hn_precise_workspace_floats.argtypes = [C.c_int64]
    lib.hn_pack_weights_precise.argtypes = [C.POINTER(Weights), _p, _p]
    lib.hn_mlp_fwd_precise.argtypes = [C.POINTER(MlpFwdPrecise), _p]
    lib.hn_mlp_bwd_data_precise.argtypes = [C.POINTER(MlpBwdDataPrecise), _p]
    lib.hn_fold_bias.argtypes = [C.POINTER(Fold), _p, _p]
    lib.hn_fold_bias_bwd.argtypes = [C.POINTER(Fold), _p, C.POINTER(FoldGrads), _p]
    lib.hn_loss_scale.argtypes = [_p, C.c_int64, C.c_float, _p, _p, _p]
    lib.hn_camera_bwd.argtypes = [C.POINTER(Camera), _p, _p, _p, _p, _p, _p, _p]
    f3 = C.POINTER(C.c_float)
    lib.hn_upsample_tail_fwd.argtypes = [_p, _p, f3, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p]
    lib.hn_upsample_tail_bwd.argtypes = [_p, _p, f3, _p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p]
    lib.hn_rgb_upsample_fwd.argtypes = [_p, f3, _p, C.c_int, C.c_int, C.c_int, _p]
    lib.hn_rgb_upsample_bwd.argtypes = [_p, f3, _p, C.c_int, C.c_int, C.c_int, _p]
    lib.hn_merge_fwd.argtypes = [_p, _p, _p, _p, C.c_int, C.c_int, C.c_int, _p]
    lib.hn_render_fwd.argtypes = [C.POINTER(RenderFwd), _p]
    lib.hn_render_bwd.argtypes = [C.POINTER(RenderBwd), _p]
    lib.hn_merge_bwd.argtypes = [_p, _p, _p, _p, _p, _p, C.c_int, C.c_int, C.c_int, _p]
    lib.hn_photo_loss_workspace_bytes.restype = C.c_size_t
    lib.hn_photo_loss_fwd.argtypes = [C.POINTER(PhotoLoss), _p]
    lib.hn_photo_loss_bwd.argtypes = [C.POINTER(PhotoLoss), _p, _p, _p, _p]
    lib.hn_adam_step.argtypes = [_p, _p, _p, _p, C.c_int64, C.POINTER(Adam), _p]
    lib.hn_fine_sample.argtypes = [C.POINTER(FineSample), _p]
    lib.hn_unfuse_r0r1.argtypes = [C.POINTER(Unfuse), _p]
    lib.hn_unfuse_r0r1.restype = C.c_int
    for name in ("hn_nr_saved_floats", "hn_nr_scratch_floats"):
        getattr(lib, name).restype = C.c_longlong
        getattr(lib, name).argtypes = [C.c_int] * 5
    lib.hn_nr_launches.argtypes = [C.c_int, C.c_int]
    lib.hn_nr_launches.restype = C.c_int
    lib.hn_nr_fwd.argtypes = [C.POINTER(NrFwd), _p]
    lib.hn_nr_bwd.argtypes = [C.POINTER(NrBwd), _p]
    lib.hn_nr_fwd.restype = lib.hn_nr_bwd.restype = C.c_int
    for name in ("hn_photo_loss_fwd", "hn_photo_loss_bwd", "hn_adam_step", "hn_fine_sample"):
        getattr(lib, name).restype = C.c_int
    lib.hn_sample_rays.argtypes = [C.POINTER(Camera), _p, _p, _p, _p, _p, _p]
    lib.hn_mlp_fwd.argtypes = [C.POINTER(MlpFwd), _p]
    lib.hn_composite_fwd.argtypes = [C.POINTER(CompositeFwd), _p]
    lib.hn_composite_bwd.argtypes = [C.POINTER(CompositeBwd), _p]
    lib.hn_mlp_bwd_data.argtypes = [C.POINTER(MlpBwdData), _p]
    lib.hn_mlp_bwd_weights.argtypes = [C.POINTER(MlpBwdWeights), _p]
    for name in ("hn_pack_weights", "hn_sample_rays", "hn_mlp_fwd", "hn_composite_fwd", "hn_composite_bwd",
                 "hn_mlp_bwd_data", "hn_mlp_bwd_weights", "hn_pack_weights_precise", "hn_mlp_fwd_precise",
                 "hn_mlp_bwd_data_precise", "hn_fold_bias", "hn_fold_bias_bwd", "hn_loss_scale", "hn_camera_bwd",
                 "hn_upsample_tail_fwd", "hn_upsample_tail_bwd", "hn_rgb_upsample_fwd", "hn_rgb_upsample_bwd", "hn_merge_fwd", "hn_merge_bwd", "hn_render_fwd", "hn_render_bwd"):
        getattr(lib, name).restype = C.c_int
    if lib.hn_abi_version() != 4:
        raise HeadNeRFLibraryError("ABI version mismatch between _lib.py and libheadnerf_b200.so")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().hn_last_error().decode("utf-8", "replace")
        raise HeadNeRFLibraryError(f"{what} failed (code {rc}): {msg}")
