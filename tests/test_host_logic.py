"""CPU-only tests: the C-ABI library loads and exports what include/headnerf_b200.h declares, the drop-in module
reproduces the reference state-dict layout, latent folding is exact, and the product path refuses to run on CPU."""
import ctypes
import os
import re

import pytest
import torch

from oracle import headnerf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(hn):
    header = open(os.path.join(ROOT, "include", "headnerf_b200.h")).read()
    declared = set(re.findall(r"\b(hn_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(hn._lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(hn._lib.EXPORTS) == declared
    lib.hn_abi_version.restype = ctypes.c_int
    assert lib.hn_abi_version() == 4
    lib.hn_packed_weights_bytes.restype = ctypes.c_size_t
    assert lib.hn_packed_weights_bytes() % 16384 == 0          # no GPU needed: host-side schedule only


def test_ctypes_structs_match_header_constants(hn):
    L = hn._lib
    header = open(os.path.join(ROOT, "include", "headnerf_b200.h")).read()
    consts = {k: int(v) for k, v in re.findall(r"#define (HN_[A-Z0-9_]+) (-?\d+)\b", header)}
    assert consts["HN_HIDDEN"] == L.HIDDEN and consts["HN_FEAT"] == L.FEAT and consts["HN_PE"] == L.PE
    assert consts["HN_BIAS_STRIDE"] == L.BIAS_STRIDE and consts["HN_BIAS_OFF_DENSITY"] == L.BIAS_OFF_DENSITY
    assert consts["HN_ACT_BLOCKS"] == L.ACT_BLOCKS and consts["HN_GRAD_BLOCKS"] == L.GRAD_BLOCKS and consts["HN_MASK_WORDS"] == L.MASK_WORDS
    # pointer-only tail of the camera struct starts 8-aligned after five 4-byte fields
    assert L.Camera.xy.offset == 24 and ctypes.sizeof(L.Camera) == 64


def test_ctypes_struct_layouts_match_the_header(hn, tmp_path):
    """Every argument struct of include/headnerf_b200.h, compiled as plain C, has the size - and every field the offset - of
    its ctypes twin in _lib.py (a silent mismatch would shift pointers inside a library call)."""
    import shutil
    import subprocess
    L = hn._lib
    pairs = {"hn_weights_t": L.Weights, "hn_camera_t": L.Camera, "hn_mlp_fwd_t": L.MlpFwd, "hn_composite_fwd_t": L.CompositeFwd,
             "hn_composite_bwd_t": L.CompositeBwd, "hn_mlp_bwd_data_t": L.MlpBwdData, "hn_mlp_bwd_weights_t": L.MlpBwdWeights,
             "hn_fold_t": L.Fold, "hn_fold_grads_t": L.FoldGrads, "hn_unfuse_t": L.Unfuse, "hn_mlp_fwd_precise_t": L.MlpFwdPrecise,
             "hn_mlp_bwd_data_precise_t": L.MlpBwdDataPrecise, "hn_render_fwd_t": L.RenderFwd, "hn_render_bwd_t": L.RenderBwd,
             "hn_photo_loss_t": L.PhotoLoss, "hn_adam_t": L.Adam, "hn_fine_sample_t": L.FineSample, "hn_nr_fwd_t": L.NrFwd, "hn_nr_bwd_t": L.NrBwd}
    header = open(os.path.join(ROOT, "include", "headnerf_b200.h")).read()
    declared = set(re.findall(r"^} (hn_[a-z0-9_]+_t);", header, flags=re.M))
    assert declared == set(pairs), declared ^ set(pairs)
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "headnerf_b200.h"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True, capture_output=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), (cname, out[cname], ctypes.sizeof(cls))
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_renderer_workspace_sizes(hn):
    """hn_nr_saved_floats / hn_nr_scratch_floats / hn_nr_launches: host-side geometry only (no GPU): the saved buffer holds exactly the
    four activations of every block plus the RGB chain, unsupported geometries are refused."""
    lib = hn._lib.load()
    B, nb, nc, mf, fs = 2, 4, 256, 32, 32                     # Reso32HR: 32 x 32 x 256 -> 512 x 512
    C = [max(nc >> i, mf) for i in range(nb + 1)]
    P = [(fs << i) ** 2 for i in range(nb + 1)]
    pad = lambda n: (n + 63) // 64 * 64
    want = sum(pad(B * 2 * C[i] * P[i]) + pad(B * 4 * C[i] * P[i]) + pad(B * C[i] * P[i + 1]) + pad(B * C[i + 1] * P[i + 1]) for i in range(nb))
    want += pad(B * 3 * P[0]) + sum(2 * pad(B * 3 * P[l]) for l in range(1, nb + 1))
    assert lib.hn_nr_saved_floats(B, nb, nc, mf, fs) == want
    assert lib.hn_nr_scratch_floats(B, nb, nc, mf, fs) > 0
    assert lib.hn_nr_saved_floats(2 * B, nb, nc, mf, fs) > want
    assert lib.hn_nr_launches(nb, 0) == 5 * nb and lib.hn_nr_launches(nb, 1) == 6 * nb + 1
    for bad in [(0, nb, nc, mf, fs), (B, 0, nc, mf, fs), (B, 5, nc, mf, fs), (B, nb, 258, mf, fs), (B, nb, 1024, 512, fs), (B, nb, nc, mf, 1)]:
        assert lib.hn_nr_saved_floats(*bad) == -1 and lib.hn_nr_scratch_floats(*bad) == -1, bad


@pytest.mark.parametrize("fs,S", [(32, 256), (32, 512), (64, 512)])
def test_state_dict_layout_matches_reference(hn, fs, S):
    opt = O.OracleOptions(featmap_size=fs, pred_img_size=S)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": fs, "featmap_nc": 256, "pred_img_size": S}), False, False)
    sd = net.state_dict()
    shapes = O.state_dict_shapes(opt)               # pinned against the real reference in test_oracle_vs_reference.py
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == shapes[k], k
    n_params = sum(p.numel() for p in net.parameters())
    assert n_params == {(32, 256): 2711309, (32, 512): 2722896, (64, 512): 3497741}[(fs, S)]
    net.load_state_dict(O.formula_state_dict(opt, "init"), strict=True)


def test_unsupported_modes_raise(hn):
    opt = hn.BaseOptions()
    with pytest.raises(NotImplementedError):
        hn.HeadNeRFNet(opt, include_vd=True, hier_sampling=False)
    with pytest.raises(NotImplementedError):
        hn.HeadNeRFNet(opt, include_vd=False, hier_sampling=True)
    opt.mlp_hidden_nchannels = 256
    with pytest.raises(NotImplementedError):
        hn.HeadNeRFNet(opt, False, False)


def test_latent_folding_is_exact(hn):
    """bias_eff reproduces W[:, latent cols] @ code + b of the reference concat orders (models.py:69,75,80)."""
    torch.manual_seed(0)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 8, "featmap_nc": 256, "pred_img_size": 32}), False, False).double()
    B = 3
    shape, appea, audio = torch.randn(B, 179).double(), torch.randn(B, 127).double(), torch.randn(B, 64).double()
    bias = net._fold_biases(shape, appea, audio)
    assert bias.shape == (B, hn._lib.BIAS_STRIDE)
    fg = net.fg_CD_predictor
    w0, b0 = fg.FeaExt_module_0.weight.flatten(1), fg.FeaExt_module_0.bias
    x0 = torch.cat([torch.zeros(B, 63).double(), shape, audio], 1)           # PE part zero: only the folded columns remain
    assert torch.allclose(bias[:, :384], x0 @ w0.t() + b0, atol=1e-12)
    w5, b5 = fg.FeaExt_module_5.weight.flatten(1), fg.FeaExt_module_5.bias
    x5 = torch.cat([torch.zeros(B, 63).double(), shape, torch.zeros(B, 384).double()], 1)
    assert torch.allclose(bias[:, 5 * 384:6 * 384], x5 @ w5.t() + b5, atol=1e-12)
    wr, br = fg.RGB_layer_1.weight.flatten(1), fg.RGB_layer_1.bias
    xr = torch.cat([torch.zeros(B, 384).double(), appea], 1)
    off = hn._lib.BIAS_OFF_R1
    assert torch.allclose(bias[:, off:off + 192], xr @ wr.t() + br, atol=1e-12)
    assert torch.allclose(bias[:, hn._lib.BIAS_OFF_DENSITY], fg.density_module.bias.expand(B), atol=0)


def test_gaze_columns_fold_like_shape_columns(hn):
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 8, "featmap_nc": 256, "pred_img_size": 32}), False, False,
                         include_gaze=True, eye_gaze_dim=2)
    assert net.fg_CD_predictor.FeaExt_module_0.weight.shape[1] == 63 + 181 + 64
    assert net.fg_CD_predictor.FeaExt_module_5.weight.shape[1] == 63 + 181 + 384
    bias = net._fold_biases(torch.randn(2, 181), torch.randn(2, 127), torch.randn(2, 64))
    assert bias.shape[1] == hn._lib.BIAS_STRIDE


def test_no_cpu_fallback(hn):
    """The product path must fail loudly without CUDA tensors — there is no CPU implementation to fall back to."""
    opt = O.OracleOptions(featmap_size=8, pred_img_size=32)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 8, "featmap_nc": 256, "pred_img_size": 32}), False, False)
    x = O.synthetic_inputs(opt, 1, seed=0)
    with pytest.raises((hn._lib.HeadNeRFLibraryError, RuntimeError)):
        net("test", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"],
            x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
    with pytest.raises(AssertionError):
        net("eval", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"],
            x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
    with pytest.raises(hn._lib.HeadNeRFLibraryError):
        hn.ops.composite(torch.zeros(128, 256), torch.zeros(128), torch.zeros(128), None, 64)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nerf-3dtalker-code_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("headnerf_oracle", "oracle") or "import" not in text or \
                    not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f"{f} imports the oracle"
