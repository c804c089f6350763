"""GPU parity tests, edge cases and full-size properties (through the C ABI):
  * every supported samples-per-ray count (32 / 64 / 128), ragged ray counts (padding to whole 128-sample tiles, a single
    ray), batch sizes 1 and 3, jittered and plain sampling, both precision modes - forward vs the oracle (max-abs-err
    <= 1e-3) and gradients vs the oracle's autograd (cosine >= 0.999);
  * error behaviour of the drop-in module (unsupported sample counts, wrong code widths, bad mode);
  * at BASELINE.json's full size (Reso64, batch 2: 524 288 ray*samples), where the CPU oracle would take minutes,
    size-independent properties: run-to-run determinism of the forward, ray-sharding invariance (the multi-GPU partition),
    linearity of the backward pass in the upstream gradient, compositing bounds."""
import pytest
import torch

from oracle import headnerf_oracle as O
from _util import cosine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CODES = ("shape_code", "appea_code", "audiostyle")


def _net(hn, fs, ns, variant="init", precision="fast"):
    opt = O.OracleOptions(featmap_size=fs, pred_img_size=4 * fs, num_sample_coarse=ns)
    bopt = hn.BaseOptions({"featmap_size": fs, "featmap_nc": 256, "pred_img_size": 4 * fs})
    bopt.num_sample_coarse = ns                                   # HeadNeRFOptions.py:20 (an attribute, not a "para" key)
    net = hn.HeadNeRFNet(bopt, False, False)
    sd = O.formula_state_dict(opt, variant)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    net.precision = precision
    return opt, sd, net


@pytest.mark.parametrize("ns,n_rays,B,jitter,precision", [
    (32, 37, 3, True, "fast"), (32, 4, 1, False, "high"), (64, 1, 1, False, "fast"), (64, 131, 2, True, "high"),
    (128, 37, 3, False, "fast"), (128, 5, 2, True, "high"), (128, 1, 1, True, "fast")])
def test_ragged_rays_and_sample_counts(hn, ns, n_rays, B, jitter, precision):
    opt, sd, net = _net(hn, 8, ns, precision=precision)
    assert net.num_sample_coarse == ns
    inp = O.synthetic_inputs(opt, B, seed=11 + ns + n_rays, n_rays=n_rays, jitter=jitter)
    mode = "train" if jitter else "test"
    # oracle: forward + gradients of codes and MLP parameters for a fixed upstream gradient
    sdo = {k: v.clone().requires_grad_(k.startswith("fg_CD_predictor")) for k, v in sd.items()}
    xo = {k: v.clone().requires_grad_(k in CODES) for k, v in inp.items()}
    r = O.render_features(sdo, opt, mode, xo["batch_xy"], xo["audiostyle"], xo["shape_code"], xo["appea_code"],
                          xo["batch_Rmats"], xo["batch_Tvecs"], xo["batch_inv_inmats"], t_rand=xo.get("t_rand"))
    gen = torch.Generator().manual_seed(1)
    gF, gb = torch.randn(r["F"].shape, generator=gen), torch.randn(r["bg_alpha"].shape, generator=gen)
    torch.autograd.backward([r["F"], r["bg_alpha"]], [gF, gb])
    # CUDA path
    xc = {k: v.to(DEV).requires_grad_(k in CODES) for k, v in inp.items()}
    Fm, bg = net.render_rays(mode, xc["batch_xy"], xc["audiostyle"], xc["shape_code"], xc["appea_code"],
                             xc["batch_Rmats"], xc["batch_Tvecs"], xc["batch_inv_inmats"], t_rand=xc.get("t_rand"))
    assert Fm.shape == (B, n_rays, 256) and bg.shape == (B, n_rays)
    torch.autograd.backward([Fm, bg], [gF.permute(0, 2, 1).contiguous().to(DEV), gb[:, 0].contiguous().to(DEV)])
    hn.ops.check_status(net.last_meta["last_status"], "edge case")
    errF = (Fm.detach().cpu() - r["F"].detach().permute(0, 2, 1)).abs().max().item()
    errA = (bg.detach().cpu() - r["bg_alpha"].detach()[:, 0]).abs().max().item()
    assert errF <= 1e-3 and errA <= 1e-3, (errF, errA)
    worst = min([(cosine(xc[k].grad, xo[k].grad), k) for k in CODES] +
                [(cosine(p.grad, sdo["fg_CD_predictor." + n].grad), n) for n, p in net.fg_CD_predictor.named_parameters()])
    print(f"ns={ns} rays={n_rays} B={B} {mode} {precision}: F err {errF:.1e}, bg err {errA:.1e}, worst gradient cosine {worst[0]:.6f} ({worst[1]})")
    assert worst[0] >= 0.999, worst


def test_gaze_columns(hn):
    """include_gaze=True (HeadNeRFNet.py:49-52; talker_trainer.py:736-747): eye_gaze_dim extra columns ride at the end of the
    shape/expression code and of the folded column blocks of FeaExt_module_0 / _5.  Oracle: the same algebra with a 2-wider code."""
    opt = O.OracleOptions(featmap_size=8, pred_img_size=32, expr_code_dims=79 + 2)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 8, "featmap_nc": 256, "pred_img_size": 32}), False, False, include_gaze=True, eye_gaze_dim=2)
    sd = O.formula_state_dict(opt, "trained")
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    assert net.precision == "auto"                               # trained-like weights: the default must hold the absolute gate
    inp = O.synthetic_inputs(opt, 2, seed=21)
    assert inp["shape_code"].shape[1] == 181
    sdo = {k: v.clone().requires_grad_(k.startswith("fg_CD_predictor")) for k, v in sd.items()}
    xo = {k: v.clone().requires_grad_(k in CODES) for k, v in inp.items()}
    r = O.render_features(sdo, opt, "test", xo["batch_xy"], xo["audiostyle"], xo["shape_code"], xo["appea_code"],
                          xo["batch_Rmats"], xo["batch_Tvecs"], xo["batch_inv_inmats"])
    gen = torch.Generator().manual_seed(2)
    gF, gb = torch.randn(r["F"].shape, generator=gen), torch.randn(r["bg_alpha"].shape, generator=gen)
    torch.autograd.backward([r["F"], r["bg_alpha"]], [gF, gb])
    xc = {k: v.to(DEV).requires_grad_(k in CODES) for k, v in inp.items()}
    Fm, bg = net.render_rays("test", xc["batch_xy"], xc["audiostyle"], xc["shape_code"], xc["appea_code"],
                             xc["batch_Rmats"], xc["batch_Tvecs"], xc["batch_inv_inmats"])
    torch.autograd.backward([Fm, bg], [gF.permute(0, 2, 1).contiguous().to(DEV), gb[:, 0].contiguous().to(DEV)])
    net.check_faults()
    errF = (Fm.detach().cpu() - r["F"].detach().permute(0, 2, 1)).abs().max().item()
    scale = float(r["F"].detach().abs().max())
    assert errF <= 1e-3, errF
    worst = min([(cosine(xc[k].grad, xo[k].grad), k) for k in CODES] +
                [(cosine(p.grad, sdo["fg_CD_predictor." + n].grad), n) for n, p in net.fg_CD_predictor.named_parameters()])
    print(f"gaze: F err {errF:.1e} (|F|max {scale:.2f}), worst gradient cosine {worst[0]:.6f} ({worst[1]})")
    assert worst[0] >= 0.999, worst


def test_bias_only_fine_tuning_gets_every_bias_gradient(hn):
    """All twelve weights frozen, biases trainable: the weight pass must still visit the layers whose bias gradients no latent code
    needs (hn_mlp_bwd_weights_t.want_all_bias) - round 1 returned silent zeros for nine of the twelve bias vectors."""
    opt, sd, net = _net(hn, 16, 64, variant="init")              # random-init weights: the single-pass kernels' own territory; 256 rays
                                                                  # per item (with 64 the deepest layer's gradient sits at 0.9990)
    for n, p in net.fg_CD_predictor.named_parameters():
        p.requires_grad_(n.endswith(".bias"))
    inp = O.synthetic_inputs(opt, 2, seed=31)
    sdo = {k: v.clone().requires_grad_(k.startswith("fg_CD_predictor") and k.endswith(".bias")) for k, v in sd.items()}
    r = O.render_features(sdo, opt, "test", inp["batch_xy"], inp["audiostyle"], inp["shape_code"], inp["appea_code"],
                          inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    gen = torch.Generator().manual_seed(3)
    gF, gb = torch.randn(r["F"].shape, generator=gen), torch.randn(r["bg_alpha"].shape, generator=gen)
    torch.autograd.backward([r["F"], r["bg_alpha"]], [gF, gb])
    xc = {k: v.to(DEV) for k, v in inp.items()}
    Fm, bg = net.render_rays("test", xc["batch_xy"], xc["audiostyle"], xc["shape_code"], xc["appea_code"],
                             xc["batch_Rmats"], xc["batch_Tvecs"], xc["batch_inv_inmats"])
    torch.autograd.backward([Fm, bg], [gF.permute(0, 2, 1).contiguous().to(DEV), gb[:, 0].contiguous().to(DEV)])
    net.check_faults()
    for n, p in net.fg_CD_predictor.named_parameters():
        if n.endswith(".bias"):
            ref = sdo["fg_CD_predictor." + n].grad
            assert p.grad is not None and float(ref.abs().max()) > 0, n
            assert cosine(p.grad, ref) >= 0.999, (n, cosine(p.grad, ref))
        else:
            assert p.grad is None, n


def test_error_behaviour(hn):
    with pytest.raises(hn._lib.HeadNeRFLibraryError):           # 48 samples per ray: outside what the kernels are specialised for
        opt, sd, net = _net(hn, 8, 48)
        inp = {k: v.to(DEV) for k, v in O.synthetic_inputs(opt, 1, seed=0, n_rays=8).items()}
        net.render_rays("test", inp["batch_xy"], inp["audiostyle"], inp["shape_code"], inp["appea_code"],
                        inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    opt, sd, net = _net(hn, 8, 64)
    inp = {k: v.to(DEV) for k, v in O.synthetic_inputs(opt, 2, seed=0).items()}
    with pytest.raises(ValueError):                               # code width does not match the network
        net.render_rays("test", inp["batch_xy"], inp["audiostyle"], inp["shape_code"][:, :100], inp["appea_code"],
                        inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    with pytest.raises(AssertionError):                           # HeadNeRFNet.py:199
        net("fit", inp["batch_xy"], None, inp["audiostyle"], None, inp["shape_code"], inp["appea_code"],
            inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    with pytest.raises(AssertionError):                           # HeadNeRFNet.py:134: bg_code must be None
        net("test", inp["batch_xy"], None, inp["audiostyle"], torch.zeros(2, 8, device=DEV), inp["shape_code"], inp["appea_code"],
            inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    with pytest.raises(AssertionError):                           # forward renders whole feature maps only (HeadNeRFNet.py:103-106)
        net("test", inp["batch_xy"][:, :, :10].contiguous(), None, inp["audiostyle"], None, inp["shape_code"], inp["appea_code"],
            inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    out = net("test", inp["batch_xy"], None, inp["audiostyle"], None, inp["shape_code"], inp["appea_code"],
              inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"], dist_expr=True, some_ignored_kwarg=1)
    assert out["coarse_dict"]["merge_img"].shape == (2, 3, 32, 32) and out["coarse_dict"]["bg_img"].shape == (1, 3, 32, 32)


@pytest.fixture(scope="module")
def full(hn):
    """Reso64, batch 2 - the benchmark configuration (BASELINE.json configs[1])."""
    opt, sd, net = _net(hn, 64, 64, variant="trained")
    inp = {k: v.to(DEV) for k, v in O.synthetic_inputs(opt, 2, seed=0, jitter=True).items()}
    return opt, net, inp


def _render(net, inp, sl=slice(None), grad=False):
    x = {k: (v.clone().requires_grad_(True) if (grad and k in CODES) else v) for k, v in inp.items()}
    Fm, bg = net.render_rays("train", x["batch_xy"][:, :, sl].contiguous(), x["audiostyle"], x["shape_code"], x["appea_code"],
                             x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"], t_rand=x["t_rand"][:, sl].contiguous())
    return Fm, bg, x


def test_full_size_determinism_and_sharding_invariance(hn, full):
    opt, net, inp = full
    with torch.no_grad():
        F1, b1, _ = _render(net, inp)
        F2, b2, _ = _render(net, inp)
        assert torch.equal(F1, F2) and torch.equal(b1, b2), "forward is not run-to-run deterministic"
        # the multi-GPU partition (dist.shard_rays): rays [0, 1366) / [1366, 4096) rendered separately == rendered together, bit for bit
        Fa, ba, _ = _render(net, inp, slice(0, 1366))
        Fb, bb, _ = _render(net, inp, slice(1366, 4096))
    hn.ops.check_status(net.last_meta["last_status"], "full size forward")
    assert torch.equal(torch.cat([Fa, Fb], 1), F1) and torch.equal(torch.cat([ba, bb], 1), b1), "a ray's result depends on its shard"
    assert torch.isfinite(F1).all() and float(b1.min()) >= -1e-6 and float(b1.max()) <= 1 + 1e-6      # bg_alpha = 1 - sum of weights


@pytest.mark.parametrize("precision", ["fast", "high"])
def test_full_size_backward_is_linear_in_the_upstream_gradient(hn, full, precision):
    opt, net, inp = full
    net.precision = precision
    gen = torch.Generator().manual_seed(9)
    gF = (torch.randn(2, 4096, 256, generator=gen) * 1e-3).to(DEV)
    gb = (torch.randn(2, 4096, generator=gen) * 1e-3).to(DEV)
    grads = []
    for mult in (1.0, 4.0):                                     # a power of two: the loss scale moves with it, mantissas do not
        net.zero_grad(set_to_none=True)
        Fm, bg, x = _render(net, inp, grad=True)
        torch.autograd.backward([Fm, bg], [gF * mult, gb * mult])
        grads.append([x[k].grad.clone() for k in CODES] + [p.grad.clone() for p in net.fg_CD_predictor.parameters()])
    hn.ops.check_status(net.last_meta["last_status"], "full size backward")
    net.precision = "fast"
    for g1, g4 in zip(*grads):
        assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0
        # atomics reorder fp32 sums between runs: equal up to reduction noise
        assert (g4 - 4.0 * g1).abs().max() <= 1e-3 * (4.0 * g1).abs().max() + 1e-12
