"""GPU parity at BASELINE.json's REAL sizes (not the fs=8/16 goldens): the CUDA path through the C ABI against the oracle
(oracle/headnerf_oracle.py, pinned bit-equal to the imported reference by tests/test_oracle_vs_reference.py) evaluated on the
same inputs in fp32 with TF32 disabled - on the CPU for config 1, on the B200 for the larger ones ("the
GPU-side oracle = the same code on the B200 box in fp32, TF32 disabled").

  config 1  HeadNeRF Reso32 forward of one latent code (32x32 rays x 64 samples)            F / bg_alpha <= 1e-3, image >= 45 dB
  config 2  HeadNeRF Reso64 forward+backward, batch 2 (the bench.py workload, mode train)   + gradient cosine >= 0.999 on every leaf
  config 3  HeadNeRF Reso32HR training step with the NeuralRenderer consumer and MSE loss    image >= 45 dB, every parameter >= 0.999
Gates are the north star's, absolute, with no scaling by the feature magnitude."""
import contextlib

import pytest
import torch

from oracle import headnerf_oracle as O
from _util import cosine, psnr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CODES = ("shape_code", "appea_code", "audiostyle")


@contextlib.contextmanager
def exact_fp32():
    """The oracle's convolutions / matmuls in true fp32 on the GPU (the reference leaves TF32 on; the oracle must not)."""
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b


def _net(hn, fs, S, variant, precision=None):
    opt = O.OracleOptions(featmap_size=fs, pred_img_size=S)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": fs, "featmap_nc": 256, "pred_img_size": S}), False, False)
    sd = O.formula_state_dict(opt, variant)
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    if precision is not None:
        net.precision = precision
    return opt, sd, net


@pytest.mark.parametrize("variant", ["init", "trained"])
def test_config1_reso32_forward(hn, variant):
    """configs[0]: model_Reso32 forward render of one synthetic latent code - oracle on the host CPU, as the config words it."""
    opt, sd, net = _net(hn, 32, 256, variant)
    inp = O.synthetic_inputs(opt, 1, seed=3)
    with torch.no_grad():
        ref_img, r = O.headnerf_forward(sd, opt, "test", inp["batch_xy"], inp["audiostyle"], inp["shape_code"], inp["appea_code"],
                                        inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
        x = {k: v.to(DEV) for k, v in inp.items()}
        Fm, bg = net.render_rays("test", x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"],
                                 x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        out = net("test", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"],
                  x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
    net.check_faults()
    errF = (Fm.cpu() - r["F"].permute(0, 2, 1)).abs().max().item()
    errA = (bg.cpu() - r["bg_alpha"][:, 0]).abs().max().item()
    p = psnr(out["coarse_dict"]["merge_img"], ref_img["coarse_dict"]["merge_img"])
    pb = psnr(out["coarse_dict"]["bg_img"], ref_img["coarse_dict"]["bg_img"])
    print(f"config 1 ({variant}, auto -> {net.last_meta['precision']}): F err {errF:.2e} (|F|max {r['F'].abs().max():.2f}), bg_alpha err {errA:.2e}, "
          f"merge_img {p:.1f} dB, bg_img {pb:.1f} dB")
    assert errF <= 1e-3 and errA <= 1e-3
    assert p >= 45.0 and pb >= 45.0


def _oracle_step_gpu(sd, opt, inp, gF, gb, mode):
    """Oracle forward + backward on the GPU in exact fp32.  gF [B,256,N_r], gb [B,1,N_r] upstream gradients."""
    with exact_fp32():
        sdo = {k: v.to(DEV).requires_grad_(k.startswith("fg_CD_predictor")) for k, v in sd.items()}
        xo = {k: v.to(DEV).requires_grad_(k in CODES) for k, v in inp.items()}
        r = O.render_features(sdo, opt, mode, xo["batch_xy"], xo["audiostyle"], xo["shape_code"], xo["appea_code"],
                              xo["batch_Rmats"], xo["batch_Tvecs"], xo["batch_inv_inmats"], t_rand=xo.get("t_rand"))
        Fo, bo = r["F"].detach().clone(), r["bg_alpha"].detach().clone()
        torch.autograd.backward([r["F"], r["bg_alpha"]], [gF.to(DEV), gb.to(DEV)])
        del r
    grads = {k: xo[k].grad for k in CODES}
    grads.update({k: v.grad for k, v in sdo.items() if v.grad is not None})
    return Fo, bo, grads


@pytest.mark.parametrize("variant,precision", [("init", "fast"), ("init", "auto"), ("trained", "auto")])
def test_config2_reso64_batch2_forward_backward(hn, variant, precision):
    """configs[1], the workload bench.py times: Reso64, batch 2, mode train (explicit jitter), upstream gradients on F and
    bg_alpha; ("init", "fast") is exactly the benchmarked kernel family on the benchmarked weights."""
    opt, sd, net = _net(hn, 64, 512, variant, precision)
    B, n_r = 2, 64 * 64
    inp = O.synthetic_inputs(opt, B, seed=0, jitter=True)
    gen = torch.Generator().manual_seed(1000)
    gF = torch.randn(B, 256, n_r, generator=gen) * 1e-3
    gb = torch.randn(B, 1, n_r, generator=gen) * 1e-3
    Fo, bo, go = _oracle_step_gpu(sd, opt, inp, gF, gb, "train")
    torch.cuda.empty_cache()
    xc = {k: v.to(DEV).requires_grad_(k in CODES) for k, v in inp.items()}
    Fm, bg = net.render_rays("train", xc["batch_xy"], xc["audiostyle"], xc["shape_code"], xc["appea_code"],
                             xc["batch_Rmats"], xc["batch_Tvecs"], xc["batch_inv_inmats"], t_rand=xc["t_rand"])
    torch.autograd.backward([Fm, bg], [gF.permute(0, 2, 1).contiguous().to(DEV), gb[:, 0].contiguous().to(DEV)])
    net.check_faults()
    errF = (Fm.detach() - Fo.permute(0, 2, 1)).abs().max().item()
    errA = (bg.detach() - bo[:, 0]).abs().max().item()
    cos = {k: cosine(xc[k].grad, go[k]) for k in CODES}
    cos.update({n: cosine(p.grad, go["fg_CD_predictor." + n]) for n, p in net.fg_CD_predictor.named_parameters()})
    worst = min(cos.items(), key=lambda t: t[1])
    print(f"config 2 ({variant}, {precision} -> {net.last_meta['precision']}): F err {errF:.2e} (|F|max {Fo.abs().max():.2f}), bg_alpha err {errA:.2e}, "
          f"worst gradient cosine {worst[1]:.6f} ({worst[0]}) over {len(cos)} leaves")
    assert errF <= 1e-3 and errA <= 1e-3
    assert len(cos) == 27 and worst[1] >= 0.999, worst


def _masked_mse_loss(img, bg_img, target, mask):
    """The photometric terms of the reference's training loss (Utils/HeadNeRFLossUtils.py:125-140) restated for the test."""
    head = ((img - target) ** 2 * mask).sum() / (mask.sum() * img.shape[1] + 1e-6)
    nonhead = ((img - target) ** 2 * (1 - mask)).sum() / ((1 - mask).sum() * img.shape[1] + 1e-6)
    return head + nonhead + ((bg_img - 1.0) ** 2).mean() * 0.0 + 0.0 * bg_img.sum()


def test_config3_reso32hr_training_step(hn):
    """configs[2]: Reso32HR (32x32 rays, 512x512 image, 4 up-sampling blocks) - full HeadNeRFNet.forward in mode train with the
    NeuralRenderer consumer, a masked photometric loss, gradients of every parameter, against the oracle end to end."""
    opt, sd, net = _net(hn, 32, 512, "trained")
    net.train()
    B = 2
    inp = O.synthetic_inputs(opt, B, seed=5, jitter=True)
    gen = torch.Generator().manual_seed(77)
    target = torch.rand(B, 3, 512, 512, generator=gen)
    mask = (torch.rand(B, 1, 512, 512, generator=gen) > 0.4).float()
    with exact_fp32():
        sdo = {k: v.to(DEV).requires_grad_(not k.endswith(".f")) for k, v in sd.items()}
        xo = {k: v.to(DEV).requires_grad_(k in CODES) for k, v in inp.items()}
        res, r = O.headnerf_forward(sdo, opt, "train", xo["batch_xy"], xo["audiostyle"], xo["shape_code"], xo["appea_code"],
                                    xo["batch_Rmats"], xo["batch_Tvecs"], xo["batch_inv_inmats"], t_rand=xo["t_rand"])
        img_o = res["coarse_dict"]["merge_img"]
        _masked_mse_loss(img_o, res["coarse_dict"]["bg_img"], target.to(DEV), mask.to(DEV)).backward()
    xc = {k: v.to(DEV).requires_grad_(k in CODES) for k, v in inp.items()}
    # HeadNeRFNet.forward draws its own jitter (like the reference); render with the explicit one and run the consumer as forward() does
    Fm, bg = net.render_rays("train", xc["batch_xy"], xc["audiostyle"], xc["shape_code"], xc["appea_code"],
                             xc["batch_Rmats"], xc["batch_Tvecs"], xc["batch_inv_inmats"], t_rand=xc["t_rand"])
    bgf = net.neural_render.get_bg_featmap()
    merge = hn.ops.MergeFunction.apply(Fm, bg, bgf)
    imgs = net.neural_render(torch.cat([merge, bgf], 0))
    img, bg_img = imgs[:B], imgs[B:]
    _masked_mse_loss(img, bg_img, target.to(DEV), mask.to(DEV)).backward()
    net.check_faults()
    p = psnr(img.detach(), img_o.detach())
    errF = (Fm.detach() - r["F"].detach().permute(0, 2, 1)).abs().max().item()
    cos = {k: cosine(xc[k].grad, xo[k].grad) for k in CODES}
    for n, prm in net.named_parameters():
        if sdo[n].grad is not None and float(sdo[n].grad.abs().max()) > 0:
            assert prm.grad is not None, n
            cos[n] = cosine(prm.grad, sdo[n].grad)
    worst = min(cos.items(), key=lambda t: t[1])
    print(f"config 3 (auto -> {net.last_meta['precision']}): merge_img {p:.1f} dB, F err {errF:.2e}, worst gradient cosine {worst[1]:.6f} ({worst[0]}) "
          f"over {len(cos)} leaves")
    assert p >= 45.0 and errF <= 1e-3
    assert worst[1] >= 0.999, worst
