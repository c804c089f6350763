"""GPU tests of the one-call NeuralRenderer (csrc/hn_nr.cu: hn_nr_fwd / hn_nr_bwd - grouped tcgen05
tf32 GEMMs over the NCHW planes + the fused tails) against
  * the oracle's CPU restatement of the reference modules (oracle.neural_render: NetWorks/neural_renderer.py:72-91,
    PixelShuffleUpsample.py:36-45) for the image, and
  * the plain PyTorch modules in full fp32 (TF32 off) for every gradient.
Tolerances are those of tf32 operands (10-bit mantissa, what cuDNN's default gives the reference on this GPU): image within 2e-3
absolute / PSNR >= 45 dB, gradient cosine >= 0.999 per tensor."""
import importlib
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _nr(hn):
    return importlib.import_module(hn.__name__ + ".neural_renderer")


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def _make(nr, feat_nc, fs, S, seed, final_actvn=True):
    torch.manual_seed(seed)
    net = nr.NeuralRenderer(feat_nc=feat_nc, featmap_size=fs, img_size=S, final_actvn=final_actvn).to(DEV)
    with torch.no_grad():                       # biases away from zero so that every bias path is exercised
        for k, p in net.named_parameters():
            if k.endswith(".bias"):
                p.normal_(0.0, 0.1)
    return net


@pytest.mark.parametrize("B,feat_nc,fs,S", [(1, 64, 8, 32), (2, 256, 32, 256), (3, 128, 16, 64), (1, 256, 16, 256)])
def test_image_matches_oracle(hn, oracle, B, feat_nc, fs, S):
    nr = _nr(hn)
    net = _make(nr, feat_nc, fs, S, 11 + fs)
    x = torch.randn(B, feat_nc, fs, fs, device=DEV)
    with torch.no_grad():
        img = net(x)
    assert net._fused_net(x)
    sd = {"neural_render." + k: v.detach().cpu() for k, v in net.state_dict().items()}
    opt = oracle.OracleOptions(featmap_size=fs, featmap_nc=feat_nc, pred_img_size=S)
    ref = oracle.neural_render(sd, x.cpu(), opt)
    err = (img.cpu() - ref).abs().max().item()
    psnr = 10 * math.log10(1.0 / ((img.cpu() - ref) ** 2).mean().item())
    assert err <= 2e-3, err
    assert psnr >= 45.0, psnr
    hn.ops.FAULTS.flush()


@pytest.mark.parametrize("B,feat_nc,fs,S,act", [(2, 64, 8, 32, True), (2, 256, 32, 256, True), (1, 128, 16, 128, False)])
def test_gradients_match_fp32_modules(hn, B, feat_nc, fs, S, act):
    nr = _nr(hn)
    net = _make(nr, feat_nc, fs, S, 5 + fs, final_actvn=act)
    x = torch.randn(B, feat_nc, fs, fs, device=DEV)
    tgt = torch.rand(B, 3, S, S, device=DEV)
    res = []
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    try:
        for fused in (False, True):
            nr.FUSED_NET = nr.FUSED_TAILS = fused
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
            net.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_(True)
            img = net(xi)
            ((img - tgt) ** 2).mean().backward()
            res.append((img.detach(), xi.grad, {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
    finally:
        nr.FUSED_NET = nr.FUSED_TAILS = True
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    (i0, g0, p0), (i1, g1, p1) = res
    assert (i0 - i1).abs().max() <= (2e-3 if act else 2e-2 * i0.abs().max())
    assert _cos(g0, g1) >= 0.999, _cos(g0, g1)
    assert (g0 - g1).abs().max() <= 6e-2 * g0.abs().max()
    assert set(p0) == set(p1)
    for k in p0:
        assert _cos(p0[k], p1[k]) >= 0.999, (k, _cos(p0[k], p1[k]))
        assert (p0[k] - p1[k]).abs().max() <= 6e-2 * p0[k].abs().max() + 1e-12, k
    hn.ops.FAULTS.flush()


def test_accumulates_into_existing_grads(hn):
    """fuse_grad_accumulation: the kernels add into the parameters' .grad buffers (e.g. the flat all-reduce bucket)."""
    nr = _nr(hn)
    net = _make(nr, 64, 8, 32, 3)
    x = torch.randn(2, 64, 8, 8, device=DEV)
    net(x).square().mean().backward()
    want = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    for p in net.parameters():
        p.grad = torch.full_like(p, 1.0)
    net.fuse_grad_accumulation(True)
    net(x).square().mean().backward()
    net.fuse_grad_accumulation(False)
    for k, p in net.named_parameters():
        if k in want:
            assert (p.grad - 1.0 - want[k]).abs().max() <= 1e-4 * (1 + want[k].abs().max()), k


def test_partial_requires_grad_and_inference(hn):
    nr = _nr(hn)
    net = _make(nr, 64, 8, 32, 4)
    x = torch.randn(1, 64, 8, 8, device=DEV)
    for p in net.parameters():
        p.requires_grad_(False)
    xi = x.clone().requires_grad_(True)                    # fitting: gradients flow to the input only
    net(xi).sum().backward()
    assert xi.grad is not None and torch.isfinite(xi.grad).all() and all(p.grad is None for p in net.parameters())
    net.feat_layers[1].bias.requires_grad_(True)           # a single bias
    net(x).sum().backward()
    assert net.feat_layers[1].bias.grad is not None and net.feat_layers[1].weight.grad is None
