"""GPU tests of the consumer-side kernels (first pieces; csrc/hn_render2d.cu): the fused tails
of NeuralRenderer's up-sampling blocks against the plain PyTorch statement of the reference modules
(NetWorks/PixelShuffleUpsample.py:36-45, NetWorks/neural_renderer.py:47-50,72-91), values and every gradient."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _nr(hn):
    return importlib.import_module(hn.__name__ + ".neural_renderer")


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 8, 8), (1, 64, 5, 7), (3, 256, 2, 2), (1, 32, 64, 64)])
def test_upsample_tail_matches_pytorch(hn, B, C, H, W):
    nr = _nr(hn)
    torch.manual_seed(B * 1000 + C + H)
    blk = nr.PixelShuffleUpsample(C).to(DEV)
    x = torch.randn(B, C, H, W, device=DEV)
    gy = torch.randn(B, C, 2 * H, 2 * W, device=DEV)
    res = []
    for fused in (False, True):
        nr.FUSED_TAILS = fused
        blk.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        y = blk(xi)
        (y * gy).sum().backward()
        res.append((y.detach(), xi.grad, [p.grad.clone() for p in blk.parameters()]))
    nr.FUSED_TAILS = True
    (y0, gx0, gp0), (y1, gx1, gp1) = res
    assert (y0 - y1).abs().max() <= 1e-5 * (1 + y0.abs().max())
    assert (gx0 - gx1).abs().max() <= 1e-4 * (1 + gx0.abs().max())
    for a, b in zip(gp0, gp1):
        assert (a - b).abs().max() <= 2e-4 * (1 + a.abs().max())


@pytest.mark.parametrize("B,H,W", [(2, 8, 8), (1, 5, 9), (2, 2, 2), (1, 128, 128)])
def test_rgb_upsample_matches_pytorch(hn, B, H, W):
    nr = _nr(hn)
    torch.manual_seed(B + H)
    net = nr.NeuralRenderer(featmap_size=8, img_size=32).to(DEV)
    x = torch.randn(B, 3, H, W, device=DEV)
    gy = torch.randn(B, 3, 2 * H, 2 * W, device=DEV)
    out = []
    for fused in (False, True):
        nr.FUSED_TAILS = fused
        xi = x.clone().requires_grad_(True)
        y = net._rgb_up(xi)
        (y * gy).sum().backward()
        out.append((y.detach(), xi.grad))
    nr.FUSED_TAILS = True
    assert (out[0][0] - out[1][0]).abs().max() <= 1e-5 * (1 + out[0][0].abs().max())
    assert (out[0][1] - out[1][1]).abs().max() <= 1e-4 * (1 + out[0][1].abs().max())


@pytest.mark.parametrize("fs,S", [(8, 32), (16, 128), (32, 512)])
def test_neural_renderer_fused_vs_plain(hn, fs, S):
    """The whole consumer, fused tails (library convolutions) against the plain modules: image and every gradient (input
    feature map, all parameters).  The one-call renderer (FUSED_NET) has its own tests in test_gpu_nr.py."""
    nr = _nr(hn)
    nr.FUSED_NET = False
    torch.manual_seed(fs)
    net = nr.NeuralRenderer(featmap_size=fs, img_size=S).to(DEV)
    x = torch.randn(2, 256, fs, fs, device=DEV)
    tgt = torch.rand(2, 3, S, S, device=DEV)
    res = []
    for fused in (False, True):
        nr.FUSED_TAILS = fused
        net.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        img = net(xi)
        ((img - tgt) ** 2).mean().backward()
        res.append((img.detach(), xi.grad, {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}))
    nr.FUSED_TAILS = nr.FUSED_NET = True
    (i0, g0, p0), (i1, g1, p1) = res
    assert (i0 - i1).abs().max() <= 2e-5
    assert (g0 - g1).abs().max() <= 1e-3 * g0.abs().max()
    assert set(p0) == set(p1)
    for k in p0:
        assert (p0[k] - p1[k]).abs().max() <= 2e-3 * (p0[k].abs().max() + 1e-12), k


@pytest.mark.parametrize("B,fs,C", [(2, 8, 256), (1, 5, 96), (3, 64, 256)])
def test_merge_kernel_matches_pytorch(hn, B, fs, C):
    """hn_merge_fwd / hn_merge_bwd against merge = F.permute + bg_alpha * bg_featmap (NetWorks/HeadNeRFNet.py:103-113)."""
    torch.manual_seed(fs)
    n_r = fs * fs
    Fm, bg, feat = torch.randn(B, n_r, C, device=DEV), torch.rand(B, n_r, device=DEV), torch.randn(1, C, fs, fs, device=DEV)
    gout = torch.randn(B, C, fs, fs, device=DEV)
    res = []
    for fused in (False, True):
        xs = [t.clone().requires_grad_(True) for t in (Fm, bg, feat)]
        if fused:
            out = hn.ops.MergeFunction.apply(*xs)
        else:
            out = xs[0].permute(0, 2, 1).reshape(B, C, fs, fs) + xs[1].view(B, 1, fs, fs) * xs[2]
        (out * gout).sum().backward()
        res.append((out.detach(), [t.grad for t in xs]))
    assert torch.equal(res[0][0], res[1][0]) or (res[0][0] - res[1][0]).abs().max() <= 1e-6
    for a, b in zip(res[0][1], res[1][1]):
        assert (a - b).abs().max() <= 1e-4 * (1 + a.abs().max())


def test_consumer_graph_capture_matches_eager(hn):
    """capture_consumer_graph(): NeuralRenderer forward + backward replayed from CUDA graphs give the eager images and gradients."""
    from oracle import headnerf_oracle as O
    torch.manual_seed(1)
    opt = O.OracleOptions(featmap_size=16, pred_img_size=64)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 16, "featmap_nc": 256, "pred_img_size": 64}), False, False).to(DEV)
    x = {k: v.to(DEV) for k, v in O.synthetic_inputs(opt, 2, seed=3).items()}
    tgt = torch.rand(2, 3, 64, 64, device=DEV)

    def step():
        net.zero_grad(set_to_none=True)
        out = net("test", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"], x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        loss = ((out["coarse_dict"]["merge_img"] - tgt) ** 2).mean() + ((out["coarse_dict"]["bg_img"] - 1.0) ** 2).mean()
        loss.backward()
        return out["coarse_dict"]["merge_img"].detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}

    img0, g0 = step()
    keys0 = list(net.state_dict().keys())
    net.capture_consumer_graph(2)
    for _ in range(2):                                            # replay twice: static buffers must be refreshed every time
        img1, g1 = step()
    # while captured: the module tree and the checkpoint layout are untouched, and every call the graphs were NOT captured for
    # (no autograd, other batch sizes, the background map alone) runs the eager module with the right shapes
    assert list(net.state_dict().keys()) == keys0 and not any("_consumer_graph" in k for k in keys0)
    call = lambda xs: net("test", xs["batch_xy"], None, xs["audiostyle"], None, xs["shape_code"], xs["appea_code"], xs["batch_Rmats"], xs["batch_Tvecs"], xs["batch_inv_inmats"])
    with torch.no_grad():                                         # validation without .eval(): training flag still matches the capture
        out = call(x)
    assert out["coarse_dict"]["merge_img"].shape == (2, 3, 64, 64) and out["coarse_dict"]["bg_img"].shape == (1, 3, 64, 64)
    assert (out["coarse_dict"]["merge_img"] - img0).abs().max() <= 1e-5
    x3 = {k: v.to(DEV) for k, v in O.synthetic_inputs(opt, 3, seed=3).items()}
    assert call(x3)["coarse_dict"]["merge_img"].shape == (3, 3, 64, 64)       # another batch size, autograd on: eager, no size mismatch
    net.release_consumer_graph()
    assert "forward" not in net.neural_render.__dict__           # the dispatcher is gone: plain eager launches again
    img2, g2 = step()
    assert (img0 - img1).abs().max() <= 1e-5 and (img0 - img2).abs().max() <= 1e-5
    assert set(g0) == set(g1) == set(g2)
    for k in g0:
        assert (g0[k] - g1[k]).abs().max() <= 2e-3 * (g0[k].abs().max() + 1e-12), k
        assert (g0[k] - g2[k]).abs().max() <= 2e-3 * (g0[k].abs().max() + 1e-12), k
