"""GPU tests of the training-step pieces either side of the path, through the C ABI:
fused photometric loss vs the oracle's restatement of Utils/HeadNeRFLossUtils.py (itself pinned bit-equal to the reference),
fused Adam vs torch.optim.Adam, the reference checkpoint layout round trip, Audio2style on the GPU vs the oracle."""
import io

import pytest
import torch

from oracle import headnerf_oracle as O
from _util import cosine

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,S,bg_type,with_nan", [(2, 64, "white", False), (3, 37, "black", False), (1, 512, "white", True)])
def test_photo_loss_matches_oracle(hn, B, S, bg_type, with_nan):
    g = torch.Generator().manual_seed(B * 100 + S)
    img = torch.rand(B, 3, S, S, generator=g)
    if with_nan:
        img.view(-1)[::1013] = float("nan")
    bg, gt, mask = torch.rand(1, 3, S, S, generator=g), torch.rand(B, 3, S, S, generator=g), torch.rand(B, 1, S, S, generator=g)
    mask.view(-1)[::7] = 0.5                                     # the boundary value belongs to the head (>= 0.5)
    bgv = 1.0 if bg_type == "white" else 0.0
    o_img, o_bg = img.clone().double().requires_grad_(True), bg.clone().double().requires_grad_(True)
    ref = O.data_loss(o_img, o_bg, gt.double(), mask.double(), bgv)
    ref["total_loss"].backward()
    lu = hn.HeadNeRFLossUtils(bg_type=bg_type, use_vgg_loss=False, device=DEV)
    c_img, c_bg = img.to(DEV).requires_grad_(True), bg.to(DEV).requires_grad_(True)
    out = lu.calc_total_loss(None, None, {"coarse_dict": {"merge_img": c_img, "bg_img": c_bg}}, gt.to(DEV), mask.to(DEV), None)
    assert set(out.keys()) == {"bg_loss", "head_loss", "nonhaed_loss", "total_loss"}
    for k in out:
        assert abs(float(out[k]) - float(ref[k])) <= 2e-6 * max(1.0, abs(float(ref[k]))), (k, float(out[k]), float(ref[k]))
    out["total_loss"].backward()
    assert (c_img.grad.cpu().double() - o_img.grad).abs().max() <= 1e-6 * o_img.grad.abs().max()
    assert (c_bg.grad.cpu().double() - o_bg.grad).abs().max() <= 1e-6 * o_bg.grad.abs().max()
    # determinism: the fixed-order fold gives bit-identical terms run to run
    again = lu.calc_total_loss(None, None, {"coarse_dict": {"merge_img": c_img.detach(), "bg_img": c_bg.detach()}}, gt.to(DEV), mask.to(DEV), None)
    assert all(torch.equal(again[k], out[k].detach()) for k in out)
    # calc_data_loss: the three terms separately differentiable, boolean masks as the reference passes them
    d = lu.calc_data_loss({"merge_img": c_img.detach().requires_grad_(True), "bg_img": c_bg.detach()}, gt.to(DEV),
                          (mask >= 0.5).to(DEV), (mask < 0.5).to(DEV))
    assert abs(float(d["head_loss"]) - float(ref["head_loss"])) <= 2e-6 * max(1.0, abs(float(ref["head_loss"])))


def test_photo_loss_has_no_cpu_path(hn):
    lu = hn.HeadNeRFLossUtils(bg_type="white", use_vgg_loss=False)
    with pytest.raises(hn._lib.HeadNeRFLibraryError):
        lu.calc_total_loss(None, None, {"coarse_dict": {"merge_img": torch.rand(1, 3, 8, 8), "bg_img": torch.rand(1, 3, 8, 8)}},
                           torch.rand(1, 3, 8, 8), torch.rand(1, 1, 8, 8), None)
    with pytest.raises(NotImplementedError):
        hn.HeadNeRFLossUtils(use_vgg_loss=True)


@pytest.mark.parametrize("weight_decay", [0.0, 1e-2])
def test_fused_adam_matches_torch_adam(hn, weight_decay):
    torch.manual_seed(0)
    shapes = [(384, 306, 1, 1), (384,), (1, 384, 1, 1), (1,), (3, 7), (5,)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, device=DEV)) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = torch.optim.Adam(ref_p, lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay, foreach=False, fused=False)
    ours = hn.FusedAdam(our_p, lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    sched_r = torch.optim.lr_scheduler.StepLR(ref, step_size=3, gamma=0.5)      # talker_trainer.py:726-727
    sched_o = torch.optim.lr_scheduler.StepLR(ours, step_size=3, gamma=0.5)
    v0 = [p._version for p in our_p]
    for it in range(7):
        gs = [torch.randn_like(p) * (0.1 + it) for p in ref_p]
        ref.zero_grad(); ours.zero_grad()
        for p, q, g in zip(ref_p, our_p, gs):
            p.grad = g.clone()
            q.grad.copy_(g)                                      # the flat views stay in place
        ref.step(); ours.step()
        sched_r.step(); sched_o.step()
    for p, q in zip(ref_p, our_p):
        assert (p - q).abs().max() <= 2e-6 * (1 + p.abs().max()), float((p - q).abs().max())
    assert all(p._version > v for p, v in zip(our_p, v0))       # caches keyed on the version counter see the update
    # torch.optim.Adam's state-dict format both ways (the checkpoints' "optim_state")
    sd_r, sd_o = ref.state_dict(), ours.state_dict()
    assert sd_r["param_groups"][0]["lr"] == sd_o["param_groups"][0]["lr"]
    for i in sd_r["state"]:
        assert float(sd_r["state"][i]["step"]) == float(sd_o["state"][i]["step"]) == 7.0
        assert (sd_r["state"][i]["exp_avg"] - sd_o["state"][i]["exp_avg"]).abs().max() <= 1e-6 * (1 + sd_r["state"][i]["exp_avg"].abs().max())
    fresh_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    fresh = hn.FusedAdam(fresh_p, lr=1.0, weight_decay=weight_decay)
    fresh.load_state_dict(sd_r)                                  # a torch.optim.Adam checkpoint continues on the fused optimizer
    gs = [torch.randn_like(p) for p in ref_p]
    for p, q, g in zip(ref_p, fresh_p, gs):
        p.grad = g.clone()
        q.grad.copy_(g)
    ref.step(); fresh.step()
    for p, q in zip(ref_p, fresh_p):
        assert (p - q).abs().max() <= 2e-6 * (1 + p.abs().max())


def test_fused_adam_grad_scale_is_the_data_parallel_average(hn):
    torch.manual_seed(1)
    p0 = torch.randn(1000, device=DEV)
    a, b = torch.nn.Parameter(p0.clone()), torch.nn.Parameter(p0.clone())
    oa, ob = hn.FusedAdam([a], lr=1e-2), hn.FusedAdam([b], lr=1e-2)
    g = torch.randn(1000, device=DEV)
    a.grad.copy_(g * 8); b.grad.copy_(g)
    oa.step(grad_scale=1.0 / 8); ob.step()
    assert (a - b).abs().max() <= 1e-6


def test_training_step_with_fused_optimizer_updates_the_rendered_image(hn):
    """Whole step on the drop-in module: forward, fused loss, backward, FusedAdam over the model's flat buffers - and the NEXT
    forward must see the new weights (the packed-operand cache follows the version counters FusedAdam bumps)."""
    opt = O.OracleOptions(featmap_size=8, pred_img_size=32)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 8, "featmap_nc": 256, "pred_img_size": 32}), False, False)
    net.load_state_dict(O.formula_state_dict(opt, "init"), strict=True)
    net = net.to(DEV).train()
    keys0 = list(net.state_dict().keys())
    optim = hn.FusedAdam(net.parameters(), lr=1e-3)
    lu = hn.HeadNeRFLossUtils(use_vgg_loss=False, device=DEV)
    x = {k: v.to(DEV) for k, v in O.synthetic_inputs(opt, 2, seed=1).items()}
    gen = torch.Generator().manual_seed(0)
    gt, mask = torch.rand(2, 3, 32, 32, generator=gen).to(DEV), (torch.rand(2, 1, 32, 32, generator=gen) > 0.5).float().to(DEV)
    losses = []
    for it in range(4):
        optim.zero_grad()
        out = net("test", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"],
                  x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        loss = lu.calc_total_loss(None, None, out, gt, mask, None)["total_loss"]
        loss.backward()
        optim.step()
        losses.append(float(loss))
    net.check_faults()
    assert losses[-1] < losses[0], losses                         # the optimizer's updates reach the kernels' packed operands
    assert list(net.state_dict().keys()) == keys0


def test_checkpoint_round_trip(hn, tmp_path):
    """{"para", "net", "audio2style", "optim_state", "scheule_state", "epoch"} (talker_trainer.py:915-936) written by the drop-in,
    read back the way FittingSingleImage_new.py:640-656 does, rendering the same image; and the reference loader's own idiom
    `model.state_dict()[k].data.copy_()` (talker_trainer.py:557-567) on a live CUDA module must not leave stale packed weights."""
    opt = O.OracleOptions(featmap_size=8, pred_img_size=32)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 8, "featmap_nc": 256, "pred_img_size": 32}), False, False)
    sd_a = O.formula_state_dict(opt, "init")
    net.load_state_dict(sd_a, strict=True)
    net = net.to(DEV).eval()
    a2s = hn.Audio2style().to(DEV)
    optim = hn.FusedAdam(net.parameters(), lr=1e-4)
    sched = torch.optim.lr_scheduler.StepLR(optim, step_size=10, gamma=0.5)
    path = str(tmp_path / "ck.pth")
    ck = hn.save_checkpoint(path, net, epoch=3, audio2style=a2s, optimizer=optim, scheduler=sched)
    assert set(ck.keys()) == {"epoch", "net", "para", "audio2style", "optim_state", "scheule_state"}
    assert ck["para"] == {"featmap_size": 8, "featmap_nc": 256, "pred_img_size": 32}
    assert list(ck["net"].keys()) == list(O.state_dict_shapes(opt).keys())      # the reference's key order and names
    net2, a2s2, raw = hn.load_checkpoint(path, device=DEV)
    assert raw["epoch"] == 3 and a2s2 is not None
    x = {k: v.to(DEV) for k, v in O.synthetic_inputs(opt, 1, seed=2).items()}
    call = lambda n: n("test", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"],
                       x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])["coarse_dict"]["merge_img"]
    feats = lambda n: n.render_rays("test", x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"],
                                    x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
    with torch.no_grad():
        (F1, b1), (F2, b2) = feats(net), feats(net2)
        assert torch.equal(F1, F2) and torch.equal(b1, b2)       # the hot path is deterministic: same weights, same bits
        img1, img2 = call(net), call(net2)
        # the consumer's convolutions are library (cuDNN, TF32 allowed as in the reference) calls whose algorithm choice may follow
        # the parameters' addresses (net's live in FusedAdam's flat buffer): images agree to the TF32 level, not bit for bit
        assert (img1 - img2).abs().max() < 2e-3
        # mid-run reload through .data (no version bump): the next forward must render the NEW weights
        sd_b = O.formula_state_dict(opt, "trained")
        for k in net.state_dict():
            net.state_dict()[k].data.copy_(sd_b[k].data)
        img3 = call(net)
        ref, _ = O.headnerf_forward(sd_b, opt, "test", *[O.synthetic_inputs(opt, 1, seed=2)[k] for k in
                                    ("batch_xy", "audiostyle", "shape_code", "appea_code", "batch_Rmats", "batch_Tvecs", "batch_inv_inmats")])
    assert (img3.cpu() - ref["coarse_dict"]["merge_img"]).abs().max() < 2e-3
    assert (img3 - img1).abs().max() > 1e-3
    # p.data writes WITHOUT a state_dict() call need the explicit hook
    with torch.no_grad():
        for k, p in net.named_parameters():
            p.data.copy_(sd_a[k].to(DEV))
        net.invalidate_caches()
        assert torch.equal(feats(net)[0], F1) and torch.equal(call(net), img1)


def test_audio2style_on_gpu_matches_oracle(hn):
    torch.manual_seed(5)
    m = hn.Audio2style().eval()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn(7, 80, 16, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = O.audio2style_forward(sd, x)
        got = m.to(DEV)(x.to(DEV))
    assert got.shape == (7, 64)
    assert (got.cpu() - want).abs().max() <= 1e-4 * (1 + want.abs().max())       # cuDNN LSTM may use TF32-free fp32 but fuses differently
    assert cosine(got, want) > 0.99999
