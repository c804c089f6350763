"""GPU parity tests (forward): CUDA library through the C ABI vs the oracle and the golden fixtures.
Tolerances are the north-star gates: feature map max-abs-err <= 1e-3, image PSNR >= 45 dB."""
import pytest
import torch

from oracle import headnerf_oracle as O
from _util import GOLDEN_CASES, load_golden, psnr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cuda(d):
    return {k: v.to(DEV) for k, v in d.items()}


@pytest.mark.parametrize("ns,jitter", [(64, False), (64, True), (32, True), (128, False)])
def test_sample_rays_matches_oracle(hn, ns, jitter):
    opt = O.OracleOptions(featmap_size=16, pred_img_size=64, num_sample_coarse=ns)
    inp = O.synthetic_inputs(opt, 2, seed=4, jitter=jitter)
    ro, rd, rl = O.gen_rays(inp["batch_xy"], inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    ref = O.sample_points(ro, rd, rl, opt, jitter, inp.get("t_rand"))
    g = _cuda(inp)
    out = hn.ops.sample_rays(g["batch_xy"], g["batch_Rmats"], g["batch_Tvecs"], g["batch_inv_inmats"],
                             g.get("t_rand"), ns, opt.world_z1, opt.world_z2)
    M = 2 * 256 * ns
    pts_ref = ref["pts"].permute(0, 2, 3, 1).reshape(M, 3)
    assert (out["pts"].cpu() - pts_ref).abs().max() < 2e-5
    assert (out["zvals"].cpu() - ref["zvals"].reshape(M)).abs().max() < 2e-5
    assert (out["z_dists"].cpu() - ref["z_dists"].reshape(M)).abs().max() < 1e-6
    assert (out["ray_d"].cpu() - rd.permute(0, 2, 1).reshape(-1, 3)).abs().max() < 1e-6


@pytest.mark.parametrize("ns,C", [(64, 256), (32, 256), (128, 256), (64, 128)])
def test_composite_forward_backward(hn, ns, C):
    torch.manual_seed(0)
    R = 96
    feat = torch.randn(1, C, R, ns, dtype=torch.float64)
    sigma = torch.relu(torch.randn(1, 1, R, ns, dtype=torch.float64) * 8 + 2)
    delta = torch.rand(1, 1, R, ns, dtype=torch.float64) * 0.05 + 0.08
    zvals = torch.rand(1, 1, R, ns, dtype=torch.float64) * 6
    leaves = [t.requires_grad_(True) for t in (feat, sigma, delta)]
    Fm, bg, depth, w = O.composite(feat, sigma, delta, zvals)
    gF, gbg, gd = torch.randn_like(Fm), torch.randn_like(bg), torch.randn_like(depth)
    (Fm * gF).sum().add((bg * gbg).sum()).add((depth * gd).sum()).backward()

    to = lambda t, c: t.detach().permute(0, 2, 3, 1).reshape(R * ns, c).float().to(DEV).contiguous()
    f_c = to(feat, C).requires_grad_(True)
    s_c = to(sigma, 1).reshape(-1).requires_grad_(True)
    d_c = to(delta, 1).reshape(-1).requires_grad_(True)
    z_c = to(zvals, 1).reshape(-1)
    F2, bg2, dp2 = hn.ops.composite(f_c, s_c, d_c, z_c, ns)
    assert (F2.cpu() - Fm[0].t().float()).abs().max() < 2e-5
    assert (bg2.cpu() - bg[0, 0].float()).abs().max() < 2e-6
    assert (dp2.cpu() - depth[0, 0].float()).abs().max() < 2e-5
    loss = (F2 * gF[0].t().float().to(DEV)).sum() + (bg2 * gbg[0, 0].float().to(DEV)).sum() + (dp2 * gd[0, 0].float().to(DEV)).sum()
    loss.backward()
    ref_df = to(feat.grad, C).cpu()
    assert (f_c.grad.cpu() - ref_df).abs().max() < 1e-5 * (1 + ref_df.abs().max())
    for got, ref in ((s_c.grad, sigma.grad), (d_c.grad, delta.grad)):
        ref = ref.reshape(-1).float()
        assert (got.cpu() - ref).abs().max() < 2e-4 * (1 + ref.abs().max())


def _render_golden(hn, name, train_override=None):
    g = load_golden(name)
    opt = g["opt"]
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": opt.featmap_size, "featmap_nc": 256, "pred_img_size": opt.pred_img_size}),
                         include_vd=False, hier_sampling=False)
    net.load_state_dict(O.formula_state_dict(opt, g["variant"]), strict=True)
    net = net.to(DEV).eval()
    return g, net, _cuda(g["inp"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_feature_map_matches_reference_golden(hn, name):
    """The product's default (precision="auto") against the real reference's outputs: the ABSOLUTE north-star gate, max-abs-err
    <= 1e-3 on F and bg_alpha, on every fixture - random-init weights and the "trained" stress fixtures (output layers x6 / x12,
    max|F| 3-5) alike.  "auto" measures on the caller's inputs whether the single-pass kernels hold the gate for this checkpoint
    (HeadNeRFNet._calibrate) and runs the split-operand kernels when they do not."""
    g, net, x = _render_golden(hn, name)
    assert net.precision == "auto"
    with torch.no_grad():
        Fm, bg = net.render_rays(g["mode"], x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"],
                                 x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"], t_rand=x.get("t_rand"))
    net.check_faults()
    F_ref = g["out"]["F"].permute(0, 2, 1)                      # [B,N_r,C]
    errF = (Fm.cpu() - F_ref).abs().max().item()
    errA = (bg.cpu() - g["out"]["bg_alpha"][:, 0]).abs().max().item()
    print(f"{name}: precision auto -> {net.last_meta['precision']} (probe error {net._calib[3]:.2e}); F max-abs-err {errF:.2e} "
          f"(|F|max {F_ref.abs().max():.2f}), bg_alpha err {errA:.2e}")
    assert errF <= 1e-3 and errA <= 1e-3
    if g["variant"] == "init":
        assert net.last_meta["precision"] == "fast"            # random-init weights: the single-pass kernels hold the gate


def test_single_pass_kernels_hold_the_gate_on_init_weights(hn):
    """precision="fast" forced (the kernels bench.py times): absolute gate on the random-init fixture."""
    g, net, x = _render_golden(hn, "fs8_test_init")
    net.precision = "fast"
    with torch.no_grad():
        Fm, bg = net.render_rays(g["mode"], x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"],
                                 x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"], t_rand=x.get("t_rand"))
    net.check_faults()
    assert (Fm.cpu() - g["out"]["F"].permute(0, 2, 1)).abs().max().item() <= 1e-3
    assert (bg.cpu() - g["out"]["bg_alpha"][:, 0]).abs().max().item() <= 1e-3


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_image_psnr_vs_reference_golden(hn, name):
    g, net, x = _render_golden(hn, name)
    if g["mode"] == "train":
        pytest.skip("HeadNeRFNet.forward draws its own jitter; covered by the feature-map test with explicit t_rand")
    with torch.no_grad():
        out = net(g["mode"], x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"],
                  x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
    assert set(out["coarse_dict"].keys()) == {"merge_img", "bg_img"}
    p = psnr(out["coarse_dict"]["merge_img"], g["out"]["merge_img"])
    pb = psnr(out["coarse_dict"]["bg_img"], g["out"]["bg_img"])
    print(f"{name}: merge_img PSNR {p:.1f} dB, bg_img PSNR {pb:.1f} dB")
    assert p >= 45.0 and pb >= 45.0


def test_intermediate_activations_match_oracle(hn):
    """Layer-by-layer check of the saved operand images and ReLU masks against the oracle's hidden states."""
    g, net, x = _render_golden(hn, "fs8_test_init")
    opt = g["opt"]
    sd = O.formula_state_dict(opt, g["variant"])
    inp = g["inp"]
    # oracle hidden states
    ro, rd, rl = O.gen_rays(inp["batch_xy"], inp["batch_Rmats"], inp["batch_Tvecs"], inp["batch_inv_inmats"])
    smp = O.sample_points(ro, rd, rl, opt, False)
    pe = O.positional_encoding(smp["pts"])
    B, n_r, ns = g["B"], 64, 64
    ex = lambda c: c.unsqueeze(-1).unsqueeze(-1).expand(-1, -1, n_r, ns)
    vps = torch.cat([pe, ex(inp["shape_code"])], 1)
    h = torch.cat([vps, ex(inp["audiostyle"])], 1)
    hs = []
    import torch.nn.functional as F
    for i in range(8):
        h = F.relu(F.conv2d(h, sd[f"fg_CD_predictor.FeaExt_module_{i}.weight"], sd[f"fg_CD_predictor.FeaExt_module_{i}.bias"]))
        hs.append(h)
        if i == 4:
            h = torch.cat([vps, h], 1)
    flat = lambda t: t.permute(0, 2, 3, 1).reshape(B * n_r * ns, -1)
    xs = {k: v.clone().requires_grad_(k == "shape_code") for k, v in x.items()}      # forces the save path
    Fm, bg = net.render_rays("test", xs["batch_xy"], xs["audiostyle"], xs["shape_code"], xs["appea_code"],
                             xs["batch_Rmats"], xs["batch_Tvecs"], xs["batch_inv_inmats"])
    # the RenderFunction node sits behind a few view ops; find it by name
    node = Fm.grad_fn
    seen = []
    stack = [node]
    act = masks = None
    while stack:
        n = stack.pop()
        if n is None or n in seen:
            continue
        seen.append(n)
        if "RenderFunction" in type(n).__name__:
            saved = n.saved_tensors
            act, masks = saved[9], saved[10]
            break
        stack.extend(f for f, _ in n.next_functions)
    assert act is not None
    M = B * n_r * ns
    n_tiles = M // 128
    pe_img = hn.ops.decode_image(act, 0, 1, n_tiles).cpu()
    assert (pe_img[:, :63] - flat(pe)).abs().max() < 2e-3          # fp16 rounding of |pts| <= 4
    for i in range(8):
        img = hn.ops.decode_image(act, 1 + 6 * i, 6, n_tiles).cpu()
        ref = flat(hs[i])
        err = (img - ref).abs().max().item()
        print(f"h{i}: max-abs-err {err:.2e} (max {ref.abs().max():.2f})")
        assert err < 5e-3 * (1 + ref.abs().max().item())
        m = hn.ops.decode_masks(masks, 12 * i, 384, M).cpu()
        disagree = (m != (ref > 0)) & (ref.abs() > 1e-2)
        assert disagree.sum() == 0


def test_fold_bias_kernel_matches_folding_algebra(hn):
    """hn_fold_bias / hn_fold_bias_bwd against the plain-PyTorch statement of the folding (HeadNeRFNet._fold_biases, itself
    checked against the reference concat orders on CPU): values and every gradient (codes, folded weight columns, biases)."""
    torch.manual_seed(3)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 8, "featmap_nc": 256, "pred_img_size": 32}), False, False).to(DEV)
    B = 3
    codes = [torch.randn(B, n, device=DEV) for n in (179, 127, 64)]                # shape, appea, audio
    gout = torch.randn(B, hn._lib.BIAS_STRIDE, device=DEV)
    gout[:, hn._lib.BIAS_OFF_DENSITY + 1:] = 0                                      # padding carries no gradient
    res = []
    for fn in ("torch", "cuda"):
        net.zero_grad(set_to_none=True)
        xs = [c.clone().requires_grad_(True) for c in codes]
        bias = net._fold_biases(*xs) if fn == "torch" else net._fold_biases_cuda(*xs, None)
        (bias * gout).sum().backward()
        res.append((bias.detach(), [x.grad for x in xs], {k: p.grad.clone() for k, p in net.fg_CD_predictor.named_parameters() if p.grad is not None}))
    (b0, g0, p0), (b1, g1, p1) = res
    assert (b0 - b1).abs().max() < 1e-5
    for a, b in zip(g0, g1):
        assert (a - b).abs().max() < 1e-4 * (1 + a.abs().max())
    assert set(p0) == set(p1)
    for k in p0:
        assert (p0[k] - p1[k]).abs().max() < 1e-4 * (1 + p0[k].abs().max()), k


def test_loss_scale_kernel(hn):
    torch.manual_seed(0)
    for n, mx in ((1000003, 3.7e-4), (256, 9.0), (5, 1e-20)):
        g = torch.randn(n, device=DEV).clamp(-1, 1) * mx * 0.5
        g[n // 3] = -mx
        for _ in range(2):                                                          # the scratch words are reused: call twice
            s = hn.ops.loss_scale(g, 64.0)
        want = 2.0 ** torch.floor(torch.log2(torch.tensor(64.0 / max(mx, 1e-30))))
        assert float(s) == float(want), (n, float(s), float(want))
