"""GPU test of the one-call C ABI (hn_render_fwd / hn_render_bwd, csrc/hn_render.cu) driven straight through ctypes the way a
C / C++ host would: same feature map as the module path (bit for bit - it launches the same kernels) and the same gradients
for codes, camera, weights and biases (up to the order of the atomics)."""
import ctypes as C

import pytest
import torch

from oracle import headnerf_oracle as O
from _util import cosine, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def test_one_call_render_matches_module(hn):
    L, ops = hn._lib, hn.ops
    lib = L.load()
    g = load_golden("fs16_test_trained")
    opt = g["opt"]
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": opt.featmap_size, "featmap_nc": 256, "pred_img_size": opt.pred_img_size}), False, False)
    net.load_state_dict(O.formula_state_dict(opt, g["variant"]), strict=True)
    net = net.to(DEV).eval()
    net.precision = "fast"
    x = {k: v.to(DEV).contiguous() for k, v in g["inp"].items()}
    B, _, n_r = x["batch_xy"].shape
    ns, M = 64, B * n_r * 64
    leaves = ("shape_code", "appea_code", "audiostyle", "batch_Rmats", "batch_Tvecs")
    gen = torch.Generator().manual_seed(4)
    gF = (torch.randn(B * n_r, 256, generator=gen) * 1e-2).to(DEV)
    gb = (torch.randn(B * n_r, generator=gen) * 1e-2).to(DEV)

    # ---- module path
    xs = {k: (v.clone().requires_grad_(True) if k in leaves else v) for k, v in x.items()}
    Fm, bg = net.render_rays("test", xs["batch_xy"], xs["audiostyle"], xs["shape_code"], xs["appea_code"],
                             xs["batch_Rmats"], xs["batch_Tvecs"], xs["batch_inv_inmats"])
    torch.autograd.backward([Fm.reshape(-1, 256), bg.reshape(-1)], [gF, gb])
    ref_params = {n: p.grad.clone() for n, p in net.fg_CD_predictor.named_parameters()}

    # ---- one-call C ABI
    lay = net.fg_CD_predictor.layers()
    ws = [m.weight.detach() for m in lay]
    bs = [m.bias.detach() for m in lay]
    packed = ops.pack_weights(ws, L.PE + net.shape_dims)
    T3 = x["batch_Tvecs"].reshape(B, 3).contiguous()
    cam = ops._camera(x["batch_xy"], x["batch_Rmats"], T3, x["batch_inv_inmats"], None, ns, net.opt.world_z1, net.opt.world_z2)
    fold = ops.FoldBiasFunction._args(x["shape_code"], x["audiostyle"], x["appea_code"], ws[0], ws[5], ws[10], bs, wr0=ws[9])
    f32 = lambda *s: torch.empty(*s, device=DEV)
    z32 = lambda *s: torch.zeros(*s, device=DEV)
    u8 = lambda n: torch.empty(n, dtype=torch.uint8, device=DEV)
    bias_eff, feat, sigma, delta = f32(B, L.BIAS_STRIDE), f32(M, 256), f32(M), f32(M)
    act, masks = u8(lib.hn_act_bytes(M)), torch.empty(M * L.MASK_WORDS, dtype=torch.int32, device=DEV)
    F2, bg2, status = f32(B * n_r, 256), f32(B * n_r), torch.zeros(64, dtype=torch.int32, device=DEV)
    a = L.RenderFwd()
    a.cam, a.fold = cam, fold
    a.w_density, a.packed = _p(ws[8].reshape(-1).contiguous()), _p(packed)
    a.bias_eff, a.feat, a.sigma, a.delta, a.act, a.masks = _p(bias_eff), _p(feat), _p(sigma), _p(delta), _p(act), _p(masks)
    a.F, a.bg_alpha, a.status = _p(F2), _p(bg2), _p(status)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L.check(lib.hn_render_fwd(C.byref(a), stream), "hn_render_fwd")
    assert torch.equal(F2.view(B, n_r, 256), Fm.detach()) and torch.equal(bg2.view(B, n_r), bg.detach())

    b = L.RenderBwd()
    b.cam, b.fold = cam, fold
    b.w_density, b.packed = a.w_density, a.packed
    b.feat, b.sigma, b.delta, b.act, b.masks = _p(feat), _p(sigma), _p(delta), _p(act), _p(masks)
    b.gF, b.g_bg, b.grad_target = _p(gF), _p(gb), 64.0
    dimg, dsig, ddel, grads = u8(lib.hn_dfeat_image_bytes(M)), f32(M), f32(M), u8(lib.hn_grads_bytes(M))
    scale, scratch, dbias = f32(1), torch.zeros(2, dtype=torch.int32, device=DEV), z32(B, L.BIAS_STRIDE)
    wksp = u8(lib.hn_wgrad_workspace_bytes(B))
    g_o, g_v, g_l = z32(B * n_r, 3), z32(B * n_r, 3), z32(B * n_r)
    b.dfeat_image, b.dsigma, b.ddelta, b.grads = _p(dimg), _p(dsig), _p(ddel), _p(grads)
    b.scale, b.scale_scratch8, b.dbias_eff = _p(scale), _p(scratch), _p(dbias)
    b.items_workspace, b.items_workspace_bytes = _p(wksp), wksp.numel()
    b.g_ray_o, b.g_ray_v, b.g_ray_l = _p(g_o), _p(g_v), _p(g_l)
    dws = [torch.zeros_like(w) for w in ws]
    dbs = [torch.zeros_like(v) for v in bs]
    for i in range(12):
        b.dw[i] = dws[i].data_ptr()
        b.ld[i] = ws[i].numel() // ws[i].shape[0]
        b.fold_grads.dbias[i] = dbs[i].data_ptr()
    b.l5_hidden_col = L.PE + net.shape_dims
    dwf = z32(L.RGB1, L.HIDDEN)                                      # dL/d(W_R1[:, :384] W_R0): RGB_layer_0 is folded into RGB_layer_1
    b.dwf = _p(dwf)
    dshape, daudio, dappea = torch.empty_like(x["shape_code"]), torch.empty_like(x["audiostyle"]), torch.empty_like(x["appea_code"])
    b.fold_grads.dshape, b.fold_grads.daudio, b.fold_grads.dappea = dshape.data_ptr(), daudio.data_ptr(), dappea.data_ptr()
    b.fold_grads.dw0, b.fold_grads.dw5, b.fold_grads.dwr1 = dws[0].data_ptr(), dws[5].data_ptr(), dws[10].data_ptr()
    dR, dT, dK = z32(B, 3, 3), z32(B, 3), z32(B, 3, 3)
    b.dR, b.dT, b.dKinv, b.status = _p(dR), _p(dT), _p(dK), _p(status)
    L.check(lib.hn_render_bwd(C.byref(b), stream), "hn_render_bwd")
    torch.cuda.synchronize()
    ops.check_status(status, "one-call render")

    for got, key in ((dshape, "shape_code"), (dappea, "appea_code"), (daudio, "audiostyle"), (dR, "batch_Rmats"), (dT, "batch_Tvecs")):
        c = cosine(got, xs[key].grad)
        assert c > 0.99999, (key, c)
    names = ["FeaExt_module_%d" % i for i in range(8)] + ["density_module", "RGB_layer_0", "RGB_layer_1", "RGB_layer_2"]
    for i, n in enumerate(names):
        assert cosine(dws[i], ref_params[n + ".weight"]) > 0.99999, n
        assert cosine(dbs[i], ref_params[n + ".bias"]) > 0.99999, n
