"""Prints per-leaf gradient cosines (CUDA vs fp32 oracle) for several loss-scale targets — precision study."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import headnerf_oracle as O
from _util import LEAVES, cosine, golden_loss, load_golden
hn = importlib.import_module("nerf-3dtalker-code_b200")
DEV = "cuda:0"
for name in ("fs8_test_init", "fs16_test_trained"):
    g = load_golden(name)
    opt = g["opt"]
    sd = {k: v.requires_grad_(not k.endswith(".f")) for k, v in O.formula_state_dict(opt, g["variant"]).items()}
    x = {k: v.clone().requires_grad_(k in LEAVES) for k, v in g["inp"].items()}
    res, _ = O.headnerf_forward(sd, opt, g["mode"], x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"],
                                x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"], t_rand=x.get("t_rand"))
    golden_loss(res["coarse_dict"]["merge_img"]).backward()
    ref = {k: x[k].grad for k in LEAVES}
    for target in (1.0, 64.0, 4096.0, 32768.0):
        net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": opt.featmap_size, "featmap_nc": 256, "pred_img_size": opt.pred_img_size}), False, False)
        net.load_state_dict(O.formula_state_dict(opt, g["variant"]))
        net = net.to(DEV).eval()
        net.grad_target = target
        xc = {k: v.to(DEV).requires_grad_(k in LEAVES) for k, v in g["inp"].items()}
        out = net(g["mode"], xc["batch_xy"], None, xc["audiostyle"], None, xc["shape_code"], xc["appea_code"],
                  xc["batch_Rmats"], xc["batch_Tvecs"], xc["batch_inv_inmats"])
        golden_loss(out["coarse_dict"]["merge_img"]).backward()
        print(name, "target", target, " ".join(f"{k}={cosine(xc[k].grad.cpu(), ref[k]):.5f}" for k in LEAVES), flush=True)
