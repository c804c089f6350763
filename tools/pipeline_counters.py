"""Prints the forward kernel's CTA-0 pipeline counters (cycles) for the Reso64 batch-2 workload."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import headnerf_oracle as O
hn = importlib.import_module("nerf-3dtalker-code_b200")
dev = "cuda:0"
opt = O.OracleOptions(featmap_size=64, pred_img_size=512)
net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 64, "featmap_nc": 256, "pred_img_size": 512}), False, False).to(dev)
net.precision = "fast"
x = {k: v.to(dev) for k, v in O.synthetic_inputs(opt, 2, seed=0).items()}
for need_grad in (False, True):
    xs = {k: v.clone().requires_grad_(need_grad and k == "shape_code") for k, v in x.items()}
    import contextlib
    with (contextlib.nullcontext() if need_grad else torch.no_grad()):
        for _ in range(3):
            Fm, bg = net.render_rays("test", xs["batch_xy"], xs["audiostyle"], xs["shape_code"], xs["appea_code"], xs["batch_Rmats"], xs["batch_Tvecs"], xs["batch_inv_inmats"])
    torch.cuda.synchronize()
    st = net.last_meta["last_status"].cpu()
    mma = st[2:18].view(torch.int64).tolist()
    epi = st[18:50].view(torch.int64).tolist()
    tiles = (2 * 4096 * 64 // 128 + 147) // 148
    mn = ["total", "wait_pe", "wait_a_ready", "wait_acc_empty", "wait_w_full", "-", "issue_mma", "commit"]
    en = ["total", "wait_acc_full", "tmem_ld+arrive", "-", "convert+stores", "tmem_st_wait", "stage/a_ready arrives", "density", "-", "feat_store"]
    print("saving" if need_grad else "inference", "| cycles per tile:", mma[0] // tiles)
    print("  MMA thread :", {n: f"{v / max(mma[0], 1):.3f}" for n, v in zip(mn, mma)})
    print("  epilogue w0 (cycles/op):", {n: int(v / tiles / 31) for n, v in zip(en, epi)})
