"""Runs the fused forward kernel a few times on the Reso64 batch-2 workload (profiling aid: short, no CPU legs).
   usage: python tools/fwd_only.py [infer|train|bwd] [reps]"""
import contextlib, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import headnerf_oracle as O
hn = importlib.import_module("nerf-3dtalker-code_b200")
mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = "cuda:0"
opt = O.OracleOptions(featmap_size=64, pred_img_size=512)
net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 64, "featmap_nc": 256, "pred_img_size": 512}), False, False).to(dev)
x = {k: v.to(dev) for k, v in O.synthetic_inputs(opt, 2, seed=0).items()}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with (torch.no_grad() if mode == "infer" else contextlib.nullcontext()):
    for i in range(reps + 1):
        if i == 1:
            e0.record()
        Fm, bg = net.render_rays("test", x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"], x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        if mode == "bwd":
            (Fm.sum() + bg.sum()).backward()
e1.record()
torch.cuda.synchronize()
hn.ops.check_status(net.last_meta["last_status"], "fwd_only")
print(f"{mode}: {e0.elapsed_time(e1) / reps:.3f} ms per call (fused forward + compositing{' + backward' if mode == 'bwd' else ''})")
