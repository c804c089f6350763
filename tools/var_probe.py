import contextlib, importlib, os, sys, time
import torch
sys.path.insert(0, os.getcwd())
from oracle import headnerf_oracle as O
hn = importlib.import_module("nerf-3dtalker-code_b200")
dev = "cuda:0"
opt = O.OracleOptions(featmap_size=64, pred_img_size=512)
net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 64, "featmap_nc": 256, "pred_img_size": 512}), False, False).to(dev)
x = {k: v.to(dev) for k, v in O.synthetic_inputs(opt, 2, seed=0).items()}
ts = []
with torch.no_grad():
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        net.render_rays("test", x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"], x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        if ts[-1] > 100: print('slow call', i, 'status', net.last_meta['last_status'].cpu()[:2].tolist())
st = net.last_meta["last_status"].cpu()[:2].tolist()
print("status", st, " ms:", " ".join(f"{t:.2f}" for t in ts))
