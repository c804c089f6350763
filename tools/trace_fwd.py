"""Timeline of one tile of the fused forward kernel (HN_TRACE build): per weight unit, when the MMA issuer started, finished its
operand/accumulator waits, got the weights, and committed; per accumulator chunk, when its epilogue saw it, released it, and
signalled its output.   usage: HN_TRACE=1 HN_LIB_PATH=build/libhn_TRACE.so python tools/trace_fwd.py [infer|train]"""
import contextlib, importlib, os, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import headnerf_oracle as O
hn = importlib.import_module("nerf-3dtalker-code_b200")
mode = sys.argv[1] if len(sys.argv) > 1 else "infer"
dev = "cuda:0"
opt = O.OracleOptions(featmap_size=64, pred_img_size=512)
net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": 64, "featmap_nc": 256, "pred_img_size": 512}), False, False).to(dev)
net.precision = "fast"
x = {k: v.to(dev) for k, v in O.synthetic_inputs(opt, 2, seed=0).items()}
with (torch.no_grad() if mode == "infer" else contextlib.nullcontext()):
    for _ in range(3):
        net.render_rays("test", x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"], x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
torch.cuda.synchronize()
st = net.last_meta["last_status"].cpu().tolist()
tr = st[64:]
t0 = tr[0]
rel = lambda v: (v - t0) & 0xFFFFFFFF
print("stage | start  wait_ops  wait_w  issue+commit | d(start)")
prev = 0
for u in range(85):
    a, b, c, d = (rel(tr[u * 4 + i]) for i in range(4))
    print(f"{u:4d} | {a:7d} {b - a:6d} {c - b:6d} {d - c:6d} | {a - prev:5d}")
    prev = a
print("chunk | acc_full_seen  released(+)  signalled(+)")
for e in range(31):
    a, b, c = (rel(tr[1024 + e * 4 + i]) for i in range(3))
    print(f"{e:3d} | {a:8d} {b - a:6d} {c - a:6d}")
