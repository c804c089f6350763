"""Throughput of the remaining BASELINE.json configurations on one B200 (device-resident inputs, CUDA events):
  config 3 : Reso32HR training step (hot path + NeuralRenderer + MSE + Adam), batch 2
  config 4 : FittingSingleImage-style latent/camera optimisation loop, 500 iterations, Reso32, frozen weights
  config 5 : forward-only sweep, 16K..4M rays x 32/64/128 samples
Writes a markdown table to stdout (committed under profiles/)."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import headnerf_oracle as O
hn = importlib.import_module("nerf-3dtalker-code_b200")
dev = torch.device("cuda", 0)


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def make(fs, S, B, ns=64, seed=0):
    opt = O.OracleOptions(featmap_size=fs, pred_img_size=S, num_sample_coarse=ns)
    bo = hn.BaseOptions({"featmap_size": fs, "featmap_nc": 256, "pred_img_size": S})
    bo.num_sample_coarse = ns
    torch.manual_seed(0)
    net = hn.HeadNeRFNet(bo, False, False).to(dev)
    x = {k: v.to(dev) for k, v in O.synthetic_inputs(opt, B, seed=seed).items()}
    return opt, net, x


print("| config | workload | ms / step | ray*samples/s |")
print("|---|---|---|---|")
# ---- config 3: Reso32HR training step
opt, net, x = make(32, 512, 2)
adam = torch.optim.Adam(net.parameters(), lr=1e-4)
target = torch.rand(2, 3, 512, 512, device=dev)
def train_step():
    adam.zero_grad(set_to_none=False)
    out = net("train", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"], x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
    loss = ((out["coarse_dict"]["merge_img"] - target) ** 2).mean() + ((out["coarse_dict"]["bg_img"] - 1.0) ** 2).mean()
    loss.backward()
    adam.step()
ms = timed(train_step, 20)
M = 2 * 1024 * 64
print(f"| 3 | Reso32HR (32x32 rays x 64, 512 px) full training step incl. NeuralRenderer + MSE + Adam, batch 2 | {ms:.3f} | {M / ms * 1e3:.3e} |")
net.capture_consumer_graph(2)               # NeuralRenderer fwd + bwd as two CUDA graphs
ms = timed(train_step, 20)
print(f"| 3 | same full training step, consumer captured in CUDA graphs (capture_consumer_graph) | {ms:.3f} | {M / ms * 1e3:.3e} |")
net.release_consumer_graph()
def hot_only():
    Fm, bg = net.render_rays("train", x["batch_xy"], x["audiostyle"], x["shape_code"], x["appea_code"], x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
    (Fm.sum() + bg.sum()).backward()
ms = timed(hot_only, 20)
print(f"| 3 | same, hot path only (fwd+bwd) | {ms:.3f} | {M / ms * 1e3:.3e} |")
hn.ops.check_status(net.last_meta["last_status"], "config 3")

# ---- config 4: fitting loop (FittingSingleImage_new.py:826-903 shape), 500 iterations
opt, net, x = make(32, 256, 1)
net.eval()
net.precision = "fast"                      # (the default "auto" would pick "high" here: camera gradients are requested)
for p in net.parameters():
    p.requires_grad_(False)
off = {k: torch.zeros_like(x[k], requires_grad=True) for k in ("shape_code", "appea_code")}
d_euler = torch.zeros(1, 3, device=dev, requires_grad=True)
d_T = torch.zeros(1, 3, 1, device=dev, requires_grad=True)
opt_fit = torch.optim.Adam([{"params": [off["shape_code"]], "lr": 0.015}, {"params": [off["appea_code"]], "lr": 0.01},
                            {"params": [d_euler], "lr": 0.001}, {"params": [d_T], "lr": 0.001}])
target = torch.rand(1, 3, 256, 256, device=dev)
def fit_iter():
    opt_fit.zero_grad()
    dR = O.euler_to_rot(d_euler)
    R = dR @ x["batch_Rmats"]
    T = dR @ x["batch_Tvecs"] + d_T
    out = net("test", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"] + off["shape_code"], x["appea_code"] + off["appea_code"], R, T, x["batch_inv_inmats"])
    ((out["coarse_dict"]["merge_img"] - target) ** 2).mean().backward()
    opt_fit.step()
for _ in range(5):
    fit_iter()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(500):
    fit_iter()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
M = 1024 * 64
print(f"| 4 | fitting loop, 500 iterations, Reso32 (32x32 x 64), grads to codes + camera only | {dt / 500 * 1e3:.3f} (total {dt:.2f} s) | {M * 500 / dt:.3e} |")
hn.ops.check_status(net.last_meta["last_status"], "config 4")
net.precision = "high"                      # split-operand mode: camera-gradient cosine 0.99999 instead of 0.998
for _ in range(5):
    fit_iter()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(500):
    fit_iter()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"| 4 | same loop, precision=\"high\" (hi+lo split operands, fp32 activations) | {dt / 500 * 1e3:.3f} (total {dt:.2f} s) | {M * 500 / dt:.3e} |")
hn.ops.check_status(net.last_meta["last_status"], "config 4 high")
net.precision = "fast"
def fit_hot():
    xs = {k: (x[k] + off[k]) for k in off}
    Fm, bg = net.render_rays("test", x["batch_xy"], x["audiostyle"], xs["shape_code"], xs["appea_code"], x["batch_Rmats"] + 0 * d_euler.sum(), x["batch_Tvecs"] + d_T, x["batch_inv_inmats"])
    (Fm.sum() + bg.sum()).backward()
for prec in ("fast", "high"):
    net.precision = prec
    ms = timed(fit_hot, 50)
    print(f"| 4 | hot path only of one fitting iteration (fwd + bwd to codes and camera), precision={prec} | {ms:.3f} | {M / ms * 1e3:.3e} |")
net.precision = "fast"

# ---- config 5: forward-only sweep
for ns in (32, 64, 128):
    opt, net, x = make(64, 512, 1, ns=ns)
    net.eval()
    for n_rays in (16384, 262144, 1048576, 4194304):
        xy = (torch.rand(1, 2, n_rays, device=dev) * 64.0)
        with torch.no_grad():
            fn = lambda: net.render_rays("test", xy, x["audiostyle"], x["shape_code"], x["appea_code"], x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
            ms = timed(fn, 3 if n_rays >= 1048576 else 10, warm=1)
        print(f"| 5 | forward sweep, {n_rays} rays x {ns} samples | {ms:.3f} | {n_rays * ns / ms * 1e3:.3e} |")
    hn.ops.check_status(net.last_meta["last_status"], "config 5")
