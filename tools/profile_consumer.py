"""Where the consumer (NeuralRenderer fwd+bwd) spends its GPU time: torch profiler table of the
Reso32HR training step (config 3) with the hot path excluded."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hn = importlib.import_module("nerf-3dtalker-code_b200")
dev = torch.device("cuda", 0)
fs, S, B = (int(a) for a in (sys.argv[1:4] + ["32", "512", "2"][len(sys.argv) - 1:]))
net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": fs, "featmap_nc": 256, "pred_img_size": S}), False, False).to(dev)
nr = net.neural_render
x = torch.randn(B, 256, fs, fs, device=dev, requires_grad=True)
def step():
    img = nr(x)
    bg = nr(nr.get_bg_featmap())
    ((img - 0.5) ** 2).mean().add(((bg - 1.0) ** 2).mean()).backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
print(f"NeuralRenderer x2 fwd+bwd, fs={fs} S={S} B={B}: {e0.elapsed_time(e1) / 10:.3f} ms per step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
