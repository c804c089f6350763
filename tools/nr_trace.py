"""Phase timestamps of one CTA of nr_gemm_kernel (library built with -DHN_NR_TRACE=<grid size of the launch to trace>)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
hn = importlib.import_module("nerf-3dtalker-code_b200")
net = hn.NeuralRenderer(featmap_size=32, img_size=512).cuda()
x = torch.randn(2, 256, 32, 32, device="cuda", requires_grad=True)
for _ in range(3):
    img = net(x)
    img.square().mean().backward()
torch.cuda.synchronize()
st = hn.ops._NR_STATUS[x.device].cpu()
t = st[16:24].view(torch.int64)
print("nkb", int(st[32]), "n_tile", int(st[33]), "problem", int(st[34]))
print("k loop %.1f us, wait done %.1f us, epilogue %.1f us" % ((t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[3] - t[2]) / 1e3))
