"""Runs the one-call NeuralRenderer (Reso32HR geometry: 32x32x256 -> 512x512, batch 2) a few times; under ncu this gives the
per-kernel durations of hn_nr_fwd / hn_nr_bwd.   usage: python tools/nr_probe.py [steps]"""
import importlib
import sys
import time

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

hn = importlib.import_module("nerf-3dtalker-code_b200")
hn.build_library()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(0)
net = hn.NeuralRenderer(featmap_size=32, img_size=512).cuda()
x = torch.randn(2, 256, 32, 32, device="cuda", requires_grad=True)
tgt = torch.rand(2, 3, 512, 512, device="cuda")
t0 = time.time()
for it in range(steps + 2):
    if it == 2:
        torch.cuda.synchronize(); t0 = time.time()
    net.zero_grad(set_to_none=True)
    ((net(x) - tgt) ** 2).mean().backward()
torch.cuda.synchronize()
print("ms per fwd+bwd:", (time.time() - t0) / max(steps, 1) * 1e3)
