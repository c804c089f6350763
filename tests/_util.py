"""Shared helpers for the test-suite: golden fixture loading and comparison metrics."""
import os

import numpy as np
import torch

from oracle import headnerf_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["fs8_test_init", "fs8_train_trained", "fs16_test_trained"]
LEAVES = ["shape_code", "appea_code", "audiostyle", "batch_Rmats", "batch_Tvecs"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    fs, S, B, train, seed = [int(v) for v in z["meta"]]
    opt = O.OracleOptions(featmap_size=fs, pred_img_size=S)
    inp = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")}
    out = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out_")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad_")}
    pnorm = {k[len("pgrad_norm_"):]: float(z[k]) for k in z.files if k.startswith("pgrad_norm_")}
    pprobe = {k[len("pgrad_probe_"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("pgrad_probe_")}
    return {"opt": opt, "B": B, "mode": "train" if train else "test", "variant": str(z["variant"]),
            "inp": inp, "out": out, "grads": grads, "pnorm": pnorm, "pprobe": pprobe}


def probe_index(n, k=32):
    return torch.from_numpy((np.arange(k, dtype=np.int64) * 2654435761 % max(n, 1)).astype(np.int64))


def golden_loss(img):
    tgt = torch.linspace(0, 1, img.numel(), dtype=img.dtype, device=img.device).view_as(img)
    return ((img - tgt) ** 2).mean()


def cosine(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    na, nb = a.norm(), b.norm()
    if na == 0 and nb == 0:
        return 1.0
    return float((a @ b) / (na * nb + 1e-300))


def psnr(a, b):
    mse = float(((a.double().cpu() - b.double().cpu()) ** 2).mean())
    return 99.0 if mse == 0 else -10.0 * np.log10(mse)
