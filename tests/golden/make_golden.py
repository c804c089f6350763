"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only:  python tests/golden/make_golden.py
Each fixture holds the seeded synthetic inputs (oracle.synthetic_inputs), the explicit jitter draw, and
the reference's outputs: composited feature map F / bg_alpha / depth, the two images, and the gradients of
a fixed quadratic loss on every input leaf plus per-parameter gradient probes (norm + 32 fixed entries).
Weights are NOT stored: they come from oracle.formula_state_dict (RNG-free), loaded into the reference net
with load_state_dict(strict=True), i.e. through the reference's own state-dict layout."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import headnerf_oracle as O   # noqa: E402
from oracle import ref_import             # noqa: E402

CASES = {
    # name: (featmap_size, pred_img_size, B, mode, weight variant, seed)
    "fs8_test_init": (8, 32, 2, "test", "init", 1),
    "fs8_train_trained": (8, 64, 1, "train", "trained", 2),
    "fs16_test_trained": (16, 64, 1, "test", "trained", 3),
}
LEAVES = ["shape_code", "appea_code", "audiostyle", "batch_Rmats", "batch_Tvecs"]


def probe_index(n, k=32):
    return (np.arange(k, dtype=np.int64) * 2654435761 % max(n, 1)).astype(np.int64)


def loss_fn(img):
    tgt = torch.linspace(0, 1, img.numel(), dtype=img.dtype, device=img.device).view_as(img)
    return ((img - tgt) ** 2).mean()


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (fs, S, B, mode, variant, seed) in CASES.items():
        opt = O.OracleOptions(featmap_size=fs, pred_img_size=S)
        _, net = ref_import.build(fs, S)
        net.load_state_dict(O.formula_state_dict(opt, variant), strict=True)
        net.eval()
        inp = O.synthetic_inputs(opt, B, seed=seed, jitter=(mode == "train"))
        x = {k: v.clone().requires_grad_(k in LEAVES) for k, v in inp.items()}
        taps = {}
        net.calc_color_func.register_forward_hook(
            lambda m, i, o: taps.update(F=o[0].detach(), bg_alpha=o[1].detach(), depth=o[2].detach()))
        # make the reference consume OUR jitter draw: rand_like at utils.py:77 is the only RNG use
        orig = torch.rand_like
        if mode == "train":
            torch.rand_like = lambda t, *a, **k: x["t_rand"].to(t)
        try:
            res = net(mode, x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"],
                      x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        finally:
            torch.rand_like = orig
        img = res["coarse_dict"]["merge_img"]
        loss_fn(img).backward()
        rec = {f"in_{k}": v.numpy() for k, v in inp.items()}
        rec.update({f"out_{k}": v.numpy() for k, v in taps.items()})
        rec["out_merge_img"] = img.detach().numpy()
        rec["out_bg_img"] = res["coarse_dict"]["bg_img"].detach().numpy()
        for k in LEAVES:
            rec[f"grad_{k}"] = x[k].grad.numpy()
        for k, p in net.named_parameters():
            g = p.grad.reshape(-1).numpy()
            rec[f"pgrad_norm_{k}"] = np.array(np.linalg.norm(g.astype(np.float64)))
            rec[f"pgrad_probe_{k}"] = g[probe_index(g.size)]
        rec["meta"] = np.array([fs, S, B, 1 if mode == "train" else 0, seed])
        rec["variant"] = np.array(variant)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **rec)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB",
              "F absmax %.3f bg_alpha range [%.3f, %.3f]" % (taps["F"].abs().max(), taps["bg_alpha"].min(), taps["bg_alpha"].max()))


if __name__ == "__main__":
    main()
