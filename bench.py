#!/usr/bin/env python
"""bench.py — HeadNeRF rendering hot path, ray·samples/s forward+backward at Reso64 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one forward+backward pass of the hot path (ray sampling -> positional encoding -> fg_CD_predictor ->
alpha compositing, and back: compositing bwd -> MLP data gradients -> MLP weight gradients + latent-code
gradients) over one batch of synthetic input: Reso64 (64x64 rays x 64 samples), batch 2 per GPU, mode "train"
(stratified jitter), random-init weights of the reference architecture, a fixed synthetic upstream gradient
(dL/dF, dL/dbg_alpha) standing in for the NeuralRenderer + loss that sit above the path.
  value : device-resident inputs, CUDA-event timed, max over ranks.
  e2e   : the same step with every input (pixels, codes, camera, upstream gradient) copied from pinned host
          memory and the results (feature map, bg_alpha, code gradients) read back, inside the timed region.
  --impl reference : the reference algorithm (oracle port of NetWorks/{utils,models}.py, plain PyTorch fp32) on
          the box's host cores, on a bounded sample of the same workload (rank 0 only).
Prints ONE JSON line (rank 0)."""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic cost per ray·sample (SURVEY.md §8d / DESIGN.md §4): latent columns folded into biases, padding not counted
MAC_FWD = 1_351_296
FLOP_FWD = 2 * MAC_FWD
FLOP_DGRAD = 2 * (MAC_FWD - 2 * 63 * 384)        # no gradient w.r.t. the positional encoding in the training step
FLOP_WGRAD = 2 * MAC_FWD
FLOP_STEP = FLOP_FWD + FLOP_DGRAD + FLOP_WGRAD
NS, FS, S_IMG, B_PER_GPU, C_FEAT = 64, 64, 512, 2, 256


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]           # upper half = samples taken under load
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(O, opt, B, seed, device):
    inp = O.synthetic_inputs(opt, B, seed=seed)
    g = torch.Generator().manual_seed(1000 + seed)
    n_r = opt.featmap_size ** 2
    inp["gF"] = torch.randn(B * n_r, C_FEAT, generator=g) * 1e-3
    inp["g_bg"] = torch.randn(B * n_r, generator=g) * 1e-3
    return inp


def run_ours(args):
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / debug log must not share stdout with the ONE JSON line
    hn = importlib.import_module("nerf-3dtalker-code_b200")
    from oracle import headnerf_oracle as O          # input factory only (synthetic_inputs); never on the timed path
    dist_mod = hn.dist
    rank, local, world = dist_mod.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    hn._lib.load()
    opt_o = O.OracleOptions(featmap_size=FS, pred_img_size=S_IMG)
    torch.manual_seed(0)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": FS, "featmap_nc": C_FEAT, "pred_img_size": S_IMG}), False, False).to(dev)
    for p in net.neural_render.parameters():
        p.requires_grad_(False)                       # the consumer is outside the timed hot path
    bucket = dist_mod.GradBucket(net.fg_CD_predictor.parameters())
    net.fuse_grad_accumulation()                      # kernels accumulate straight into the bucket's views (no per-parameter add kernels)
    host = build_inputs(O, opt_o, B_PER_GPU, seed=rank, device=dev)
    pinned = {k: v.contiguous().pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    code_keys = ["shape_code", "appea_code", "audiostyle"]
    M = B_PER_GPU * FS * FS * NS
    timer = hn.ops.TIMER

    def step(x):
        bucket.zero()
        codes = {k: x[k].detach().requires_grad_(True) for k in code_keys}
        Fm, bg = net.render_rays("train", x["batch_xy"], codes["audiostyle"], codes["shape_code"], codes["appea_code"],
                                 x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        torch.autograd.backward([Fm.reshape(-1, C_FEAT), bg.reshape(-1)], [x["gF"], x["g_bg"]])
        bucket.all_reduce()
        return Fm, bg, codes

    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    late_keys = ("gF", "g_bg")                        # needed only by the backward pass
    pending = []                                      # (event, host results) of steps whose read-back may still be in flight
    rays_ = B_PER_GPU * FS * FS
    host_out = [{"F": torch.empty(B_PER_GPU, FS * FS, C_FEAT).pin_memory(), "bg": torch.empty(B_PER_GPU, FS * FS).pin_memory(),
                 **{k: torch.empty_like(host[k]).pin_memory() for k in code_keys}} for _ in range(2)]   # pinned result buffers, double-buffered
    e2e_count = [0]

    def step_e2e():
        """The same step from HOST buffers: every input is copied from pinned memory and every result read back inside the timed
        region.  The upstream gradient (8 MB) travels on a copy stream while the forward runs, and the feature map (8 MB) is read
        back on another while the backward runs; the host waits for step i-1's results while step i is already queued (double
        buffering) - nothing is skipped, copies and launches just overlap the kernels.  timed() drains the last step."""
        main = torch.cuda.current_stream()
        x = {k: pinned[k].to(dev, non_blocking=True) for k in pinned if k not in late_keys}
        copy_in.wait_stream(main)
        with torch.cuda.stream(copy_in):
            late = {k: pinned[k].to(dev, non_blocking=True) for k in late_keys}
        bucket.zero()
        codes = {k: x[k].detach().requires_grad_(True) for k in code_keys}
        Fm, bg = net.render_rays("train", x["batch_xy"], codes["audiostyle"], codes["shape_code"], codes["appea_code"],
                                 x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        ho = host_out[e2e_count[0] & 1]                # last used two steps ago: that step's event has been waited for
        e2e_count[0] += 1
        copy_out.wait_stream(main)
        with torch.cuda.stream(copy_out):
            ho["F"].copy_(Fm.detach(), non_blocking=True)
            ho["bg"].copy_(bg.detach(), non_blocking=True)
            Fm.record_stream(copy_out); bg.record_stream(copy_out)
        outs = ho
        main.wait_stream(copy_in)
        for t in late.values():
            t.record_stream(main)
        torch.autograd.backward([Fm.reshape(-1, C_FEAT), bg.reshape(-1)], [late["gF"], late["g_bg"]])
        bucket.all_reduce()
        for k in code_keys:
            ho[k].copy_(codes[k].grad, non_blocking=True)
        main.wait_stream(copy_out)
        ev = torch.cuda.Event()
        ev.record(main)
        pending.append((ev, outs))
        if len(pending) > 1:
            done_ev, done = pending.pop(0)
            done_ev.synchronize()                     # the previous step's feature map, bg_alpha and code gradients are on the host
            return done
        return None

    def timed(fn, k):
        dist_mod.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        dist_mod.barrier()
        return dist_mod.max_over_ranks(e0.elapsed_time(e1), dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                               # nvidia-smi needs ~0.2 s to emit its first sample
    for _ in range(max(args.warmup, 3)):
        step(resident)
    torch.cuda.synchronize()
    hn.ops.check_status(net.last_meta["last_status"], "warm-up")
    timer.reset(); timer.enabled = True
    ms_total = timed(lambda: step(resident), args.steps)
    timer.enabled = False
    kern = timer.summary()
    launches = timer.launches
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(4):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    hn.ops.check_status(net.last_meta["last_status"], "timed region")

    # the same step in the high-precision mode (split-operand GEMMs, fp32 activations): reported beside the headline, N = 1 only
    high = None
    if world == 1 and not args.no_high:
        net.precision = "high"
        for _ in range(2):
            step(resident)
        torch.cuda.synchronize()
        timer.reset(); timer.enabled = True
        n_high = max(2, min(args.steps, 10))
        ms_high = timed(lambda: step(resident), n_high)
        timer.enabled = False
        hn.ops.check_status(net.last_meta["last_status"], "high-precision timed region")
        high = {"ms_per_step": round(ms_high / n_high, 4), "value": round(M / (ms_high / n_high * 1e-3), 1), "unit": "ray*samples/s", "steps": n_high,
                "gpu_launches_per_step": timer.launches / n_high,
                "calls_ms": {k: round(v["ms_avg"], 4) for k, v in timer.summary().items()},
                "note": "precision='high': hi+lo split operands, 3 tcgen05 products per GEMM, fp32 activations in HBM (csrc/hn_precise.cu)"}
        net.precision = "fast"

    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()
    if rank != 0:
        return
    peaks = load_peaks()
    ms_step = ms_total / args.steps
    value = world * M / (ms_step * 1e-3)
    e2e_val = world * M / (ms_e2e / args.steps * 1e-3)
    flops = {"hn_mlp_fwd": FLOP_FWD * M, "hn_mlp_bwd_data": FLOP_DGRAD * M, "hn_mlp_bwd_weights": FLOP_WGRAD * M}
    kernels = {}
    for name, d in kern.items():
        k = {"launches_per_step": d["n"] / args.steps, "ms_avg": round(d["ms_avg"], 4)}
        if name in flops:
            k["tflops_algorithmic"] = round(flops[name] / (d["ms_avg"] * 1e-3) / 1e12, 1)
            k["frac_of_tensor_peak"] = round(k["tflops_algorithmic"] / peaks["tflops"], 4)
        kernels[name] = k
    rays = B_PER_GPU * FS * FS
    comp_bytes = {"hn_composite_fwd": M * (C_FEAT + 2) * 4 + rays * (C_FEAT + 1) * 4,
                  "hn_composite_bwd": M * (C_FEAT + 2) * 4 + rays * (C_FEAT + 1) * 4 + M * (C_FEAT * 2 + 4)}
    for name, nbytes in comp_bytes.items():
        if name in kernels:
            gbs = nbytes / (kernels[name]["ms_avg"] * 1e-3) / 1e9
            kernels[name].update({"gbs_algorithmic": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 4)})
    # DRAM traffic per launch comes from the committed ncu --set full capture of this workload (profiles/)
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r01h_traffic.json")) as f:
            traffic = {k: v["dram_gbytes_per_launch"] * 1e9 for k, v in json.load(f)["kernels"].items()}
    except Exception:
        pass
    # weight gradients: every saved activation / gradient operand block is read once (58 + 58 + 4 blocks of 16 KiB per 128-sample tile)
    WGRAD_BYTES_PER_SAMPLE = 120 * 16384 // 128
    if "hn_mlp_bwd_weights" in kernels:
        gbs = WGRAD_BYTES_PER_SAMPLE * M / (kernels["hn_mlp_bwd_weights"]["ms_avg"] * 1e-3) / 1e9
        kernels["hn_mlp_bwd_weights"].update({"gbs_algorithmic": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 4)})
    dom = max((n for n in flops if n in kernels), key=lambda n: kernels[n]["ms_avg"])
    # which roof bounds the dominant kernel: arithmetic intensity (algorithmic FLOP per algorithmic HBM byte) against the ridge point
    ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    dom_bytes = {"hn_mlp_bwd_weights": WGRAD_BYTES_PER_SAMPLE * M}.get(dom)
    if dom_bytes is not None and flops[dom] / dom_bytes < ridge:
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["gbs_algorithmic"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": kernels[dom]["frac_of_hbm_peak"], "traffic": traffic.get(dom),
                    "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r01h_traffic.json)", "peak_source": peaks["src"],
                    "arithmetic_intensity_flop_per_byte": round(flops[dom] / dom_bytes, 1), "ridge_flop_per_byte": round(ridge, 1),
                    "algorithmic_bytes_per_launch": dom_bytes, "tensor_frac_same_kernel": kernels[dom]["frac_of_tensor_peak"],
                    "step_tflops_algorithmic": round(FLOP_STEP * M / (ms_step * 1e-3) / 1e12, 1)}
    else:
        roofline = {"kernel": dom, "bound": "tensor", "achieved": kernels[dom]["tflops_algorithmic"], "peak": peaks["tflops"],
                    "unit": "TFLOP/s", "frac": kernels[dom]["frac_of_tensor_peak"], "traffic": traffic.get(dom),
                    "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r01h_traffic.json)", "peak_source": peaks["src"],
                    "step_tflops_algorithmic": round(FLOP_STEP * M / (ms_step * 1e-3) / 1e12, 1)}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = (rays * C_FEAT + rays) * 4 + sum(host[k].numel() for k in code_keys) * 4
    out = {
        "metric": "ray_samples_per_sec_fwd_bwd_reso64", "value": round(value, 1), "unit": "ray*samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
        "config": {"workload": "HeadNeRF Reso64 hot path fwd+bwd, batch 2 per GPU, 64x64 rays x 64 samples, mode train, random-init weights",
                   "rays_per_gpu": rays, "samples_per_ray": NS, "ray_samples_per_step_per_gpu": M, "parallelism": f"dp{world} (batch-sharded, flat-bucket NCCL all-reduce)",
                   "l2": "per-step working set (saved activations 4.0 GB + gradients 4.1 GB) far exceeds the 126 MB L2; no explicit flush"},
        "e2e": {"value": round(e2e_val, 1), "unit": "ray*samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 4)},
        "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
        "roofline": roofline, "kernels": kernels, "clocks": clocks,
    }
    if high is not None:
        high["tflops_algorithmic_step"] = round(FLOP_STEP * M / (high["ms_per_step"] * 1e-3) / 1e12, 1)
        out["high_precision"] = high
    if world == 1:
        out["cpu_baseline"] = cpu_baseline(sample_rays=1024, repeats=2)
    print(json.dumps(out), file=RESULT_OUT, flush=True)


def _reference_pass(O, opt, sd, inp, n_rays_sample):
    """One fwd+bwd of the reference algorithm (oracle port) on the first `n_rays_sample` rays of every item."""
    x = {k: (v[..., :n_rays_sample].contiguous() if k == "batch_xy" else v) for k, v in inp.items()}
    B = x["batch_xy"].shape[0]
    codes = {k: x[k].clone().requires_grad_(True) for k in ("shape_code", "appea_code", "audiostyle")}
    r = O.render_features(sd, opt, "train", x["batch_xy"], codes["audiostyle"], codes["shape_code"], codes["appea_code"],
                          x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
    gF = x["gF"].view(B, -1, C_FEAT)[:, :n_rays_sample].permute(0, 2, 1)
    gb = x["g_bg"].view(B, 1, -1)[:, :, :n_rays_sample]
    torch.autograd.backward([r["F"], r["bg_alpha"]], [gF, gb])
    return B * n_rays_sample * opt.num_sample_coarse


def _reference_setup():
    from oracle import headnerf_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    opt = O.OracleOptions(featmap_size=FS, pred_img_size=S_IMG)
    torch.manual_seed(0)
    sd = {k: v.requires_grad_(k.startswith("fg_CD_predictor")) for k, v in O.formula_state_dict(opt, "init").items()}
    inp = build_inputs(O, opt, B_PER_GPU, seed=0, device="cpu")
    return O, opt, sd, inp


def cpu_baseline(sample_rays, repeats):
    O, opt, sd, inp = _reference_setup()
    _reference_pass(O, opt, sd, inp, 16)
    best = 1e30
    n = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        n = _reference_pass(O, opt, sd, inp, sample_rays)
        best = min(best, time.perf_counter() - t0)
    return {"value": round(n / best, 1), "unit": "ray*samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first {sample_rays} rays x {NS} samples of each of {B_PER_GPU} items ({n} ray*samples), fwd+bwd, best of {repeats}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O, opt, sd, inp = _reference_setup()
    t0 = time.perf_counter()
    n0 = _reference_pass(O, opt, sd, inp, 32)
    rate = n0 / (time.perf_counter() - t0)
    total_steps = args.steps + args.warmup
    budget_s = float(os.environ.get("HN_BENCH_REF_BUDGET_S", "150"))      # host seconds for all steps together (tests shrink it)
    rays = int(max(32, min(FS * FS, rate * budget_s / total_steps / (B_PER_GPU * NS))))
    rays -= rays % 2
    for _ in range(args.warmup):
        _reference_pass(O, opt, sd, inp, rays)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        n += _reference_pass(O, opt, sd, inp, rays)
    dt = time.perf_counter() - t0
    value = n / dt
    sample = f"first {rays} rays x {NS} samples of each of {B_PER_GPU} items per step (bounded sample of the Reso64 batch-2 workload)"
    out = {"impl": "reference", "metric": "ray_samples_per_sec_fwd_bwd_reso64", "value": round(value, 1), "unit": "ray*samples/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "HeadNeRF Reso64 hot path fwd+bwd, batch 2, mode train, reference algorithm on host CPU", "sample": sample},
           "cpu_baseline": {"value": round(value, 1), "unit": "ray*samples/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
           "e2e": {"value": round(value, 1), "unit": "ray*samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), file=RESULT_OUT, flush=True)


RESULT_OUT = sys.stdout


def main():
    # stdout carries exactly ONE line (the JSON result): file descriptor 1 is pointed at stderr for the whole run, so that
    # library chatter (NCCL's version banner, torchrun notices, ...) cannot end up in front of it
    global RESULT_OUT
    sys.stdout.flush()
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-high", action="store_true", help="skip the auxiliary high-precision-mode measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
