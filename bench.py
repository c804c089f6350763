#!/usr/bin/env python
"""bench.py — HeadNeRF rendering hot path, ray·samples/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--config 2 (default, the configuration BASELINE.json's metric is quoted on): a step = one forward+backward pass of the hot path
(ray sampling -> positional encoding -> fg_CD_predictor -> alpha compositing, and back: compositing bwd -> MLP data gradients
-> MLP weight gradients + latent-code gradients) over one batch of synthetic input: Reso64 (64x64 rays x 64 samples), batch 2
per GPU, mode "train" (stratified jitter), random-init weights of the reference architecture, a fixed synthetic upstream
gradient (dL/dF, dL/dbg_alpha) standing in for the NeuralRenderer + loss above the path, one flat-bucket NCCL all-reduce.
  value     : device-resident inputs, CUDA-event timed, max over ranks.
  e2e       : the same step with every input copied from pinned host memory and the results read back, inside the timed region.
  sustained : the same step repeated for >= --sustain-s seconds (the K-step region is ~0.1 s: too short for the power cap).
  roofline  : the dominant MLP kernel against the TENSOR roof (algorithmic FLOP / measured duration /
              measured sustained bf16 peak); its HBM view (algorithmic operand bytes, ncu DRAM traffic) is a secondary key.
--config 3 : HeadNeRF Reso32HR full training step (HeadNeRFNet.forward mode train incl. the NeuralRenderer consumer, the
             reference's photometric loss, backward, NCCL all-reduce, Adam) - batch 2 per GPU (`--shard items`, weak scaling) or
             ONE item of batch 2 ray-sharded over the GPUs (`--shard rays`, all-gather / reduce-scatter of the feature slices).
--config 4 : the FittingSingleImage_new latent-code optimisation loop: --steps iterations (default 500) on one GPU.
--config 5 : forward-only sweep, 16K-4M rays x 32/64/128 samples, rays sharded over the GPUs.
--impl reference : the reference algorithm (oracle port of NetWorks/{utils,models}.py, plain PyTorch fp32) on the box's host
             cores, all threads, on a bounded sample of the same workload in chunks of 1024 rays per item (rank 0 only) - the
             same chunked routine the `cpu_baseline` key of our arm times.
Prints ONE JSON line (rank 0)."""
import argparse
import importlib
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic cost per ray·sample (SURVEY.md §8d / DESIGN.md §4): latent columns folded into biases, padding not counted.  The
# kernels EXECUTE less: RGB_layer_0 (no activation) is multiplied into RGB_layer_1 once per weight version (SURVEY.md appendix A4,
# "optional algebraic fusion"), which removes 384 x 384 MACs per ray·sample from each pass; both figures are reported.
MAC_R0 = 384 * 384
MAC_FWD = 1_351_296
FLOP_FWD = 2 * MAC_FWD
FLOP_DGRAD = 2 * (MAC_FWD - 2 * 63 * 384)        # no gradient w.r.t. the positional encoding in the training step
FLOP_WGRAD = 2 * MAC_FWD
FLOP_STEP = FLOP_FWD + FLOP_DGRAD + FLOP_WGRAD
NS, FS, S_IMG, B_PER_GPU, C_FEAT = 64, 64, 512, 2, 256
CPU_CHUNK_RAYS = 1024                             # rays per item per chunk of the CPU arm (both cpu_baseline and --impl reference)
CODE_KEYS = ("shape_code", "appea_code", "audiostyle")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "tflops_burst": p["bf16_tflops"],
                "src": "measured (MEASURED_PEAKS.json: bf16 sustained TFLOP/s, copy GB/s)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1700.0, "src": "fallback (B200_PROFILING.md)"}


def latest_traffic():
    """DRAM bytes per launch from the newest committed ncu --set full capture of this workload (profiles/rNN*_traffic.json)."""
    d = os.path.join(ROOT, "profiles")
    try:
        names = sorted(n for n in os.listdir(d) if n.endswith("_traffic.json"))
        with open(os.path.join(d, names[-1])) as f:
            return names[-1], {k: v["dram_gbytes_per_launch"] * 1e9 for k, v in json.load(f)["kernels"].items()}
    except Exception:
        return None, {}


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


def build_inputs(O, opt, B, seed, jitter=False):
    inp = O.synthetic_inputs(opt, B, seed=seed, jitter=jitter)
    g = torch.Generator().manual_seed(1000 + seed)
    n_r = opt.featmap_size ** 2
    inp["gF"] = torch.randn(B * n_r, C_FEAT, generator=g) * 1e-3
    inp["g_bg"] = torch.randn(B * n_r, generator=g) * 1e-3
    return inp


def workload_config(cfg, world, shard="items"):
    """The `config` object of the JSON line: identical in both arms (--impl ours / reference), so the driver's same-config check
    compares like with like; what each arm actually touched per step is stated in its own keys (e2e, cpu_baseline.sample)."""
    if cfg == 2:
        return {"workload": "HeadNeRF Reso64 hot path fwd+bwd, batch 2 per GPU, 64x64 rays x 64 samples, mode train, random-init weights",
                "baseline_config": 2, "rays_per_gpu": B_PER_GPU * FS * FS, "samples_per_ray": NS,
                "ray_samples_per_step_per_gpu": B_PER_GPU * FS * FS * NS,
                "parallelism": f"dp{world} (batch-sharded, flat-bucket NCCL all-reduce)",
                "l2": "per-step working set (saved activations 4.0 GB + gradients 4.1 GB) far exceeds the 126 MB L2; no explicit flush"}
    if cfg == 3:
        per_gpu = 2 * 1024 * 64 if shard == "items" else 2 * 1024 * 64 // world
        return {"workload": "HeadNeRF Reso32HR full training step (32x32 rays x 64 samples -> 512x512 image): HeadNeRFNet.forward mode train "
                            "incl. NeuralRenderer, photometric loss, backward, NCCL all-reduce, Adam; random-init weights",
                "baseline_config": 3, "samples_per_ray": 64, "ray_samples_per_step_per_gpu": per_gpu,
                "parallelism": (f"dp{world}: batch 2 per GPU (weak), bucket all-reduce with the consumer's range overlapped" if shard == "items" else
                                f"rays{world}: ONE batch of 2 items, each item's rays sharded over {world} GPUs (strong): all-gather of [B,N_r/G,257] "
                                "feature slices forward, reduce-scatter backward, replicated consumer, bucket all-reduce"),
                "l2": "working set per step (~0.5 GB of saved operands per item) exceeds the 126 MB L2; no explicit flush"}
    if cfg == 4:
        return {"workload": "FittingSingleImage_new latent-code optimisation loop: Reso32 (32x32 rays x 64 samples -> 256x256 image), batch 1, mode test, "
                            "frozen weights, gradients to shape/appearance code offsets + delta Euler / delta T, photometric loss, Adam (4 groups)",
                "baseline_config": 4, "samples_per_ray": 64, "ray_samples_per_step_per_gpu": 1024 * 64, "parallelism": "1 GPU (the loop is sequential)",
                "l2": "the per-iteration working set (~60 MB of fp32 activations) fits L2 by construction of the workload; iterations depend on each other"}
    return {"workload": "forward-only sweep (sampling + PE + MLP + compositing): 16K / 256K / 1M / 4M rays x 32 / 64 / 128 samples, mode test, random-init weights",
            "baseline_config": 5, "parallelism": f"rays sharded over {world} GPUs, no collective", "value_is": "4M rays x 64 samples",
            "l2": "every point streams far more than the 126 MB L2 (>= 0.5 GB of features); no explicit flush"}


def timed_region(dist_mod, dev, fn, k):
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


def sustained_run(dist_mod, dev, fn, ms_step, seconds, units_per_step):
    """The same step for >= `seconds` of device time (count fixed up front from the measured step so that all ranks agree)."""
    n = max(8, int(math.ceil(seconds * 1e3 / max(ms_step, 1e-3))))
    n = int(dist_mod.max_over_ranks(float(n), dev))
    ms = timed_region(dist_mod, dev, fn, n)
    return {"steps": n, "seconds": round(ms * 1e-3, 3), "ms_per_step": round(ms / n, 4), "value": round(units_per_step / (ms / n * 1e-3), 1),
            "unit": "ray*samples/s"}


def setup(cfg_name):
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's banner / debug log must not share stdout with the ONE JSON line
    hn = importlib.import_module("nerf-3dtalker-code_b200")
    from oracle import headnerf_oracle as O          # input factory only (synthetic_inputs); never on the timed path
    rank, local, world = hn.dist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit(f"bench.py ({cfg_name}) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    hn._lib.load()
    return hn, O, rank, local, world, torch.device("cuda", local)


def finish(world):
    if world > 1:
        import torch.distributed as tdist
        tdist.barrier()
        tdist.destroy_process_group()


# =====================================================================================================================
# config 2 — the headline
# =====================================================================================================================
def run_config2(args):
    hn, O, rank, local, world, dev = setup("ours")
    dist_mod = hn.dist
    opt_o = O.OracleOptions(featmap_size=FS, pred_img_size=S_IMG)
    torch.manual_seed(0)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": FS, "featmap_nc": C_FEAT, "pred_img_size": S_IMG}), False, False).to(dev)
    for p in net.neural_render.parameters():
        p.requires_grad_(False)                       # the consumer is outside the timed hot path
    bucket = dist_mod.GradBucket(net.fg_CD_predictor.parameters())
    net.fuse_grad_accumulation()                      # kernels accumulate straight into the bucket's views (no per-parameter add kernels)
    host = build_inputs(O, opt_o, B_PER_GPU, seed=rank)
    pinned = {k: v.contiguous().pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    M = B_PER_GPU * FS * FS * NS
    timer = hn.ops.TIMER

    def step(x):
        bucket.zero()
        codes = {k: x[k].detach().requires_grad_(True) for k in CODE_KEYS}
        Fm, bg = net.render_rays("train", x["batch_xy"], codes["audiostyle"], codes["shape_code"], codes["appea_code"],
                                 x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        torch.autograd.backward([Fm.reshape(-1, C_FEAT), bg.reshape(-1)], [x["gF"], x["g_bg"]])
        bucket.all_reduce()
        return Fm, bg, codes

    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    late_keys = ("gF", "g_bg")                        # needed only by the backward pass
    pending = []                                      # (event, host results) of steps whose read-back may still be in flight
    host_out = [{"F": torch.empty(B_PER_GPU, FS * FS, C_FEAT).pin_memory(), "bg": torch.empty(B_PER_GPU, FS * FS).pin_memory(),
                 **{k: torch.empty_like(host[k]).pin_memory() for k in CODE_KEYS}} for _ in range(2)]   # pinned result buffers, double-buffered
    e2e_count = [0]

    def step_e2e():
        """The same step from HOST buffers: every input is copied from pinned memory and every result read back inside the timed
        region.  The upstream gradient (8 MB) travels on a copy stream while the forward runs, and the feature map (8 MB) is read
        back on another while the backward runs; the host waits for step i-1's results while step i is already queued (double
        buffering) - nothing is skipped, copies and launches just overlap the kernels.  timed_region() drains the last step."""
        main = torch.cuda.current_stream()
        x = {k: pinned[k].to(dev, non_blocking=True) for k in pinned if k not in late_keys}
        copy_in.wait_stream(main)
        with torch.cuda.stream(copy_in):
            late = {k: pinned[k].to(dev, non_blocking=True) for k in late_keys}
        bucket.zero()
        codes = {k: x[k].detach().requires_grad_(True) for k in CODE_KEYS}
        Fm, bg = net.render_rays("train", x["batch_xy"], codes["audiostyle"], codes["shape_code"], codes["appea_code"],
                                 x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        ho = host_out[e2e_count[0] & 1]                # last used two steps ago: that step's event has been waited for
        e2e_count[0] += 1
        copy_out.wait_stream(main)
        with torch.cuda.stream(copy_out):
            ho["F"].copy_(Fm.detach(), non_blocking=True)
            ho["bg"].copy_(bg.detach(), non_blocking=True)
            Fm.record_stream(copy_out); bg.record_stream(copy_out)
        main.wait_stream(copy_in)
        for t in late.values():
            t.record_stream(main)
        torch.autograd.backward([Fm.reshape(-1, C_FEAT), bg.reshape(-1)], [late["gF"], late["g_bg"]])
        bucket.all_reduce()
        for k in CODE_KEYS:
            ho[k].copy_(codes[k].grad, non_blocking=True)
        main.wait_stream(copy_out)
        ev = torch.cuda.Event()
        ev.record(main)
        pending.append((ev, ho))
        if len(pending) > 1:
            done_ev, done = pending.pop(0)
            done_ev.synchronize()                     # the previous step's feature map, bg_alpha and code gradients are on the host
            return done
        return None

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                               # nvidia-smi needs ~0.2 s to emit its first sample
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(resident)
    torch.cuda.synchronize()
    net.check_faults()
    assert net.last_meta["precision"] == "fast", "random-init weights: precision='auto' must select the single-pass kernels"
    timer.reset(); timer.enabled = True
    ms_total = timed_region(dist_mod, dev, lambda: step(resident), args.steps)
    timer.enabled = False
    kern = timer.summary()
    launches = timer.launches
    ms_step = ms_total / args.steps
    sustained = sustained_run(dist_mod, dev, lambda: step(resident), ms_step, args.sustain_s, world * M)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(4):
        step_e2e()
    ms_e2e = timed_region(dist_mod, dev, step_e2e, args.steps)
    net.check_faults()

    # the same step in the high-precision mode (split-operand GEMMs, fp32 activations): reported beside the headline, N = 1 only
    high = None
    if world == 1 and not args.no_high:
        net.precision = "high"
        for _ in range(2):
            step(resident)
        torch.cuda.synchronize()
        timer.reset(); timer.enabled = True
        n_high = max(2, min(args.steps, 10))
        ms_high = timed_region(dist_mod, dev, lambda: step(resident), n_high)
        timer.enabled = False
        net.check_faults()
        high = {"ms_per_step": round(ms_high / n_high, 4), "value": round(M / (ms_high / n_high * 1e-3), 1), "unit": "ray*samples/s", "steps": n_high,
                "gpu_launches_per_step": timer.launches / n_high,
                "calls_ms": {k: round(v["ms_avg"], 4) for k, v in timer.summary().items()},
                "note": "precision='high': hi+lo split operands, 3 tcgen05 products per GEMM, fp32 activations in HBM (csrc/hn_precise.cu); "
                        "what precision='auto' selects for checkpoints whose feature scale puts the single-pass kernels outside the 1e-3 gate"}
        net.precision = "auto"

    finish(world)
    if rank != 0:
        return
    peaks = load_peaks()
    value = world * M / (ms_step * 1e-3)
    e2e_val = world * M / (ms_e2e / args.steps * 1e-3)
    flops = {"hn_mlp_fwd": FLOP_FWD * M, "hn_mlp_bwd_data": FLOP_DGRAD * M, "hn_mlp_bwd_weights": FLOP_WGRAD * M}
    tfile, traffic = latest_traffic()
    kernels = {}
    for name, d in kern.items():
        k = {"launches_per_step": d["n"] / args.steps, "ms_avg": round(d["ms_avg"], 4)}
        if name in flops:
            k["tflops_algorithmic"] = round(flops[name] / (d["ms_avg"] * 1e-3) / 1e12, 1)
            k["frac_of_tensor_peak"] = round(k["tflops_algorithmic"] / peaks["tflops"], 4)
            k["tflops_executed"] = round((flops[name] - 2 * MAC_R0 * M) / (d["ms_avg"] * 1e-3) / 1e12, 1)
            k["executed_frac_of_tensor_peak"] = round(k["tflops_executed"] / peaks["tflops"], 4)
        if name in traffic:
            k["dram_bytes_per_launch_ncu"] = traffic[name]
        kernels[name] = k
    rays = B_PER_GPU * FS * FS
    comp_bytes = {"hn_composite_fwd": M * (C_FEAT + 2) * 4 + rays * (C_FEAT + 1) * 4,
                  "hn_composite_bwd": M * (C_FEAT + 2) * 4 + rays * (C_FEAT + 1) * 4 + M * (C_FEAT * 2 + 4)}
    for name, nbytes in comp_bytes.items():
        if name in kernels:
            gbs = nbytes / (kernels[name]["ms_avg"] * 1e-3) / 1e9
            kernels[name].update({"algorithmic_bytes_per_launch": nbytes, "gbs_algorithmic": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 4)})
    # operand bytes the weight-gradient pass must read once (58 activation + 58 gradient + 4 dL/dfeat blocks of 16 KiB per 128-sample
    # tile): a property of this design's saved-operand layout, reported as the secondary (HBM) view, never as the roofline fraction
    WGRAD_OPERAND_BYTES = 120 * 16384 // 128 * M
    if "hn_mlp_bwd_weights" in kernels:
        gbs = WGRAD_OPERAND_BYTES / (kernels["hn_mlp_bwd_weights"]["ms_avg"] * 1e-3) / 1e9
        kernels["hn_mlp_bwd_weights"].update({"operand_bytes_per_launch": WGRAD_OPERAND_BYTES, "gbs_operands": round(gbs, 1),
                                              "operand_frac_of_hbm_peak": round(gbs / peaks["hbm_gbs"], 4)})
    dom = max((n for n in flops if n in kernels), key=lambda n: kernels[n]["ms_avg"])
    step_tflops = FLOP_STEP * M / (ms_step * 1e-3) / 1e12
    roofline = {"kernel": dom, "bound": "tensor", "achieved": kernels[dom]["tflops_algorithmic"], "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": kernels[dom]["frac_of_tensor_peak"], "traffic": traffic.get(dom),
                "traffic_unit": f"bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/{tfile})",
                "peak_source": peaks["src"], "algorithmic_flop_per_launch": flops[dom],
                "algorithmic_flop_per_ray_sample": {"hn_mlp_fwd": FLOP_FWD, "hn_mlp_bwd_data": FLOP_DGRAD, "hn_mlp_bwd_weights": FLOP_WGRAD}[dom],
                "all_mlp_kernels_frac": {n: kernels[n]["frac_of_tensor_peak"] for n in flops if n in kernels},
                "executed": {"note": "RGB_layer_0 multiplied into RGB_layer_1 (SURVEY.md appendix A4): 147 456 MAC per ray*sample fewer per pass than the "
                                     "section 8d figure `frac` is quoted on; the same kernels against the MACs they actually issue:",
                             "mac_per_ray_sample_fwd": MAC_FWD - MAC_R0, "frac": kernels[dom]["executed_frac_of_tensor_peak"],
                             "all_mlp_kernels_frac": {n: kernels[n]["executed_frac_of_tensor_peak"] for n in flops if n in kernels}},
                "step_tflops_algorithmic": round(step_tflops, 1), "step_frac_of_tensor_peak": round(step_tflops / peaks["tflops"], 4),
                "hbm_view": ({"operand_gbs": kernels[dom].get("gbs_operands"), "operand_frac_of_hbm_peak": kernels[dom].get("operand_frac_of_hbm_peak"),
                              "peak_gbs": peaks["hbm_gbs"]} if dom == "hn_mlp_bwd_weights" else None),
                "compositing": {n: {"gbs": kernels[n]["gbs_algorithmic"], "frac_of_hbm_peak": kernels[n]["frac_of_hbm_peak"]} for n in comp_bytes if n in kernels}}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values())
    d2h = (rays * C_FEAT + rays) * 4 + sum(host[k].numel() for k in CODE_KEYS) * 4
    out = {
        "metric": "ray_samples_per_sec_fwd_bwd_reso64", "value": round(value, 1), "unit": "ray*samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
        "config": workload_config(2, world),
        "e2e": {"value": round(e2e_val, 1), "unit": "ray*samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 4)},
        "sustained": sustained,
        "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
        "roofline": roofline, "kernels": kernels, "clocks": clocks, "precision": "auto -> fast (probe error %.1e <= %.0e)" % (net._calib[3], net.auto_tolerance),
    }
    if high is not None:
        high["tflops_algorithmic_step"] = round(FLOP_STEP * M / (high["ms_per_step"] * 1e-3) / 1e12, 1)
        out["high_precision"] = high
    if world == 1:
        out["cpu_baseline"] = cpu_arm(2, budget_s=15.0, repeats=1)["cpu_baseline"]
    print(json.dumps(out), file=RESULT_OUT, flush=True)


# =====================================================================================================================
# config 3 — Reso32HR full training step
# =====================================================================================================================
def run_config3(args):
    hn, O, rank, local, world, dev = setup("config 3")
    dist_mod = hn.dist
    fs, S, B = 32, 512, 2
    opt_o = O.OracleOptions(featmap_size=fs, pred_img_size=S)
    torch.manual_seed(0)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": fs, "featmap_nc": C_FEAT, "pred_img_size": S}), False, False).to(dev).train()
    early = [p for n, p in net.neural_render.named_parameters() if n != "bg_featmap"]
    bucket = dist_mod.GradBucket(net.parameters(), early=early)
    net.fuse_grad_accumulation()
    adam = hn.FusedAdam(bucket.params, lr=1e-4, bucket=bucket)
    lu = hn.HeadNeRFLossUtils(bg_type="white", use_vgg_loss=False, device=dev)
    rays_mode = args.shard == "rays"
    seed = 0 if rays_mode else rank                 # ray sharding: every rank holds the SAME batch and renders a slice of its rays
    host = O.synthetic_inputs(opt_o, B, seed=seed)
    gen = torch.Generator().manual_seed(50 + seed)
    host["gt"] = torch.rand(B, 3, S, S, generator=gen)
    host["mask"] = (torch.rand(B, 1, S, S, generator=gen) > 0.4).float()
    pinned = {k: v.contiguous().pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    if rays_mode:
        net.set_ray_sharding(rank, world)
    net.on_consumer_grads_ready = bucket.all_reduce_early
    loss_host = torch.zeros(4).pin_memory()
    M_rank = B * fs * fs * 64 // (world if rays_mode else 1)
    M_job = B * fs * fs * 64 * (1 if rays_mode else world)
    timer = hn.ops.TIMER

    def step(x):
        adam.zero_grad()
        out = net("train", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"], x["appea_code"],
                  x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
        loss = lu.calc_total_loss(None, None, out, x["gt"], x["mask"], None)["total_loss"]
        loss.backward()
        bucket.all_reduce(average=False)
        adam.step(grad_scale=1.0 / world)
        return loss

    copy_in = torch.cuda.Stream(device=dev)
    e2e_state = {"next": None, "pending": [], "count": 0, "last_loss": None}
    loss_slots = [torch.zeros(1).pin_memory() for _ in range(2)]

    def prefetch():
        """Next step's inputs (8.4 MB: images, masks, codes, cameras) travel on a copy stream while the current step computes."""
        with torch.cuda.stream(copy_in):
            x = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
            ev = torch.cuda.Event()
            ev.record(copy_in)
        return x, ev

    def step_e2e():
        """The training step from HOST buffers, as a prefetching loader + asynchronous logging would drive it: every step's inputs are
        copied from pinned memory and its loss is read back inside the timed region (talker_trainer.py:1069 logs it); the copy of
        step i+1 overlaps step i, and the host waits for step i-1's loss after step i is queued.  timed_region() drains the tail."""
        main = torch.cuda.current_stream()
        if e2e_state["next"] is None:
            e2e_state["next"] = prefetch()
        x, ev = e2e_state["next"]
        main.wait_event(ev)
        for t in x.values():
            t.record_stream(main)
        e2e_state["next"] = prefetch()
        loss = step(x)
        slot = loss_slots[e2e_state["count"] & 1]
        e2e_state["count"] += 1
        slot.copy_(loss.detach().reshape(1), non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        e2e_state["pending"].append((done, slot))
        if len(e2e_state["pending"]) > 1:
            d_ev, d_slot = e2e_state["pending"].pop(0)
            d_ev.synchronize()
            e2e_state["last_loss"] = float(d_slot[0])
        return e2e_state["last_loss"]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(resident)
    torch.cuda.synchronize()
    net.check_faults()
    timer.reset(); timer.enabled = True
    ms_total = timed_region(dist_mod, dev, lambda: step(resident), args.steps)
    timer.enabled = False
    kern, launches = timer.summary(), timer.launches
    ms_step = ms_total / args.steps
    sustained = sustained_run(dist_mod, dev, lambda: step(resident), ms_step, args.sustain_s, M_job)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(3):
        step_e2e()
    ms_e2e = timed_region(dist_mod, dev, step_e2e, args.steps)
    net.check_faults()
    finish(world)
    if rank != 0:
        return
    peaks = load_peaks()
    hot = sum(v["ms_avg"] * v["n"] / args.steps for k, v in kern.items() if k.startswith(("hn_mlp", "hn_composite", "hn_fold", "hn_loss_scale", "hn_pack")))
    out = {"metric": "ray_samples_per_sec_train_step_reso32hr", "value": round(M_job / (ms_step * 1e-3), 1), "unit": "ray*samples/s", "n_gpus": world,
           "steps": args.steps, "warmup": warm, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
           "scaling": "strong" if rays_mode else "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate (hot path), f32 (consumer, loss, Adam)",
           "data": "synthetic", "config": workload_config(3, world, args.shard),
           "e2e": {"value": round(M_job / (ms_e2e / args.steps * 1e-3), 1), "unit": "ray*samples/s",
                   "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in pinned.values()), "d2h_bytes_per_step": 4,
                   "ms_per_step": round(ms_e2e / args.steps, 4)},
           "sustained": sustained, "gpu_launches": launches, "gpu_launches_per_step": launches / args.steps,
           "library_ms_per_step": round(sum(v["ms_avg"] * v["n"] / args.steps for v in kern.values()), 4), "hot_path_ms_per_step": round(hot, 4),
           "roofline": {"kernel": "whole step: hot-path FLOPs (section 8d) over the step time; ~60 % of the step is the 512x512 NeuralRenderer (memory-bound tf32 convolutions + tails, DESIGN.md 5b)", "bound": "tensor",
                        "achieved": round(FLOP_STEP * M_rank / (ms_step * 1e-3) / 1e12, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": round(FLOP_STEP * M_rank / (ms_step * 1e-3) / 1e12 / peaks["tflops"], 4), "traffic": None, "peak_source": peaks["src"]},
           "kernels": {k: {"launches_per_step": v["n"] / args.steps, "ms_avg": round(v["ms_avg"], 4)} for k, v in kern.items()}, "clocks": clocks}
    if world == 1:
        out["cpu_baseline"] = cpu_arm(3, budget_s=15.0, repeats=1)["cpu_baseline"]
    print(json.dumps(out), file=RESULT_OUT, flush=True)


# =====================================================================================================================
# config 4 — fitting loop
# =====================================================================================================================
def run_config4(args):
    hn, O, rank, local, world, dev = setup("config 4")
    if world > 1:
        if rank != 0:
            finish(world)
            return
    fs, S = 32, 256
    iters = args.steps if args.steps_given else 500
    opt_o = O.OracleOptions(featmap_size=fs, pred_img_size=S)
    torch.manual_seed(0)
    net = hn.HeadNeRFNet(hn.BaseOptions({"featmap_size": fs, "featmap_nc": C_FEAT, "pred_img_size": S}), False, False).to(dev).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    host = O.synthetic_inputs(opt_o, 1, seed=0)
    gen = torch.Generator().manual_seed(4)
    host["gt"] = torch.rand(1, 3, S, S, generator=gen)
    host["mask"] = (torch.rand(1, 1, S, S, generator=gen) > 0.4).float()
    x = {k: v.to(dev) for k, v in host.items()}
    lu = hn.HeadNeRFLossUtils(use_vgg_loss=False, device=dev)
    off = {k: torch.zeros_like(x[k], requires_grad=True) for k in ("shape_code", "appea_code")}
    d_euler = torch.zeros(1, 3, device=dev, requires_grad=True)
    d_T = torch.zeros(1, 3, 1, device=dev, requires_grad=True)
    # FittingSingleImage_new.py:847-859: four Adam groups
    opt_fit = torch.optim.Adam([{"params": [off["shape_code"]], "lr": 0.015}, {"params": [off["appea_code"]], "lr": 0.01},
                                {"params": [d_euler], "lr": 0.001}, {"params": [d_T], "lr": 0.001}])
    loss_host = torch.zeros(1).pin_memory()
    timer = hn.ops.TIMER

    def fit_iter():
        opt_fit.zero_grad()
        dR = O.euler_to_rot(d_euler)
        R = dR @ x["batch_Rmats"]
        T = dR @ x["batch_Tvecs"] + d_T
        out = net("test", x["batch_xy"], None, x["audiostyle"], None, x["shape_code"] + off["shape_code"], x["appea_code"] + off["appea_code"],
                  R, T, x["batch_inv_inmats"])
        loss = lu.calc_total_loss(None, None, out, x["gt"], x["mask"], None)["total_loss"]
        loss.backward()
        opt_fit.step()
        return loss

    def fit_iter_e2e():
        loss = fit_iter()
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the reference's loop prints the loss every iteration (FittingSingleImage_new.py:895-900)

    sampler = ClockSampler(local)
    sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        fit_iter()
    torch.cuda.synchronize()
    net.check_faults()
    prec = net.last_meta["precision"]
    timer.reset(); timer.enabled = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fit_iter()
    e1.record()
    torch.cuda.synchronize()
    timer.enabled = False
    ms_total = e0.elapsed_time(e1)
    kern, launches = timer.summary(), timer.launches
    clocks = sampler.stop()
    e0.record()
    for _ in range(iters):
        fit_iter_e2e()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    net.check_faults()
    finish(world)
    M = 1024 * 64
    peaks = load_peaks()
    flop_iter = (FLOP_FWD + 2 * MAC_FWD) * M           # forward + data gradients incl. dL/dPE, no weight gradients (5 405 184 / ray*sample)
    ms_it = ms_total / iters
    out = {"metric": "ray_samples_per_sec_fitting_loop_reso32", "value": round(M / (ms_it * 1e-3), 1), "unit": "ray*samples/s", "n_gpus": 1,
           "steps": iters, "warmup": warm, "ms_per_step": round(ms_it, 4), "loop_seconds": round(ms_total * 1e-3, 3), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": f"precision auto -> {prec} (camera gradients: hi+lo split f16 operands, f32 accumulate)",
           "data": "synthetic", "config": workload_config(4, 1),
           "e2e": {"value": round(M / (ms_e2e / iters * 1e-3), 1), "unit": "ray*samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                   "ms_per_step": round(ms_e2e / iters, 4),
                   "note": "inputs of the loop are device-resident by definition (one image, fitted for 500 iterations); the per-iteration loss is read back"},
           "gpu_launches": launches, "gpu_launches_per_step": launches / iters,
           "library_ms_per_step": round(sum(v["ms_avg"] * v["n"] / iters for v in kern.values()), 4),
           "roofline": {"kernel": "whole iteration: hot-path FLOPs (section 8d) over the iteration time; the split-operand kernels issue 3x those FLOPs (DESIGN.md 6)", "bound": "tensor", "achieved": round(flop_iter / (ms_it * 1e-3) / 1e12, 1),
                        "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": round(flop_iter / (ms_it * 1e-3) / 1e12 / peaks["tflops"], 4), "traffic": None,
                        "peak_source": peaks["src"]},
           "kernels": {k: {"launches_per_step": v["n"] / iters, "ms_avg": round(v["ms_avg"], 4)} for k, v in kern.items()}, "clocks": clocks}
    out["cpu_baseline"] = cpu_arm(4, budget_s=15.0, repeats=1)["cpu_baseline"]
    print(json.dumps(out), file=RESULT_OUT, flush=True)


# =====================================================================================================================
# config 5 — forward-only sweep
# =====================================================================================================================
def run_config5(args):
    hn, O, rank, local, world, dev = setup("config 5")
    dist_mod = hn.dist
    points = []
    reps = max(2, min(args.steps, 5)) if args.steps_given else 3
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e2e = None
    launches = 0
    hn.ops.TIMER.reset()
    for ns in (32, 64, 128):
        opt_o = O.OracleOptions(featmap_size=64, pred_img_size=512, num_sample_coarse=ns)
        bo = hn.BaseOptions({"featmap_size": 64, "featmap_nc": C_FEAT, "pred_img_size": 512})
        bo.num_sample_coarse = ns
        torch.manual_seed(0)
        net = hn.HeadNeRFNet(bo, False, False).to(dev).eval()
        x = {k: v.to(dev) for k, v in O.synthetic_inputs(opt_o, 1, seed=0).items()}
        for n_rays in (16384, 262144, 1048576, 4194304):
            lo, hi = dist_mod.shard_range(n_rays // 2, rank, world)
            n_loc = (hi - lo) * 2
            gen = torch.Generator().manual_seed(n_rays + rank)
            xy_host = (torch.rand(1, 2, n_loc, generator=gen) * 64.0).pin_memory()
            xy = xy_host.to(dev)
            fn = lambda xy_=xy: net.render_rays("test", xy_, x["audiostyle"], x["shape_code"], x["appea_code"], x["batch_Rmats"], x["batch_Tvecs"], x["batch_inv_inmats"])
            with torch.no_grad():
                fn()
                ms = timed_region(dist_mod, dev, fn, reps) / reps
                if ns == 64 and n_rays == 4194304:
                    out_host = torch.empty(1, n_loc, C_FEAT + 1).pin_memory()

                    def fn_e2e():
                        Fm, bg = fn(xy_host.to(dev, non_blocking=True))
                        out_host[..., :C_FEAT].copy_(Fm, non_blocking=True)
                        out_host[..., C_FEAT].copy_(bg, non_blocking=True)
                        torch.cuda.current_stream().synchronize()
                    fn_e2e()
                    ms_e = timed_region(dist_mod, dev, fn_e2e, reps) / reps
                    e2e = {"value": round(n_rays * ns / (ms_e * 1e-3), 1), "unit": "ray*samples/s", "h2d_bytes_per_step": xy_host.numel() * 4 * world,
                           "d2h_bytes_per_step": out_host.numel() * 4 * world, "ms_per_step": round(ms_e, 4)}
            points.append({"rays": n_rays, "samples_per_ray": ns, "ms": round(ms, 4), "value": round(n_rays * ns / (ms * 1e-3), 1),
                           "frac_of_tensor_peak": round(FLOP_FWD * n_rays * ns / world / (ms * 1e-3) / 1e12 / load_peaks()["tflops"], 4)})
        net.check_faults()
    launches = hn.ops.TIMER.launches
    clocks = sampler.stop() if rank == 0 else None
    finish(world)
    if rank != 0:
        return
    peaks = load_peaks()
    head = [p for p in points if p["rays"] == 4194304 and p["samples_per_ray"] == 64][0]
    out = {"metric": "ray_samples_per_sec_fwd_sweep", "value": head["value"], "unit": "ray*samples/s", "n_gpus": world, "steps": reps, "warmup": 1,
           "ms_per_step": head["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate",
           "data": "synthetic", "config": workload_config(5, world), "e2e": e2e, "gpu_launches": launches, "sweep": points,
           "roofline": {"kernel": "hn_mlp_fwd (+ hn_composite_fwd) at 4M rays x 64 samples", "bound": "tensor",
                        "achieved": round(FLOP_FWD * 4194304 * 64 / world / (head["ms"] * 1e-3) / 1e12, 1), "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": head["frac_of_tensor_peak"], "traffic": None, "peak_source": peaks["src"]}, "clocks": clocks}
    if world == 1:
        out["cpu_baseline"] = cpu_arm(5, budget_s=10.0, repeats=1)["cpu_baseline"]
    print(json.dumps(out), file=RESULT_OUT, flush=True)


# =====================================================================================================================
# the CPU arm: reference algorithm (oracle port) on the host cores, chunks of CPU_CHUNK_RAYS rays per item
# =====================================================================================================================
def _cpu_problem(cfg):
    from oracle import headnerf_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    fs, S, B, mode = {2: (FS, S_IMG, B_PER_GPU, "train"), 3: (32, 512, 2, "train"), 4: (32, 256, 1, "test"), 5: (64, 512, 1, "test")}[cfg]
    opt = O.OracleOptions(featmap_size=fs, pred_img_size=S)
    torch.manual_seed(0)
    train = cfg in (2, 3)
    sd = {k: v.requires_grad_(train and not k.endswith(".f") and (cfg == 3 or k.startswith("fg_CD_predictor"))) for k, v in O.formula_state_dict(opt, "init").items()}
    inp = build_inputs(O, opt, B, seed=0, jitter=(mode == "train"))
    return O, opt, sd, inp, B, mode


def _cpu_chunk(O, opt, sd, inp, cfg, mode, lo, hi):
    """fwd (+bwd) of the hot path on rays [lo, hi) of every item: the reference algorithm's cost is per ray, so chunks of a fixed
    size time the same arithmetic as the whole batch while bounding memory (full Reso64 batch 2 keeps ~20 GB of activations)."""
    B = inp["batch_xy"].shape[0]
    xy = inp["batch_xy"][..., lo:hi].contiguous()
    tr = inp["t_rand"][:, lo:hi].contiguous() if "t_rand" in inp else None
    grad = cfg != 5
    codes = {k: inp[k].clone().requires_grad_(grad) for k in CODE_KEYS}
    cam = {k: inp[k].clone().requires_grad_(cfg == 4) for k in ("batch_Rmats", "batch_Tvecs")}
    with torch.set_grad_enabled(grad):
        r = O.render_features(sd, opt, mode, xy, codes["audiostyle"], codes["shape_code"], codes["appea_code"],
                              cam["batch_Rmats"], cam["batch_Tvecs"], inp["batch_inv_inmats"], t_rand=tr)
        if grad:
            gF = inp["gF"].view(B, -1, C_FEAT)[:, lo:hi].permute(0, 2, 1)
            gb = inp["g_bg"].view(B, 1, -1)[:, :, lo:hi]
            torch.autograd.backward([r["F"], r["bg_alpha"]], [gF, gb])
    return B * (hi - lo) * opt.num_sample_coarse


def cpu_arm(cfg, budget_s, repeats, steps=1, warmup=0):
    """Times the reference algorithm's hot path for `cfg` on all host threads.  One "step" = as many CPU_CHUNK_RAYS-ray chunks of the
    workload's batch as fit the time budget (all of them if the budget allows: then the step IS the full-size workload)."""
    O, opt, sd, inp, B, mode = _cpu_problem(cfg)
    n_r = inp["batch_xy"].shape[-1]
    chunk = min(CPU_CHUNK_RAYS, n_r)
    _cpu_chunk(O, opt, sd, inp, cfg, mode, 0, min(64, n_r))                       # page in, spin up the thread pool
    t0 = time.perf_counter()
    n0 = _cpu_chunk(O, opt, sd, inp, cfg, mode, 0, chunk)
    t_chunk = time.perf_counter() - t0
    total_chunks = (n_r + chunk - 1) // chunk
    per_step = max(1, min(total_chunks, int(budget_s / max(steps + warmup, 1) / max(t_chunk, 1e-6))))
    bounds = [(i * chunk, min(n_r, (i + 1) * chunk)) for i in range(per_step)]

    def one_step():
        return sum(_cpu_chunk(O, opt, sd, inp, cfg, mode, lo, hi) for lo, hi in bounds)

    for _ in range(warmup):
        one_step()
    best, total_t, n = 1e30, 0.0, 0
    for _ in range(max(steps, repeats)):
        t0 = time.perf_counter()
        n = one_step()
        dt = time.perf_counter() - t0
        best, total_t = min(best, dt), total_t + dt
    full = per_step == total_chunks
    what = {2: "hot path fwd+bwd (weight + code gradients)", 3: "hot path fwd+bwd of the training step (consumer, loss and Adam not included)",
            4: "hot path fwd+bwd to codes and camera, no weight gradients", 5: "hot path forward only"}[cfg]
    sample = (f"{'the whole batch' if full else 'rays [0, %d)' % bounds[-1][1]} of each of {B} item(s) x {opt.num_sample_coarse} samples = {n} ray*samples per step, "
              f"in chunks of {chunk} rays per item; {what}; oracle port of the reference (pinned bit-equal to it), fp32, {torch.get_num_threads()} threads")
    mean_t = total_t / max(steps, repeats)
    return {"value": n / mean_t, "ms_per_step": mean_t * 1e3, "full_size": full,
            "cpu_baseline": {"value": round(n / mean_t, 1), "unit": "ray*samples/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    budget_s = float(os.environ.get("HN_BENCH_REF_BUDGET_S", "150"))      # host seconds for all steps together (tests shrink it)
    steps = args.steps if (cfg != 4 or args.steps_given) else 20
    r = cpu_arm(cfg, budget_s, repeats=1, steps=steps, warmup=args.warmup)
    metric = {2: "ray_samples_per_sec_fwd_bwd_reso64", 3: "ray_samples_per_sec_train_step_reso32hr", 4: "ray_samples_per_sec_fitting_loop_reso32",
              5: "ray_samples_per_sec_fwd_sweep"}[cfg]
    out = {"impl": "reference", "metric": metric, "value": round(r["value"], 1), "unit": "ray*samples/s",
           "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 2),
           "higher_is_better": True, "scaling": "weak" if cfg in (2, 3, 4) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(cfg, args.gpus, args.shard), "full_size_step": r["full_size"],
           "cpu_baseline": r["cpu_baseline"],
           "e2e": {"value": round(r["value"], 1), "unit": "ray*samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configuration (2 = the headline)")
    ap.add_argument("--shard", default="items", choices=["items", "rays"], help="config 3: batch items per GPU (weak) or one batch ray-sharded (strong)")
    ap.add_argument("--sustain-s", type=float, default=2.5, help="length of the sustained run (seconds of device time)")
    ap.add_argument("--no-high", action="store_true", help="skip the auxiliary high-precision-mode measurement")
    args = ap.parse_args()
    args.steps_given = args.steps is not None
    if args.steps is None:
        args.steps = 50
    if args.impl == "reference":
        run_reference(args)
    else:
        {2: run_config2, 3: run_config3, 4: run_config4, 5: run_config5}[args.config](args)


if __name__ == "__main__":
    main()
