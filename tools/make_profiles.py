"""Turns the ncu outputs of one profiling call (gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep) into the tracked
summaries under profiles/: launch list with shares, per-kernel --set full summary, and the DRAM-traffic table bench.py reads.
   usage: python tools/make_profiles.py <tag> [title]          (runs here, no GPU: needs the ncu CLI)"""
import collections, csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else tag
go = os.path.join(ROOT, "gpurun_out")
prof = os.path.join(ROOT, "profiles")

# ---- launch list
rows = [r for r in csv.reader(l for l in open(os.path.join(go, f"launches_{tag}.csv")) if not l.startswith("=="))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki])
    unit = r[hdr.index("Metric Unit")]
    v = float(r[vi].replace(",", "")) * ({"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0))
    d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(prof, f"{tag}_launch_list.md"), "w") as f:
    f.write(f"# ncu launch list, {title}: first {sum(v[0] for v in agg.values())} launches of `python bench.py --steps 3 --warmup 3` (Reso64 batch 2)\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 120` - cold-cache, serialised times: compare SHARES, not absolutes.\n\n")
    f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        f.write(f"| {k[:90]} | {n} | {us:.1f} | {100 * us / tot:.1f} % |\n")

# ---- full capture
raw = subprocess.run(["ncu", "-i", os.path.join(go, f"prof_{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum"]
seen, traffic = set(), {}
names = {"mlp_fwd_kernel": "hn_mlp_fwd", "mlp_chain_kernel<0>": "hn_mlp_fwd", "mlp_chain_kernel<1>": "hn_mlp_bwd_data",
         "composite_fwd_kernel": "hn_composite_fwd", "composite_bwd_kernel": "hn_composite_bwd",
         "mlp_bwd_kernel": "hn_mlp_bwd_data", "mlp_wgrad_kernel": "hn_mlp_bwd_weights"}
with open(os.path.join(prof, f"{tag}_ncu_full_summary.md"), "w") as f:
    f.write(f"# ncu --set full summary, {title} (bench.py --steps 3 --warmup 3; Reso64 batch 2)\n\n")
    f.write("Source: `ncu --set full --clock-control none --import-source on -k regex:\"mlp_|composite\" -s 12 -c 6` on a B200 (gpurun); the .ncu-rep is not committed.\n")
    f.write("Times under ncu are cold-cache and serialised; the bench's CUDA-event times are the reported ones.\n")
    f.write("Captured with HN_WGRAD_SIDE=0 (the second-stream split of the weight-gradient pass is meaningless under ncu's serialisation).\n")
    for r in rr[2:]:
        kname = r[h.index("Kernel Name")]
        short = re.sub(r"\(.*", "", kname)
        if short in seen:
            continue
        seen.add(short)
        f.write(f"\n## {kname}\n\n| metric | value | unit |\n|---|---|---|\n")
        vals = {}
        for m in want:
            if m in h:
                vals[m] = r[h.index(m)]; f.write(f"| {m} | {r[h.index(m)]} | {units[h.index(m)]} |\n")
        def gb(m):
            u = units[h.index(m)]; v = float(vals[m].replace(",", ""))
            return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}[u]
        tr = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
        f.write(f"| dram traffic (read+write) | {tr:.3f} | Gbyte |\n")
        for key, api in names.items():
            if key in short and api in traffic and key == "mlp_wgrad_kernel":       # cluster + remainder launches of one call
                traffic[api]["dram_gbytes_per_launch"] = round(traffic[api]["dram_gbytes_per_launch"] + tr, 4)
            elif key in short and api not in traffic:
                traffic[api] = {"dram_gbytes_per_launch": round(tr, 4),
                                "tensor_pipe_pct": round(float(vals["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]), 2),
                                "dram_pct": round(float(vals["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]), 2)}
prev = os.path.join(prof, "r01h_traffic.json")
if os.path.isfile(prev):                                                           # kernels outside this capture window keep their last figure
    for k, v in json.load(open(prev))["kernels"].items():
        traffic.setdefault(k, dict(v, note="from profiles/r01h_traffic.json (kernel not in this capture)"))
with open(os.path.join(prof, f"{tag}_traffic.json"), "w") as f:
    json.dump({"source": f"profiles/{tag}_ncu_full_summary.md (ncu --set full, Reso64 batch 2)", "kernels": traffic}, f, indent=1)
print(json.dumps(traffic, indent=1))
