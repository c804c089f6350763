#!/bin/bash
# Quick A/B on the GPU box: a subset of the parity tests + short bench runs of the library and of diagnostic variants.
#   usage: bash tools/gpu_quick.sh <tag> [variant ...]      (variants = build/libhn_<v>.so)
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests/test_gpu_backward.py tests/test_gpu_precise.py tests/test_gpu_full_size.py -m gpu -q --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.log
B="python bench.py --no-high --warmup 5"
$B --steps 30 --sustain-s 1.5 > $out/bench_c2.json 2> $out/bench_c2.err; echo "bench rc=$?"
for v in "$@"; do
  HN_LIB_PATH=build/libhn_$v.so $B --steps 15 --sustain-s 0.2 > $out/bench_$v.json 2> $out/bench_$v.err; echo "$v rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/bench_*.json")):
    try:
        d = json.load(open(f)); k = d["kernels"]
        print(f.split("bench_")[1][:-5].ljust(8), "step %.3f sust %.3f | fwd %.3f dgrad %.3f wgrad %.3f" % (d["ms_per_step"], d["sustained"]["ms_per_step"], k["hn_mlp_fwd"]["ms_avg"], k["hn_mlp_bwd_data"]["ms_avg"], k["hn_mlp_bwd_weights"]["ms_avg"]), d["roofline"]["all_mlp_kernels_frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
