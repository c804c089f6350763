#!/bin/bash
# A/B of library variants built by tools/build_variant.sh: prints name, ms/step and the per-kernel times.
#   usage: bash tools/sweep_lib.sh default pf4 pf16 ...      ("default" = the in-tree library)
for v in "$@"; do
  if [ "$v" = default ]; then unset HN_LIB_PATH; else export HN_LIB_PATH=$PWD/build/$v.so; fi
  timeout 100 python bench.py --steps 20 --warmup 3 --no-high 2>/dev/null | V=$v python -c '
import json, os, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(os.environ["V"], d["ms_per_step"], {k: v["ms_avg"] for k, v in d["kernels"].items() if "mlp" in k})'
done
