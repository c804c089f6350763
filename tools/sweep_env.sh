#!/bin/bash
# sweep of one environment knob: bash tools/sweep_env.sh NAME v1 v2 ...   (prints value, ms/step, ms of hn_mlp_bwd_weights)
name=$1; shift
for v in "$@"; do
  env $name=$v timeout 100 python bench.py --steps 20 --warmup 3 --no-high 2>/dev/null | V=$v python -c '
import json, os, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(os.environ["V"], d["ms_per_step"], d["kernels"]["hn_mlp_bwd_weights"]["ms_avg"])'
done
