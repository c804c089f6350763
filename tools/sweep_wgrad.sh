#!/bin/bash
# sweep of the weight-gradient cost-model slope (HN_WGRAD_SLOPE): prints slope, ms/step, ms of hn_mlp_bwd_weights
for sl in "$@"; do
  HN_WGRAD_SLOPE=$sl timeout 100 python bench.py --steps 20 --warmup 3 --no-high 2>/dev/null | SL=$sl python -c '
import json, os, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(os.environ["SL"], d["ms_per_step"], d["kernels"]["hn_mlp_bwd_weights"]["ms_avg"])'
done
