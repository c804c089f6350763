#!/bin/bash
# Multi-GPU session (gpurun --gpus N): NCCL 1-vs-2-rank equivalence test, then the bench configurations at N GPUs.
N=${1:-2}; tag=${2:-r2m}
out=gpurun_out/$tag; mkdir -p $out
python -m pytest tests/test_gpu_multi.py -m gpu -q -s --timeout 900 > $out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -3 $out/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
$TR bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_c2_n$N.json 2> $out/bench_c2_n$N.err; echo "c2 rc=$?"
$TR bench.py --gpus $N --config 3 --shard items --steps 20 > $out/bench_c3_items_n$N.json 2> $out/bench_c3_items_n$N.err; echo "c3 items rc=$?"
$TR bench.py --gpus $N --config 3 --shard rays --steps 20 > $out/bench_c3_rays_n$N.json 2> $out/bench_c3_rays_n$N.err; echo "c3 rays rc=$?"
$TR bench.py --gpus $N --config 5 > $out/bench_c5_n$N.json 2> $out/bench_c5_n$N.err; echo "c5 rc=$?"
python bench.py --config 3 --steps 20 > $out/bench_c3_n1.json 2> $out/bench_c3_n1.err; echo "c3 n1 rc=$?"
python bench.py --config 5 > $out/bench_c5_n1.json 2> $out/bench_c5_n1.err; echo "c5 n1 rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("$out/bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("bench_")[1][:-5].ljust(14), d["metric"], "%.3e" % d["value"], "ms/step", d["ms_per_step"], "e2e", "%.3e" % d["e2e"]["value"])
    except Exception as e:
        print(f, "ERR", e)
PY
