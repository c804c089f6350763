#!/bin/bash
# One GPU-box session of round 2 (run through gpurun): parity tests, the bench, A/B switches of the weight-gradient kernel and one
# ncu --set full capture of the MLP kernels.   usage: bash tools/gpu_session.sh <tag>
tag=${1:-r2c}
out=gpurun_out/$tag
mkdir -p $out
python -m pytest tests -m gpu -q --timeout 900 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
B="python bench.py --no-high --warmup 5"
$B --steps 30 --sustain-s 1.5 > $out/bench_c2.json 2> $out/bench_c2.err; echo "bench rc=$?"
HN_WGRAD_DUOS=0 $B --steps 15 --sustain-s 0.2 > $out/bench_noduo.json 2> $out/bench_noduo.err; echo "noduo rc=$?"
HN_WGRAD_SIDE=0 $B --steps 15 --sustain-s 0.2 > $out/bench_noside.json 2> $out/bench_noside.err; echo "noside rc=$?"
export HN_WGRAD_SIDE=0
ncu --set full --clock-control none --import-source on -k regex:"mlp_chain|mlp_wgrad" -s 10 -c 5 -f -o $out/prof python bench.py --steps 2 --warmup 3 --no-high --sustain-s 0.01 > $out/ncu.log 2>&1
ls -la $out/prof.ncu-rep
