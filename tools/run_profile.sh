#!/bin/bash
# One profiling call on the GPU box (B200_PROFILING.md recipe): plain bench first, then the ncu launch list and one
# --set full capture of OUR kernels of the same command.   usage (through gpurun): bash tools/run_profile.sh <tag>
# Afterwards, here: python tools/make_profiles.py <tag> "<title>"
tag=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-high --sustain-s 0.01"
$CMD > gpurun_out/plain_${tag}.log 2> gpurun_out/plain_${tag}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${tag}.err; exit 1; }
# ncu serialises kernels: the weight-gradient items that normally run CONCURRENTLY on the idle SMs (second stream) would show up as
# a 13-CTA kernel running alone, so the captures run with that split disabled (HN_WGRAD_SIDE=0); the plain run above keeps it
export HN_WGRAD_SIDE=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${tag}.csv $CMD > gpurun_out/ncu1_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mlp_|composite" -s 14 -c 7 -f -o gpurun_out/prof_${tag} $CMD > gpurun_out/ncu2_${tag}.log 2>&1
wc -l gpurun_out/launches_${tag}.csv; ls -la gpurun_out/prof_${tag}.ncu-rep
