#!/bin/bash
# usage: tools/run_probe_ts.sh  (on a GPU box; binary built here by nvcc into build/)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/probe_ts.log
: > $out
run() { timeout 60 ./build/probe_ts "$@" >> $out 2>&1; echo "rc=$? args=$*" >> $out; }
# args: mode N ts CL reps fills replicas pieces stages
for st in 2 4 6 8; do run 2 1 0 0 0 4096 1 1 $st; done
for st in 4 8; do run 2 1 0 0 0 4096 148 1 $st; done
run 2 1 0 0 0 4096 1 4 8
for st in 4 8; do run 2 2 0 0 0 4096 1 1 $st; done
run 2 3 0 0 0 4096 1 1 6
for st in 4 8; do run 2 4 0 0 0 4096 1 1 $st; done
# MMA rate with fill
for n in 128; do run 1 $n 0 1 8192 4096 1 1 8; run 1 $n 1 1 8192 4096 1 1 8; run 1 $n 1 2 8192 4096 1 1 8; run 1 $n 1 4 8192 4096 1 1 8; done
cat $out
