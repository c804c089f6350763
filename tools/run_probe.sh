#!/bin/bash
# Runs every probe configuration as its own process (a faulting config must not hide the rest).
# usage: tools/run_probe.sh  (on a GPU box, after tools/build_probe.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/probe_umma.log
: > $out
run() { timeout 60 ./build/probe_umma "$@" >> $out 2>&1; echo "rc=$? args=$*" >> $out; }
# K-major x K-major (forward / data-gradient GEMMs)
run 0 0 0 0 128 64
run 0 0 0 0 128 128
run 0 0 0 0 128 256
run 0 0 0 0 64 128
run 0 0 0 0 16 128
run 0 0 0 0 256 128
run 0 0 0 0 192 64
# MN-major x MN-major (weight-gradient GEMMs: contraction over the 128 rows)
run 1 1 0 0 128 128
run 1 1 0 0 256 128
run 1 1 0 0 64 128
run 1 1 0 0 128 256
# mixed major
run 1 0 0 0 128 128
run 0 1 0 0 128 128
# formats: bf16 x bf16, and mixed f16/bf16 operands
run 0 0 1 1 128 128
run 0 0 1 0 128 128
run 0 0 0 1 128 128
run 1 1 1 0 128 128
run 1 1 0 1 128 128
cat $out
