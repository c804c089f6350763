#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/probe_fill.log
: > $out
run() { timeout 60 ./build/probe_fill "$@" >> $out 2>&1; echo "rc=$? args=$*" >> $out; }
# mech CL stages box_rows wait_flavor(>=10: same-warp lanes)
run 3 1 2 128 0
run 3 1 2 128 10
run 3 1 2 64 10
cat $out
