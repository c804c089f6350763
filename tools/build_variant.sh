#!/bin/bash
# Diagnostic build of the SAME library with extra -D flags into build/<name>.so (loaded with HN_LIB_PATH; never a fallback).
#   usage: bash tools/build_variant.sh <name> <flags...>
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/var_$name
objs=""
for f in hn_api hn_composite hn_mlp_sched hn_mlp_pack hn_mlp_fwd hn_mlp_bwd hn_mlp_wgrad hn_precise hn_fold hn_render2d hn_render hn_train hn_fine hn_nr; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DHN_BUILDING_DSO "$@" -c nerf-3dtalker-code_b200/csrc/$f.cu -o build/var_$name/$f.o &
  objs="$objs build/var_$name/$f.o"
done
wait
nvcc -shared -o build/$name.so $objs
echo build/$name.so
