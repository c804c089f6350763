#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -o build/probe_umma tools/probe_umma.cu
