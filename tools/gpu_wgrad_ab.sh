#!/bin/bash
# per-kernel durations of the weight-gradient pass under the A/B switches (ncu, serialised, second stream off)
out=gpurun_out/$1; mkdir -p $out
export HN_WGRAD_SIDE=0
run() { name=$1; shift; env "$@" ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:mlp_wgrad -s 9 -c 6 --csv python bench.py --steps 2 --warmup 3 --no-high --sustain-s 0.01 2>/dev/null | grep -E "mlp_wgrad" | awk -F'","' '{print $5, $(NF-2), $(NF)}' | sed "s/^/$name: /" | tee -a $out/wgrad_ab.log; }
run default X=1
run nofold HN_WGRAD_DENS_FOLD=0
run noduo HN_WGRAD_DUOS=0
run old HN_WGRAD_DENS_FOLD=0 HN_WGRAD_DUOS=0
