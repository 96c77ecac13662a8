#!/bin/bash
# GPU box: one full ncu capture of the first launch of a kernel (regex) in a short bench run
# usage: tools/gpu_ncu_kernel.sh <kernel-regex> <output-name> [ENV=..]...
K=$1; O=$2; shift 2
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -c 1 -f -o gpurun_out/$O python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$O.log 2>&1; tail -2 gpurun_out/ncu_$O.log | cut -c1-300
