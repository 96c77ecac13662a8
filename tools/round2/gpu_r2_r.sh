#!/bin/bash
# round 2, GPU run R (1 GPU): occupancy of the trimmed pair loop: 8/9/10 CTAs per SM, with and without the constants pinned in registers
mkdir -p gpurun_out
for V in mb8 mb8pf1 mb9pf1 h0mb8 h0mb9 h0mb10; do
  export B200SPH_LIB=$PWD/build/variants/libb200sph_$V.so
  timeout 300 python bench.py --workload dambreak2m --quick --steps 20 --warmup 10 > gpurun_out/r_$V.json 2> gpurun_out/r_$V.err; python -c "
import json; d=json.load(open('gpurun_out/r_$V.json')); print('$V ms/step', round(d['ms_per_step'],4), 'forces kernel ms', round(d['roofline']['kernel_ms'],4))"
done
