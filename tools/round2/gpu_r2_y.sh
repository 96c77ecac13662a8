#!/bin/bash
# round 2, GPU run Y (1 GPU): one block of 4M particles = the interleaved layout through the blocked code path
mkdir -p gpurun_out
for V in bbig default; do
  if [ $V = default ]; then unset B200SPH_LIB; else export B200SPH_LIB=$PWD/build/variants/libb200sph_$V.so; fi
  timeout 300 python bench.py --workload dambreak2m --quick --steps 20 --warmup 10 > gpurun_out/y_$V.json 2> gpurun_out/y_$V.err; python -c "
import json; d=json.load(open('gpurun_out/y_$V.json')); print('$V ms/step', round(d['ms_per_step'],4), 'forces kernel ms', round(d['roofline']['kernel_ms'],4), 'rebuild', round(d['roofline']['neighbour_rebuild_ms'],3))"
done
