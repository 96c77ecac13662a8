#!/bin/bash
# round 2, GPU run Z (1 GPU): large list blocks (256k / 1M / 2M particles): the pair kernel likes the interleaved layout, the list builder rows that are not too far apart
mkdir -p gpurun_out
for V in b256k b1m b2m; do
  export B200SPH_LIB=$PWD/build/variants/libb200sph_$V.so
  for W in dambreak2m dambreak8m; do
  timeout 300 python bench.py --workload $W --quick --steps 20 --warmup 10 > gpurun_out/z_${V}_$W.json 2> gpurun_out/z_${V}_$W.err; python -c "
import json; d=json.load(open('gpurun_out/z_${V}_$W.json')); print('$V $W ms/step', round(d['ms_per_step'],4), 'forces kernel ms', round(d['roofline']['kernel_ms'],4), 'rebuild', round(d['roofline']['neighbour_rebuild_ms'],3))"
  done
done
