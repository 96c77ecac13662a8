#!/bin/bash
# round 2, GPU run X (1 GPU): blocked list: where do the 4 % of the pair kernel go? previous commit's library / blocked / blocked with a constant row step
mkdir -p gpurun_out
for V in head default al; do
  if [ $V = default ]; then unset B200SPH_LIB; else export B200SPH_LIB=$PWD/build/variants/libb200sph_$V.so; fi
  timeout 300 python bench.py --workload dambreak2m --quick --steps 20 --warmup 10 > gpurun_out/x_$V.json 2> gpurun_out/x_$V.err; python -c "
import json; d=json.load(open('gpurun_out/x_$V.json')); print('$V ms/step', round(d['ms_per_step'],4), 'forces kernel ms', round(d['roofline']['kernel_ms'],4), 'rebuild', round(d['roofline']['neighbour_rebuild_ms'],3))"
done
