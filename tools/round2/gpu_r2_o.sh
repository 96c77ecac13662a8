#!/bin/bash
# round 2, GPU run O (1 GPU): which of the three pair-loop changes costs time (A/B of compile-time variants, dambreak2m)
mkdir -p gpurun_out
for V in base e1 r1 p1 r1p1 default; do
  if [ $V = default ]; then unset B200SPH_LIB; else export B200SPH_LIB=$PWD/build/variants/libb200sph_$V.so; fi
  timeout 300 python bench.py --workload dambreak2m --quick --steps 20 --warmup 10 > gpurun_out/o_$V.json 2> gpurun_out/o_$V.err; python -c "
import json; d=json.load(open('gpurun_out/o_$V.json')); print('$V ms/step', round(d['ms_per_step'],4), 'forces kernel ms', round(d['roofline']['kernel_ms'],4))"
done
