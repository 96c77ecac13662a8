#!/bin/bash
# round 2, GPU run W (1 GPU): block size of the neighbour-list layout (128 = default build, 1k, 4k, 16k, 64k particles) at 2M and 8M
mkdir -p gpurun_out
for V in default b1k b4k b16k b64k; do
  if [ $V = default ]; then unset B200SPH_LIB; else export B200SPH_LIB=$PWD/build/variants/libb200sph_$V.so; fi
  for W in dambreak2m dambreak8m; do
  timeout 300 python bench.py --workload $W --quick --steps 20 --warmup 10 > gpurun_out/w_${V}_$W.json 2> gpurun_out/w_${V}_$W.err; python -c "
import json; d=json.load(open('gpurun_out/w_${V}_$W.json')); print('$V $W ms/step', round(d['ms_per_step'],4), 'forces kernel ms', round(d['roofline']['kernel_ms'],4), 'rebuild', round(d['roofline']['neighbour_rebuild_ms'],3))"
  done
done
