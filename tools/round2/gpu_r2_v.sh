#!/bin/bash
# round 2, GPU run V (1 GPU): blocked list layout: L1 policy of the list stream (default / no_allocate / evict_first), then the GPU suite
mkdir -p gpurun_out
for V in default lc1 lc2; do
  if [ $V = default ]; then unset B200SPH_LIB; else export B200SPH_LIB=$PWD/build/variants/libb200sph_$V.so; fi
  timeout 300 python bench.py --workload dambreak2m --quick --steps 20 --warmup 10 > gpurun_out/v_$V.json 2> gpurun_out/v_$V.err; python -c "
import json; d=json.load(open('gpurun_out/v_$V.json')); print('$V ms/step', round(d['ms_per_step'],4), 'forces kernel ms', round(d['roofline']['kernel_ms'],4), 'rebuild', round(d['roofline']['neighbour_rebuild_ms'],3))"
done
unset B200SPH_LIB
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/v_pytest.log 2>&1; tail -3 gpurun_out/v_pytest.log
