#!/bin/bash
# round 2, GPU run Q (1 GPU): occupancy / list read-ahead variants of the trimmed pair loop (dambreak2m), parity of the new default
mkdir -p gpurun_out
for V in default mb8 mb6 pf1 pf3; do
  if [ $V = default ]; then unset B200SPH_LIB; else export B200SPH_LIB=$PWD/build/variants/libb200sph_$V.so; fi
  timeout 300 python bench.py --workload dambreak2m --quick --steps 20 --warmup 10 > gpurun_out/q_$V.json 2> gpurun_out/q_$V.err; python -c "
import json; d=json.load(open('gpurun_out/q_$V.json')); print('$V ms/step', round(d['ms_per_step'],4), 'forces kernel ms', round(d['roofline']['kernel_ms'],4))"
done
unset B200SPH_LIB
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extras.py tests/test_golden.py -m gpu -q 2>&1 | tail -30) > gpurun_out/q_pytest.log 2>&1; tail -2 gpurun_out/q_pytest.log
