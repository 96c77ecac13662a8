#!/bin/bash
# round 2, GPU run N (1 GPU): pair kernel with the ratio-space physics and the copy-free look-ahead: parity + timing
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extras.py tests/test_golden.py -m gpu -q 2>&1 | tail -30) > gpurun_out/n_pytest.log 2>&1; tail -4 gpurun_out/n_pytest.log
for W in dambreak2m dambreak8m; do
timeout 300 python bench.py --workload $W --quick --steps 20 --warmup 10 > gpurun_out/n_$W.json 2> gpurun_out/n_$W.err; python -c "
import json; d=json.load(open('gpurun_out/n_$W.json')); print('$W ms/step', d['ms_per_step'], 'value', d['value'], 'roofline', d['roofline'])"; tail -2 gpurun_out/n_$W.err | cut -c1-300
done
