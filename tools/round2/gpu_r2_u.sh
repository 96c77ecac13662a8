#!/bin/bash
# round 2, GPU run U (1 GPU): blocked neighbour-list layout: full GPU suite, 2M / 8M timing
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/u_pytest.log 2>&1; tail -3 gpurun_out/u_pytest.log
for W in dambreak2m dambreak8m; do
timeout 300 python bench.py --workload $W --quick --steps 20 --warmup 10 > gpurun_out/u_$W.json 2> gpurun_out/u_$W.err; python -c "
import json; d=json.load(open('gpurun_out/u_$W.json')); print('$W ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'kernel ms', round(d['roofline']['kernel_ms'],4), 'rebuild', round(d['roofline']['neighbour_rebuild_ms'],3))"; tail -2 gpurun_out/u_$W.err | cut -c1-300
done
