#!/bin/bash
# round 2, GPU run S (1 GPU): new pair-kernel defaults (ratio-space physics, 8 CTAs/SM, one list row ahead): full GPU suite,
# 2M / 8M / poiseuille timing, ncu of the pair kernel
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30) > gpurun_out/s_pytest.log 2>&1; tail -3 gpurun_out/s_pytest.log
for W in dambreak2m dambreak8m poiseuille1m; do
timeout 300 python bench.py --workload $W --quick --steps 20 --warmup 10 > gpurun_out/s_$W.json 2> gpurun_out/s_$W.err; python -c "
import json; d=json.load(open('gpurun_out/s_$W.json')); print('$W ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'kernel ms', round(d['roofline']['kernel_ms'],4), 'rebuild', round(d['roofline']['neighbour_rebuild_ms'],3))"; tail -2 gpurun_out/s_$W.err | cut -c1-300
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forces_gather -s 6 -c 1 -f -o gpurun_out/prof_pair_final python bench.py --workload dambreak2m --steps 3 --warmup 3 --quick > gpurun_out/s_ncu.log 2>&1; tail -1 gpurun_out/s_ncu.log | cut -c1-200
