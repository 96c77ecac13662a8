#!/bin/bash
# round 2, GPU run EE (1 GPU): final tree: full GPU suite (incl. the 7.87M list-layout test), smoke, default bench line
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/ee_pytest.log 2>&1; tail -3 gpurun_out/ee_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/ee_ours.json 2> gpurun_out/ee_ours.err; python -c "
import json; d=json.load(open('gpurun_out/ee_ours.json')); print('ours', d['config']['workload'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'launches', d['gpu_launches'])"; tail -2 gpurun_out/ee_ours.err | cut -c1-300
