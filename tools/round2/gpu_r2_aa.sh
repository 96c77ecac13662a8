#!/bin/bash
# round 2, GPU run AA (1 GPU): the tree as it stands: full GPU suite, smoke, default bench (both arms), launch list of a bench run
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/aa_pytest.log 2>&1; tail -3 gpurun_out/aa_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/aa_ref.json 2> gpurun_out/aa_ref.err; cut -c1-300 gpurun_out/aa_ref.json
timeout 600 python bench.py > gpurun_out/aa_ours.json 2> gpurun_out/aa_ours.err; python -c "
import json; d=json.load(open('gpurun_out/aa_ours.json')); print('ours', d['config']['workload'], 'ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', d['e2e'], 'launches', d['gpu_launches'], 'clocks', d['clocks'], 'cpu', d['cpu_baseline']); print(d['roofline'])"; tail -2 gpurun_out/aa_ours.err | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/aa_launches.csv python bench.py --steps 10 --warmup 3 --quick > gpurun_out/aa_ncu.log 2>&1; tail -1 gpurun_out/aa_ncu.log | cut -c1-200
