#!/bin/bash
# round 2, GPU run CC (1 GPU): the 16 M problem on one GPU with the final tree (the N=1 point of the strong-scaling table), 2M full line
mkdir -p gpurun_out
timeout 900 python bench.py --workload dambreak16m --steps 10 --warmup 10 --no-cpu-baseline > gpurun_out/cc_ours_16m_n1.json 2> gpurun_out/cc_err16.log; python -c "
import json; d=json.load(open('gpurun_out/cc_ours_16m_n1.json')); print('16m N=1 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'kernel', d['roofline']['kernel_ms'], 'rebuild', d['roofline']['neighbour_rebuild_ms'])"; tail -2 gpurun_out/cc_err16.log | cut -c1-300
timeout 600 python bench.py --workload dambreak2m --no-cpu-baseline > gpurun_out/cc_ours_2m.json 2> gpurun_out/cc_err2.log; python -c "
import json; d=json.load(open('gpurun_out/cc_ours_2m.json')); print('2m ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'e2e value', d['e2e']['value'])"
