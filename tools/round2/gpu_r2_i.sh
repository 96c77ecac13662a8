#!/bin/bash
# round 2, GPU run I (2 GPUs): pipelined slab host step: bitwise test + N=2 bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/i_pytest.log 2>&1; tail -3 gpurun_out/i_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 10 > gpurun_out/i_ours_n2.json 2> gpurun_out/i_ours_n2.err; python -c "
import json; d=json.load(open('gpurun_out/i_ours_n2.json')); print('N=2', d['config']['workload'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'e2e value', d['e2e']['value'])"; tail -5 gpurun_out/i_ours_n2.err | cut -c1-300
