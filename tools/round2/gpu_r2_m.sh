#!/bin/bash
# round 2, GPU run M (2 GPUs): ring-of-stripes host step (periodic COORD3) tests, poiseuille1m e2e, N=2 e2e with finer pieces
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_extras.py tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -30) > gpurun_out/m_pytest.log 2>&1; tail -3 gpurun_out/m_pytest.log
timeout 300 python bench.py --workload poiseuille1m --no-cpu-baseline --steps 20 --warmup 10 > gpurun_out/m_pois1m.json 2> gpurun_out/m_pois1m.err; python -c "
import json; d=json.load(open('gpurun_out/m_pois1m.json')); print('pois1m ms/step', d['ms_per_step'], 'e2e', d['e2e'])"; tail -2 gpurun_out/m_pois1m.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 10 > gpurun_out/m_ours_n2.json 2> gpurun_out/m_ours_n2.err; python -c "
import json; d=json.load(open('gpurun_out/m_ours_n2.json')); print('N=2', d['config']['workload'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'e2e value', d['e2e']['value'])"; tail -2 gpurun_out/m_ours_n2.err | cut -c1-300
