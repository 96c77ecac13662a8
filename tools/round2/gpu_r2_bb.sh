#!/bin/bash
# round 2, GPU run BB (2 GPUs): final tree: 2-GPU tests (resident, host state, periodic slab axis) and the N=2 bench, both arms
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -30) > gpurun_out/bb_pytest.log 2>&1; tail -2 gpurun_out/bb_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 20 --warmup 10 > gpurun_out/bb_ours_n2.json 2> gpurun_out/bb_ours_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bb_ours_n2.json')); print('N=2', d['config']['workload'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'e2e value', d['e2e']['value'])"; tail -2 gpurun_out/bb_ours_n2.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --impl reference --gpus 2 --steps 20 --warmup 10 > gpurun_out/bb_ref_n2.json 2> gpurun_out/bb_ref_n2.err; cut -c1-400 gpurun_out/bb_ref_n2.json
