#!/bin/bash
# round 2, GPU run H (2 GPUs): 2-GPU tests (resident + host state), N=2 bench with the pipelined slab e2e, Poiseuille ppH 100
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -30) > gpurun_out/h_pytest.log 2>&1; tail -3 gpurun_out/h_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 10 > gpurun_out/h_ours_n2.json 2> gpurun_out/h_ours_n2.err; python -c "
import json; d=json.load(open('gpurun_out/h_ours_n2.json')); print('N=2', d['config']['workload'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'upd/s', d['particle_updates_per_s'], 'launches', d['gpu_launches'])"; tail -5 gpurun_out/h_ours_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --workload dambreak16m --steps 10 --warmup 10 --no-cpu-baseline > gpurun_out/h_ours_16m_n1.json 2> gpurun_out/h_err16.log; python -c "
import json; d=json.load(open('gpurun_out/h_ours_16m_n1.json')); print('16m N=1 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'])"
(CUDA_VISIBLE_DEVICES=1 timeout 1200 python tools/validate_poiseuille.py --ppH 100 --iters 2000 --out gpurun_out/h_poiseuille100.json 2>&1 | tail -3) > gpurun_out/h_poiseuille.log 2>&1; tail -2 gpurun_out/h_poiseuille.log | cut -c1-900
