#!/bin/bash
# round 2, GPU run J (8 GPUs): slab tests at world 2/4/8 (incl. periodic slab axis), N=4 and N=8 bench, reference N=8
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/j_pytest.log 2>&1; tail -3 gpurun_out/j_pytest.log
for N in 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 10 > gpurun_out/j_ours_n$N.json 2> gpurun_out/j_ours_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/j_ours_n$N.json')); print('N=$N', d['config']['workload'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'e2e value', d['e2e']['value'])"; tail -3 gpurun_out/j_ours_n$N.err | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 8 --steps 20 --warmup 10 > gpurun_out/j_ref_n8.json 2> gpurun_out/j_ref_n8.err; cut -c1-600 gpurun_out/j_ref_n8.json; tail -3 gpurun_out/j_ref_n8.err | cut -c1-300
