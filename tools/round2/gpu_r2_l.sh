#!/bin/bash
# round 2, GPU run L (8 GPUs): time line of the resident N=8 step; N=8 and N=2 bench with the chained slab step_host
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/diag_slab_step.py > gpurun_out/l_step_n8.txt 2> gpurun_out/l_step_n8.err; head -12 gpurun_out/l_step_n8.txt | cut -c1-400; tail -3 gpurun_out/l_step_n8.err | cut -c1-300
for N in 8 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 10 > gpurun_out/l_ours_n$N.json 2> gpurun_out/l_ours_n$N.err; python -c "
import json; d=json.load(open('gpurun_out/l_ours_n$N.json')); print('N=$N', d['config']['workload'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'e2e value', d['e2e']['value'])"; tail -2 gpurun_out/l_ours_n$N.err | cut -c1-300
done
