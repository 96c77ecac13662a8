#!/bin/bash
# round 2, GPU run HH (4 GPUs): final tree, N=4 bench line (strong scaling of the 16 M problem)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 4 --steps 20 --warmup 10 > gpurun_out/hh_ours_n4.json 2> gpurun_out/hh_ours_n4.err; python -c "
import json; d=json.load(open('gpurun_out/hh_ours_n4.json')); print('N=4', d['config']['workload'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'e2e value', d['e2e']['value'])"; tail -2 gpurun_out/hh_ours_n4.err | cut -c1-300
