#!/bin/bash
# round 2, GPU run DD (2 GPUs): predictor integration fused into the pair kernel in the slab step: bitwise tests, N=2 bench with / without
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -30) > gpurun_out/dd_pytest.log 2>&1; tail -2 gpurun_out/dd_pytest.log
for F in 1 0; do
B200SPH_SLAB_FUSED_PREDICTOR=$F timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2960$F bench.py --gpus 2 --steps 20 --warmup 10 --quick > gpurun_out/dd_n2_f$F.json 2> gpurun_out/dd_n2_f$F.err; python -c "
import json; d=json.load(open('gpurun_out/dd_n2_f$F.json')); print('fused predictor $F: N=2', d['config']['workload'], 'ms/step', d['ms_per_step'], 'value', d['value'])"; tail -1 gpurun_out/dd_n2_f$F.err | cut -c1-200
done
for F in 1 0; do
B200SPH_SLAB_FUSED_PREDICTOR=$F timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$F bench.py --gpus 2 --steps 20 --warmup 10 --quick --workload dambreak2m > gpurun_out/dd_n2_2m_f$F.json 2> gpurun_out/dd_n2_2m_f$F.err; python -c "
import json; d=json.load(open('gpurun_out/dd_n2_2m_f$F.json')); print('fused predictor $F: N=2', d['config']['workload'], 'ms/step', d['ms_per_step'], 'value', d['value'])"
done
