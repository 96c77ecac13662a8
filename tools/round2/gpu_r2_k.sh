#!/bin/bash
# round 2, GPU run K (4 GPUs): slab step_host chained piece by piece: bitwise test, time line at N=4 with / without NUMA binding
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|model name") > gpurun_out/k_topo.txt 2>&1
(timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -k "two_gpu or periodic" 2>&1 | tail -40) > gpurun_out/k_pytest.log 2>&1; tail -3 gpurun_out/k_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 tools/diag_slab_step_host.py > gpurun_out/k_trace_bound.txt 2> gpurun_out/k_trace_bound.err; grep "ms/step" gpurun_out/k_trace_bound.txt; tail -3 gpurun_out/k_trace_bound.err | cut -c1-300
B200SPH_BIND_NUMA=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 tools/diag_slab_step_host.py > gpurun_out/k_trace_unbound.txt 2> gpurun_out/k_trace_unbound.err; grep "ms/step" gpurun_out/k_trace_unbound.txt
