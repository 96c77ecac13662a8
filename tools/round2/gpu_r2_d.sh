#!/bin/bash
# round 2, GPU run D (2 GPUs): full GPU test suite incl. the 2-GPU bitwise test, 1-GPU benches, 2-GPU strong-scaling bench of both arms
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -300) > gpurun_out/d_pytest_gpu.log 2>&1; tail -4 gpurun_out/d_pytest_gpu.log
for wl in dambreak2m lattice2m; do
  CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --quick 2>gpurun_out/d_err.log > gpurun_out/d_${wl}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/d_${wl}.json")); print("$wl", "ms/step", round(d["ms_per_step"],4), "forces ms", round(d["roofline"]["kernel_ms"],4), "rebuild ms", d["roofline"]["neighbour_rebuild_ms"])
except Exception as e: print("$wl failed", e); print(open("gpurun_out/d_err.log").read()[-1500:])
PY
done
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --steps 20 --warmup 10 > gpurun_out/d_ours_8m.json 2> gpurun_out/d_ours_8m.err; python -c "
import json; d=json.load(open('gpurun_out/d_ours_8m.json')); print('8m ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'rebuild', d['roofline']['neighbour_rebuild_ms'], 'value', d['value'])"; tail -3 gpurun_out/d_ours_8m.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 10 > gpurun_out/d_ours_n2.json 2> gpurun_out/d_ours_n2.err; python -c "
import json; d=json.load(open('gpurun_out/d_ours_n2.json')); print('N=2', d['config']['workload'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'value', d['value'], 'upd/s', d['particle_updates_per_s'])"; tail -5 gpurun_out/d_ours_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 10 > gpurun_out/d_ref_n2.json 2> gpurun_out/d_ref_n2.err; python -c "
import json; d=json.load(open('gpurun_out/d_ref_n2.json')); print('ref N=2', d.get('ms_per_step'), d.get('particle_updates_per_s'), d.get('reference_phase_ms_per_step'), d.get('unavailable'))"; tail -3 gpurun_out/d_ref_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --workload dambreak16m --steps 10 --warmup 10 --quick > gpurun_out/d_ours_16m_n1.json 2> gpurun_out/d_err16.log; python -c "
import json; d=json.load(open('gpurun_out/d_ours_16m_n1.json')); print('16m N=1 ms/step', d['ms_per_step'], 'value', d['value'], 'npp', d['config']['neibs_per_particle'], 'particles', d['config']['particles'])"
