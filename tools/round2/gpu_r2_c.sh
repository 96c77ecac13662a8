#!/bin/bash
# round 2, GPU run C: locality-scheduled (sweep) pair kernel: full GPU test suite, A/B benches, ncu
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -400) > gpurun_out/c_pytest_gpu.log 2>&1; tail -5 gpurun_out/c_pytest_gpu.log
for v in sweep gather; do
  br=1; [ $v = gather ] && br=0
  for wl in dambreak2m lattice2m; do
  B200SPH_FORCES_SWEEP=$br timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --quick 2>gpurun_out/c_err_$v.log > gpurun_out/c_${wl}_$v.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c_${wl}_$v.json")); print("$wl $v", "ms/step", round(d["ms_per_step"],4), "forces ms", round(d["roofline"]["kernel_ms"],4), "rebuild ms", d["roofline"]["neighbour_rebuild_ms"], "npp", d["config"]["neibs_per_particle"])
except Exception as e: print("$wl $v failed", e); print(open("gpurun_out/c_err_$v.log").read()[-1500:])
PY
  done
done
timeout 900 python bench.py --steps 20 --warmup 10 > gpurun_out/c_ours_8m.json 2> gpurun_out/c_ours_8m.err; python -c "
import json; d=json.load(open('gpurun_out/c_ours_8m.json')); print('8m ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'rebuild', d['roofline']['neighbour_rebuild_ms'])"; tail -3 gpurun_out/c_ours_8m.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forces_sweep -s 6 -c 1 -f -o gpurun_out/prof_sweep_r2c python bench.py --workload dambreak2m --steps 3 --warmup 3 --quick > gpurun_out/c_ncu.log 2>&1; tail -2 gpurun_out/c_ncu.log
