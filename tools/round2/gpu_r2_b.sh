#!/bin/bash
# round 2, GPU run B: brick (staged) pair kernel + new list builder: full GPU test suite, fixtures, A/B benches, ncu captures
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q -rA 2>&1 | grep -v "^PASSED" | tail -150) > gpurun_out/b_pytest_gpu.log 2>&1; tail -5 gpurun_out/b_pytest_gpu.log
(timeout 600 python oracle/gen_golden.py dambreak_dp050_mls10 dambreak_dp050_brezzi dambreak_dp050_planes dambreak_dp050_obstacle 2>&1 | tail -8) > gpurun_out/b_gen_golden.log 2>&1; tail -4 gpurun_out/b_gen_golden.log
for v in bricks gather; do
  br=1; [ $v = gather ] && br=0
  for wl in dambreak2m lattice2m; do
  B200SPH_FORCES_BRICKS=$br timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --quick 2>gpurun_out/b_err_$v.log > gpurun_out/b_${wl}_$v.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_${wl}_$v.json")); print("$wl $v", "ms/step", round(d["ms_per_step"],4), "forces ms", round(d["roofline"]["kernel_ms"],4), "rebuild ms", d["roofline"]["neighbour_rebuild_ms"], "npp", d["config"]["neibs_per_particle"])
except Exception as e: print("$wl $v failed", e); print(open("gpurun_out/b_err_$v.log").read()[-1500:])
PY
  done
done
timeout 900 python bench.py --steps 20 --warmup 10 > gpurun_out/b_ours_8m.json 2> gpurun_out/b_ours_8m.err; tail -c 1200 gpurun_out/b_ours_8m.json; tail -3 gpurun_out/b_ours_8m.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forces_brick -s 6 -c 1 -f -o gpurun_out/prof_brick_r2b python bench.py --workload dambreak2m --steps 3 --warmup 3 --quick > gpurun_out/b_ncu.log 2>&1; tail -2 gpurun_out/b_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_neibs -s 1 -c 1 -f -o gpurun_out/prof_neibs_r2b python bench.py --workload dambreak2m --steps 3 --warmup 3 --quick > gpurun_out/b_ncu2.log 2>&1; tail -2 gpurun_out/b_ncu2.log
timeout 600 ncu --set full --clock-control none -k regex:build_neibs -s 1 -c 1 -f -o gpurun_out/prof_neibs8m_r2b python bench.py --workload dambreak8m --steps 3 --warmup 3 --quick > gpurun_out/b_ncu3.log 2>&1; tail -2 gpurun_out/b_ncu3.log
(timeout 1500 python tools/validate_poiseuille.py --ppH 16 32 64 --iters 1000 --out gpurun_out/b_poiseuille.json 2>&1 | tail -5) > gpurun_out/b_poiseuille.log 2>&1; tail -3 gpurun_out/b_poiseuille.log
