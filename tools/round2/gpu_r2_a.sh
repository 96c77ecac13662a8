#!/bin/bash
# round 2, GPU run A: full GPU test suite (incl. the stock-vs-drop-in binaries), new reference fixtures, kernel variants,
# the 8 M default bench with its reference arm, one full ncu capture of the pair kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
(timeout 2400 python -m pytest tests -m gpu -q -x -rA 2>&1 | tail -80) > gpurun_out/a_pytest_gpu.log 2>&1; tail -5 gpurun_out/a_pytest_gpu.log
(timeout 600 python oracle/gen_golden.py dambreak_dp050_mls10 dambreak_dp050_brezzi dambreak_dp050_planes dambreak_dp050_obstacle 2>&1 | tail -8) > gpurun_out/a_gen_golden.log 2>&1; tail -4 gpurun_out/a_gen_golden.log
for v in base ahead8 ahead7 ahead6 mb7 pf1 pf3; do
  lib=$PWD/build/variants/libb200sph_$v.so; [ $v = base ] && lib=$PWD/gpusph_b200/libb200sph.so
  B200SPH_LIB=$lib timeout 300 python bench.py --workload dambreak2m --steps 20 --warmup 5 --quick 2>gpurun_out/a_err_$v.log > gpurun_out/a_2m_$v.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/a_2m_$v.json")); print("$v", "ms/step", round(d["ms_per_step"],4), "forces ms", round(d["roofline"]["kernel_ms"],4), "rebuild ms", d["roofline"]["neighbour_rebuild_ms"], "npp", d["config"]["neibs_per_particle"])
except Exception as e: print("$v failed", e); print(open("gpurun_out/a_err_$v.log").read()[-1200:])
PY
done
timeout 900 python bench.py --impl reference --steps 20 --warmup 10 > gpurun_out/a_ref_8m.json 2> gpurun_out/a_ref_8m.err; tail -c 600 gpurun_out/a_ref_8m.json
timeout 900 python bench.py --steps 20 --warmup 10 > gpurun_out/a_ours_8m.json 2> gpurun_out/a_ours_8m.err; tail -c 1500 gpurun_out/a_ours_8m.json; tail -3 gpurun_out/a_ours_8m.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forces_gather -s 6 -c 1 -f -o gpurun_out/prof_forces_r2a python bench.py --workload dambreak2m --steps 3 --warmup 3 --quick > gpurun_out/a_ncu.log 2>&1; tail -2 gpurun_out/a_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/a_launches_2m.csv python bench.py --workload dambreak2m --steps 12 --warmup 10 --quick > gpurun_out/a_ncu2.log 2>&1; tail -2 gpurun_out/a_ncu2.log
