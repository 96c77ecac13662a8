#!/bin/bash
# round 2, GPU run E: cell-table variants of the pair kernel, ncu evidence for the current default
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5) > gpurun_out/e_pytest.log 2>&1; tail -2 gpurun_out/e_pytest.log
for v in base ct ct8 ct6; do
  lib=$PWD/build/variants/libb200sph_$v.so; [ $v = base ] && lib=$PWD/gpusph_b200/libb200sph.so
  for wl in dambreak2m dambreak8m; do
  B200SPH_LIB=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --quick 2>gpurun_out/e_err_$v.log > gpurun_out/e_${wl}_$v.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/e_${wl}_$v.json")); print("$wl $v", "ms/step", round(d["ms_per_step"],4), "forces ms", round(d["roofline"]["kernel_ms"],4), "rebuild ms", d["roofline"]["neighbour_rebuild_ms"])
except Exception as e: print("$wl $v failed", e); print(open("gpurun_out/e_err_$v.log").read()[-1500:])
PY
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forces_gather -s 6 -c 1 -f -o gpurun_out/prof_gather_r2e python bench.py --workload dambreak2m --steps 3 --warmup 3 --quick > gpurun_out/e_ncu.log 2>&1; tail -1 gpurun_out/e_ncu.log
B200SPH_LIB=$PWD/build/variants/libb200sph_ct.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:forces_gather -s 6 -c 1 -f -o gpurun_out/prof_gather_ct_r2e python bench.py --workload dambreak2m --steps 3 --warmup 3 --quick > gpurun_out/e_ncu2.log 2>&1; tail -1 gpurun_out/e_ncu2.log
