#!/bin/bash
# round 2, GPU run G: unrolled gather-ahead variants
mkdir -p gpurun_out
for v in base un7 un8 un6; do
  lib=$PWD/build/variants/libb200sph_$v.so; [ $v = base ] && lib=$PWD/gpusph_b200/libb200sph.so
  for wl in dambreak2m dambreak8m lattice2m; do
  B200SPH_LIB=$lib timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --quick 2>gpurun_out/g_err_$v.log > gpurun_out/g_${wl}_$v.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/g_${wl}_$v.json")); print("$wl $v", "ms/step", round(d["ms_per_step"],4), "forces ms", round(d["roofline"]["kernel_ms"],4))
except Exception as e: print("$wl $v failed", e); print(open("gpurun_out/g_err_$v.log").read()[-1500:])
PY
  done
done
B200SPH_LIB=$PWD/build/variants/libb200sph_un7.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q 2>&1 | tail -3
