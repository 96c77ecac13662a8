#!/bin/bash
# A/B of the list builder's L2 policies at 8 M and 2 M (rebuild time from bench.py's roofline block)
for wl in dambreak8m dambreak2m; do for v in base nlcs nlel nlcsel; do
  B200SPH_LIB=$PWD/build/variants/libb200sph_$v.so timeout 300 python bench.py --workload $wl --steps 10 --warmup 10 --no-cpu-baseline 2>gpurun_out/sweep_err.log > gpurun_out/sweep_${wl}_$v.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_${wl}_$v.json")); print("$wl $v", "ms/step", round(d["ms_per_step"],4), "forces ms", round(d["roofline"]["kernel_ms"],4), "rebuild ms", round(d["roofline"]["neighbour_rebuild_ms"],3))
except Exception as e: print("$wl $v failed", e); print(open("gpurun_out/sweep_err.log").read()[-800:])
PY
done; done
B200SPH_LIB=$PWD/build/variants/libb200sph_nlcsel.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -2
