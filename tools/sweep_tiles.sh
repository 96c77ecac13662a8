#!/bin/bash
# forces kernel: gather fallback vs staged tile configurations
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -5)
for wl in dambreak2m lattice2m; do
for cfg in g 0 1 2 3 4 5; do
  if [ $cfg = g ]; then export B200SPH_FORCES_TILES=0; else export B200SPH_FORCES_TILES=1 B200SPH_TILE_CFG=$cfg; fi
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_err.log > /tmp/o.json
  python - <<PY
import json
try:
    d=json.load(open("/tmp/o.json")); print("$wl cfg=$cfg", "ms/step", round(d["ms_per_step"],3), "forces ms", round(d["roofline"]["kernel_ms"],3))
except Exception as e: print("$wl cfg=$cfg failed", e); print(open("gpurun_out/bench_err.log").read()[-800:])
PY
done; done
