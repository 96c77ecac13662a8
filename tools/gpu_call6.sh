#!/bin/bash
mkdir -p gpurun_out
for L in 1 2; do (B200SPH_HOST_LANES=$L timeout 600 python -m pytest tests/test_gpu_extras.py -m gpu -q --timeout 300 -k "host" 2>&1 | tail -3); done
B200SPH_HOST_LANES=2 python tools/diag_step_host.py dambreak2m 2>&1 | grep -v "GB/s" | head -24
run() { tag=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline 2>gpurun_out/err_$tag.log > gpurun_out/b_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_$tag.json")); print("$tag", "ms/step", round(d["ms_per_step"],4), "MIPS", round(d["value"]), "forces ms", round(d["roofline"]["kernel_ms"],4), "rebuild ms", round(d["roofline"]["neighbour_rebuild_ms"],3), "e2e ms", round(d["e2e"]["ms_per_step"],4), "e2e MIPS", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
except Exception as e: print("$tag failed", e); print(open("gpurun_out/err_$tag.log").read()[-1500:])
PY
}
B200SPH_HOST_LANES=2 run 2m_l2 --workload dambreak2m
B200SPH_HOST_LANES=1 run 2m_l1 --workload dambreak2m
B200SPH_HOST_LANES=2 B200SPH_HOST_STRIPES=12 B200SPH_HOST_STRIPE_MIN=100000 run 2m_l2s12 --workload dambreak2m
B200SPH_HOST_LANES=2 B200SPH_HOST_STRIPES=5 run 2m_l2s5 --workload dambreak2m
B200SPH_HOST_LANES=2 run 8m_l2 --workload dambreak8m --steps 10 --warmup 10
B200SPH_HOST_LANES=2 B200SPH_HOST_STRIPES=16 run 8m_l2s16 --workload dambreak8m --steps 10 --warmup 10
