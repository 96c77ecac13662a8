#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30) > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
run() { tag=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline 2>gpurun_out/err_$tag.log > gpurun_out/b_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_$tag.json")); print("$tag", "ms/step", round(d["ms_per_step"],4), "MIPS", round(d["value"]), "forces ms", round(d["roofline"]["kernel_ms"],4), "rebuild ms", round(d["roofline"]["neighbour_rebuild_ms"],3), "e2e ms", round(d["e2e"]["ms_per_step"],4), "e2e MIPS", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
except Exception as e: print("$tag failed", e); print(open("gpurun_out/err_$tag.log").read()[-1500:])
PY
}
run 2m_fused --workload dambreak2m
B200SPH_FUSED_EULER=0 run 2m_unfused --workload dambreak2m
run 8m_fused --workload dambreak8m --steps 10 --warmup 10
run 84k_fused --workload dambreak84k --steps 100 --warmup 20
