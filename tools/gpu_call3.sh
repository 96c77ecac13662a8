#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_extras.py -m gpu -q --timeout 300 2>&1 | tail -30) > gpurun_out/pytest_gpu_extras.log 2>&1; tail -6 gpurun_out/pytest_gpu_extras.log
run() { tag=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline 2>gpurun_out/err_$tag.log > gpurun_out/b_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_$tag.json")); print("$tag", "ms/step", round(d["ms_per_step"],4), "MIPS", round(d["value"]), "forces ms", round(d["roofline"]["kernel_ms"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4), "e2e MIPS", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
except Exception as e: print("$tag failed", e); print(open("gpurun_out/err_$tag.log").read()[-1500:])
PY
}
run 2m_pipe --workload dambreak2m
run 2m_plain --workload dambreak2m --no-pipeline
run 84k --workload dambreak84k --steps 100 --warmup 20 --graphs
run 8m_pipe --workload dambreak8m --steps 10 --warmup 10
