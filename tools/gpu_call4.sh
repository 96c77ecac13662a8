#!/bin/bash
# GPU box session: new host-step tests first, full GPU suite, then e2e A/B (stripe counts) on the bench workloads
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_extras.py -m gpu -q --timeout 300 -k "host" 2>&1 | tail -30) > gpurun_out/pytest_host.log 2>&1; tail -15 gpurun_out/pytest_host.log
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30) > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
run() { tag=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline 2>gpurun_out/err_$tag.log > gpurun_out/b_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_$tag.json")); print("$tag", "ms/step", round(d["ms_per_step"],4), "MIPS", round(d["value"]), "forces ms", round(d["roofline"]["kernel_ms"],4), "rebuild ms", round(d["roofline"]["neighbour_rebuild_ms"],3), "e2e ms", round(d["e2e"]["ms_per_step"],4), "e2e MIPS", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
except Exception as e: print("$tag failed", e); print(open("gpurun_out/err_$tag.log").read()[-1500:])
PY
}
run 2m_s8 --workload dambreak2m
B200SPH_HOST_STRIPES=16 B200SPH_HOST_STRIPE_MIN=100000 run 2m_s16 --workload dambreak2m
B200SPH_HOST_STRIPES=4 run 2m_s4 --workload dambreak2m
run 2m_plain --workload dambreak2m --no-pipeline
run 8m_s8 --workload dambreak8m --steps 10 --warmup 10
B200SPH_HOST_STRIPES=16 run 8m_s16 --workload dambreak8m --steps 10 --warmup 10
