#!/bin/bash
# GPU box: parity tests with the default kernel selection, then A/B of env knobs on the bench workloads
# usage: tools/gpu_ab.sh "ENV1=.. ENV2=..|ENV1=..|..." "dambreak2m lattice2m"
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
IFS='|' read -ra CFGS <<< "${1:-}"
for wl in ${2:-dambreak2m}; do for cfg in "${CFGS[@]}"; do
  tag=$(echo "$cfg" | tr ' =/' '___')
  env $cfg timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/ab_err.log > gpurun_out/ab_${wl}_${tag}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_${wl}_${tag}.json")); print("$wl [$cfg]", "ms/step", round(d["ms_per_step"],4), "forces ms", round(d["roofline"]["kernel_ms"],4), "MIPS", round(d["value"]))
except Exception as e: print("$wl [$cfg] failed", e); print(open("gpurun_out/ab_err.log").read()[-1200:])
PY
done; done
