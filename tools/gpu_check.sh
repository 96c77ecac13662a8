#!/bin/bash
# GPU box: parity tests (default gather kernel, then the opt-in staged kernel), then quick benches
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
(B200SPH_FORCES_TILES=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -4)
for wl in ${WORKLOADS:-dambreak2m lattice2m}; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 10 --no-cpu-baseline 2>gpurun_out/bench_err.log > gpurun_out/bench_${wl}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${wl}.json")); print("$wl", "ms/step", round(d["ms_per_step"],3), "MIPS", round(d["value"]), "upd/s", round(d["particle_updates_per_s"]/1e6,1), "forces ms", round(d["roofline"]["kernel_ms"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3))
except Exception as e: print("$wl failed", e); print(open("gpurun_out/bench_err.log").read()[-1500:])
PY
done
