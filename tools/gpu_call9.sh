#!/bin/bash
mkdir -p gpurun_out
(B200SPH_HOST_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_extras.py -m gpu -q --timeout 300 -k "host" 2>&1 | tail -3)
run() { tag=$1; shift; timeout 300 python bench.py "$@" --no-cpu-baseline 2>gpurun_out/err_$tag.log > gpurun_out/b_$tag.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/b_$tag.json")); print("$tag", "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_step"],4), "e2e MIPS", round(d["e2e"]["value"]))
except Exception as e: print("$tag failed", e); print(open("gpurun_out/err_$tag.log").read()[-1500:])
PY
}
B200SPH_HOST_SPLIT=0 run 2m_split0 --workload dambreak2m
B200SPH_HOST_SPLIT=1 run 2m_split1 --workload dambreak2m
B200SPH_HOST_SPLIT=1 run 8m_split1 --workload dambreak8m --steps 10 --warmup 10
B200SPH_HOST_SPLIT=1 B200SPH_HOST_TRACE=1 python tools/diag_step_host.py dambreak2m 2>&1 | grep -v "GB/s" | head -11
