#!/bin/bash
# GPU box: time tuning variants of the forces kernel (tools/build_variants.sh) on the bench workloads
# usage: tools/sweep_variants.sh "base h6 h7" "0 1" "dambreak2m lattice2m"
VARS=${1:-base}; RECS=${2:-0}; WLS=${3:-dambreak2m}
for wl in $WLS; do for v in $VARS; do for r in $RECS; do
  B200SPH_LIB=$PWD/build/variants/libb200sph_$v.so B200SPH_FORCES_COOP=$r timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/sweep_err.log > gpurun_out/sweep_${wl}_${v}_coop$r.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_${wl}_${v}_coop$r.json")); print("$wl $v coop=$r", "ms/step", round(d["ms_per_step"],4), "forces ms", round(d["roofline"]["kernel_ms"],4))
except Exception as e: print("$wl $v coop=$r failed", e); print(open("gpurun_out/sweep_err.log").read()[-800:])
PY
done; done; done
if [ -n "$PARITY" ]; then
  v=${PARITY%%:*}; r=${PARITY#*:}
  B200SPH_LIB=$PWD/build/variants/libb200sph_$v.so B200SPH_FORCES_COOP=$r timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -5
fi
