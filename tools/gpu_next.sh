#!/bin/bash
# First GPU session of the next round: what was written at the end of round 1 without GPU time left.
#   here:  tools/build_variants.sh g4:"-DB200_NL_GROUP4=1" zc:"-DB200_HOST_ZEROCOPY=1" && cp gpusph_b200/libb200sph.so build/variants/libb200sph_base.so
#   then:  gpurun --timeout 600 -- tools/gpu_next.sh
mkdir -p gpurun_out
# 1. HotFile relay back into the reference at a rebuild iteration (expected: XPASS -> drop the xfail marker)
(timeout 300 python -m pytest tests/test_zz_hotfile_gpu.py -m gpu -q -rxX 2>&1 | tail -6)
# 2. list builder with four-at-a-time candidate tests: bit-exact list, then rebuild time at 2 M and 8 M
(B200SPH_LIB=$PWD/build/variants/libb200sph_g4.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -2)
for wl in dambreak2m dambreak8m; do for v in base g4; do
  B200SPH_LIB=$PWD/build/variants/libb200sph_$v.so timeout 300 python bench.py --workload $wl --steps 10 --warmup 10 --no-cpu-baseline 2>gpurun_out/sweep_err.log > gpurun_out/sweep_${wl}_$v.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/sweep_${wl}_$v.json")); print("$wl $v", "ms/step", round(d["ms_per_step"],4), "rebuild ms", round(d["roofline"]["neighbour_rebuild_ms"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3))
except Exception as e: print("$wl $v failed", e); print(open("gpurun_out/sweep_err.log").read()[-800:])
PY
done; done
# 3. zero-copy downloads in b200sph_step_host (the corrector's epilogue stores state n+1 straight into the mapped host buffers)
(B200SPH_LIB=$PWD/build/variants/libb200sph_zc.so B200SPH_HOST_ZEROCOPY=1 timeout 600 python -m pytest tests/test_gpu_extras.py -m gpu -q -k "host" 2>&1 | tail -2)
for wl in dambreak2m dambreak8m; do for z in 0 1; do
  B200SPH_LIB=$PWD/build/variants/libb200sph_zc.so B200SPH_HOST_ZEROCOPY=$z timeout 300 python bench.py --workload $wl --steps 10 --warmup 10 --no-cpu-baseline 2>gpurun_out/sweep_err.log > gpurun_out/sweep_${wl}_zc$z.json
  python -c "import json; d=json.load(open('gpurun_out/sweep_${wl}_zc$z.json')); print('$wl zerocopy=$z ms/step', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))" || tail -5 gpurun_out/sweep_err.log
done; done
