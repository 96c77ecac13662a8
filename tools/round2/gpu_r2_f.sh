#!/bin/bash
# round 2, GPU run F: e2e (host-buffer step) tuning at 8 M: stripes, zero-copy epilogue
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline 2>gpurun_out/f_err_$name.log > gpurun_out/f_$name.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/f_$name.json")); print("$name", "ms/step", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "GB/s per dir", round(d["e2e"]["h2d_bytes_per_step"]/d["e2e"]["ms_per_step"]/1e6,1))
except Exception as e: print("$name failed", e); print(open("gpurun_out/f_err_$name.log").read()[-800:])
PY
}
run s8 B200SPH_HOST_STRIPES=8
run s16 B200SPH_HOST_STRIPES=16
run s24 B200SPH_HOST_STRIPES=24
run s32 B200SPH_HOST_STRIPES=32
run zc8 B200SPH_LIB=$PWD/build/variants/libb200sph_zc.so B200SPH_HOST_ZEROCOPY=1 B200SPH_HOST_STRIPES=8
run zc16 B200SPH_LIB=$PWD/build/variants/libb200sph_zc.so B200SPH_HOST_ZEROCOPY=1 B200SPH_HOST_STRIPES=16
run l1s16 B200SPH_HOST_LANES=1 B200SPH_HOST_STRIPES=16
B200SPH_HOST_TRACE=1 B200SPH_HOST_STRIPES=16 timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > /dev/null 2> gpurun_out/f_trace_s16.log; grep -A20 "host trace" gpurun_out/f_trace_s16.log | tail -22
