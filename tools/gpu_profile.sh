#!/bin/bash
# profiles for profiles/: launch list of the default bench command + full captures of the two heaviest kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 11 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"build_neibs_kernel" -c 1 -f -o gpurun_out/prof_buildneibs python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
