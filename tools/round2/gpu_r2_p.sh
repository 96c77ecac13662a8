#!/bin/bash
# round 2, GPU run P (1 GPU): ncu of the trimmed pair kernel (r1p1: 96-instruction loop) and of the slow combination
mkdir -p gpurun_out
for V in r1p1 default; do
  if [ $V = default ]; then unset B200SPH_LIB; else export B200SPH_LIB=$PWD/build/variants/libb200sph_$V.so; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:forces_gather -s 6 -c 1 -f -o gpurun_out/prof_pair_$V python bench.py --workload dambreak2m --steps 3 --warmup 3 --quick > gpurun_out/p_ncu_$V.log 2>&1; tail -1 gpurun_out/p_ncu_$V.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep | tail -3
