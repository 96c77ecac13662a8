#!/bin/bash
# round 2, GPU run T (1 GPU): does the distance between the rows of the neighbour list matter? dambreak2m with the list allocated for 1x / 4x / 8x the particles
mkdir -p gpurun_out
for S in 1 4 8; do
B200SPH_TEST_ALLOC_SCALE=$S timeout 300 python bench.py --workload dambreak2m --quick --steps 20 --warmup 10 > gpurun_out/t_$S.json 2> gpurun_out/t_$S.err; python -c "
import json; d=json.load(open('gpurun_out/t_$S.json')); print('alloc x$S ms/step', round(d['ms_per_step'],4), 'kernel ms', round(d['roofline']['kernel_ms'],4), 'rebuild', round(d['roofline']['neighbour_rebuild_ms'],3))"; tail -2 gpurun_out/t_$S.err | cut -c1-200
done
