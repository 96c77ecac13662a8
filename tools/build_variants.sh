#!/bin/bash
# Build tuning variants of libb200sph.so that differ only in compile-time knobs of forces.cu / neibs.cu:
#   tools/build_variants.sh name1:"-DFLAG=.. -DFLAG2=.." name2:"..."   ->  build/variants/libb200sph_<name>.so
# (select one at run time with B200SPH_LIB=<path>; build/ travels to the GPU box). Knobs: B200_MIN_BLOCKS, B200_HOIST,
# GATHER_PF, B200_GATHER_AHEAD, B200_LIST_CACHE (forces.cu / pair_physics.cuh); B200_NL_GROUP4, B200_NL_STORE_CS,
# B200_NL_LOAD_EL (neibs.cu).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd); C=$ROOT/gpusph_b200/csrc; O=$ROOT/build/variants; mkdir -p $O/obj
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I$ROOT/include"
for f in api euler filters hoststep; do
  if [ ! -f $O/obj/$f.o ] || [ $C/$f.cu -nt $O/obj/$f.o ] || [ $C/common.cuh -nt $O/obj/$f.o ] || [ $ROOT/include/b200sph.h -nt $O/obj/$f.o ]; then nvcc $FL -c -o $O/obj/$f.o $C/$f.cu & fi
done
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  nvcc $FL $flags -c -o $O/obj/forces_$name.o $C/forces.cu &
  nvcc $FL $flags -c -o $O/obj/neibs_$name.o $C/neibs.cu &
done
wait
for v in "$@"; do
  name=${v%%:*}
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $O/libb200sph_$name.so $O/obj/api.o $O/obj/neibs_$name.o $O/obj/euler.o $O/obj/filters.o $O/obj/hoststep.o $O/obj/forces_$name.o
  echo built $O/libb200sph_$name.so
done
