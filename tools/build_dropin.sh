#!/usr/bin/env bash
# Drop-in build: the REFERENCE's unmodified host code (GPUSPH orchestrator, integrator, GPUWorker, problem API, writers)
# and its unmodified problem files, with OUR framework seam: gpusph_b200/host/cudasimframework.cu is found before
# src/cuda/cudasimframework.cu on the include path, so `#include "cudasimframework.cu"` in the problem file
# (src/problems/DamBreak3D.cu:35, src/problems/Poiseuille.inc:50) yields the B200 engines (-> libb200sph.so).
# No reference file is edited and none of the reference's CUDA engines is compiled or linked.
# Needs the host objects produced by oracle/build_ref.sh (the texture shim that script applies only touches src/cuda
# files, which this build does not use). Output: build/dropin/<Problem>_b200 (git-ignored, travels to the GPU box).
#
# Usage: tools/build_dropin.sh [Problem ...]      (default: DamBreak3D Poiseuille)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
WORK="${TMPDIR:-/tmp}/gpusph_b200_refbuild"
PROBLEMS=("$@"); [ ${#PROBLEMS[@]} -eq 0 ] && PROBLEMS=(DamBreak3D Poiseuille)
[ -f "$WORK/build/GPUWorker.o" ] || "$HERE/oracle/build_ref.sh" DamBreak3D
OUT="$HERE/build/dropin"; mkdir -p "$OUT" "$WORK/build_b200"
cd "$WORK"
# our host directory first: that is the whole integration
INC="-I$HERE/gpusph_b200/host -I$HERE/include -Isrc -Isrc/adaptors -Isrc/cuda -Isrc/geometries -Isrc/integrators -Isrc/problem_api -Isrc/problems -Isrc/writers -Isrc/problems/user -Ioptions"
CPPFLAGS="-include cstdint -include climits -include cstring $INC -D__STDC_CONSTANT_MACROS -D__STDC_LIMIT_MACROS -D_GLIBCXX_USE_C99_MATH -DUSE_HDF5=0 -D__COMPUTE__=100"
HOSTOBJS=$(find build -name '*.o' ! -name '*.gen.o' ! -path 'build/problems/*' $(for p in build/*.gen.o; do b=$(basename "$p" .gen.o); echo "! -name $b.o"; done) | tr '\n' ' ')
for P in "${PROBLEMS[@]}"; do
  echo "== drop-in build of $P"
  [ -f "build/$P.gen.o" ] || { sed -e "s/PROBLEM/$P/g" src/problem_gen.tpl > "options/$P.gen.cc"; \
    g++ $CPPFLAGS -I/usr/local/cuda/include -m64 -std=c++11 -O3 -w -c -o "build/$P.gen.o" "options/$P.gen.cc"; }
  /usr/local/cuda/bin/nvcc $CPPFLAGS -arch=sm_100 -std=c++11 --compiler-options -m64,-O3,-w -w -c -o "build_b200/$P.o" "src/problems/$P.cu"
  /usr/local/cuda/bin/nvcc -arch=sm_100 -o "$OUT/${P}_b200" $HOSTOBJS "build/$P.gen.o" "build_b200/$P.o" \
      -L"$HERE/gpusph_b200" -lb200sph -Xlinker -rpath -Xlinker '$ORIGIN/../../gpusph_b200' -lpthread -lrt
  echo "-> $OUT/${P}_b200"
done
