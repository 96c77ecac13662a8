#!/usr/bin/env bash
# TEST / BASELINE INFRASTRUCTURE ONLY.
#
# Builds the UNMODIFIED GPUSPH reference (its own engines, kernels, launch
# configurations and summation order) for sm_100 from the sources where they
# lie under $REF (default /root/reference), into oracle/_ref/.
#
# Nothing from the reference is copied into this repository: sources are
# staged into a scratch directory under ${TMPDIR:-/tmp}, two mechanical
# edits are applied there, and only the linked binaries land in oracle/_ref/
# (git-ignored). We do NOT run the reference's Makefile: every compile
# command is issued from this script.
#
# Why edits are needed at all (finding 1):
#  * CUDA 12 removed *texture references* (texture<T,1,...> + cudaBindTexture +
#    tex1Dfetch). The reference uses them everywhere (src/cuda/textures.cuh:63-89).
#    We add a ~20 line shim (texref_shim.h, authored here) that re-creates the
#    texture<> template as a {pointer} struct in __device__ memory, maps
#    tex1Dfetch -> __ldg (value-identical for point-sampled, unnormalised,
#    element-type reads, which is the only mode the reference uses) and
#    cudaBindTexture -> cudaMemcpyToSymbol. Every kernel, launch configuration
#    and floating point summation order of the reference is untouched.
#  * gcc 13 needs <cstdint>/<climits>/<cstring> force-included.
#
# Generated files the reference Makefile would produce (options/*.opt,
# <Problem>.gen.cc, parse/describe-debugflags.h) are produced here with the
# same one-line contents / the reference's own awk scripts.
#
# Usage: oracle/build_ref.sh [Problem ...]     (default: DamBreak3D)
#        LINEARIZATION=xzy REF_SUFFIX=_xzy oracle/build_ref.sh DamBreak3D
#            the same unmodified sources with another cell linearisation (the reference's own `linearization=` build
#            option, Makefile + src/linearization.h) -> oracle/_ref/DamBreak3D_xzy: used by bench.py's reference arm on
#            N > 1 GPUs, where DamBreak3D's Y split needs Y to be the slowest hash digit for contiguous halo bursts
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
REF_SUFFIX="${REF_SUFFIX:-}"
WORK="${TMPDIR:-/tmp}/gpusph_b200_refbuild${REF_SUFFIX}"
PROBLEMS=("$@"); [ ${#PROBLEMS[@]} -eq 0 ] && PROBLEMS=(DamBreak3D)
JOBS="${JOBS:-$(nproc)}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
CXX="${CXX:-g++}"
ARCH="${REF_ARCH:-sm_100}"

[ -d "$REF/src" ] || { echo "build_ref: $REF/src not found (reference absent) - nothing to do"; exit 0; }
mkdir -p "$OUT" "$WORK"
rsync -a --delete "$REF/src/" "$WORK/src/" 2>/dev/null || { rm -rf "$WORK/src"; cp -r "$REF/src" "$WORK/src"; }
chmod -R u+w "$WORK/src"
mkdir -p "$WORK/options" "$WORK/build"

# ---- texture-reference shim (our code, not the reference's) ----
cat > "$WORK/src/cuda/texref_shim.h" <<'SHIM'
#pragma once
#include <cuda_runtime.h>
// Emulation of the texture-reference API removed in CUDA 12 (point sampling only).
template<typename T, int dim = 1, cudaTextureReadMode mode = cudaReadModeElementType>
struct texture {
	const T* ptr; cudaTextureObject_t obj;
	cudaTextureAddressMode addressMode[3]; cudaTextureFilterMode filterMode; bool normalized;
};
template<typename T, int d, cudaTextureReadMode m>
__device__ __forceinline__ T tex1Dfetch(texture<T,d,m> const& t, int i) { return __ldg(t.ptr + i); }
template<typename T, cudaTextureReadMode m>
__device__ __forceinline__ T tex2D(texture<T,2,m> const& t, float x, float y) { return tex2D<T>(t.obj, x, y); }
template<typename T, int d, cudaTextureReadMode m>
inline cudaError_t cudaBindTexture(size_t*, texture<T,d,m>& t, const void* p, size_t = 0) {
	texture<T,d,m> h{}; h.ptr = (const T*)p; return cudaMemcpyToSymbol(t, &h, sizeof(h)); }
template<typename T, int d, cudaTextureReadMode m>
inline cudaError_t cudaUnbindTexture(texture<T,d,m>&) { return cudaSuccess; }
template<typename T, int d, cudaTextureReadMode m>
inline cudaError_t cudaBindTextureToArray(texture<T,d,m>&, cudaArray_t, cudaChannelFormatDesc const&) { return cudaErrorNotSupported; }
SHIM
# file-scope texture<> declarations must live in device memory now
sed -i -E 's/^texture</__device__ texture</' "$WORK/src/cuda/textures.cuh" "$WORK/src/cuda/geom_core.cu"
# pull the shim in before the first declaration
sed -i '0,/^__device__ texture</s//#include "texref_shim.h"\n__device__ texture</' "$WORK/src/cuda/textures.cuh"

# ---- generated option headers (same content the reference Makefile writes) ----
o="$WORK/options"
echo '#define USE_CATALYST 0'  > $o/catalyst_select.opt
echo '#define USE_CHRONO 0'    > $o/chrono_select.opt
echo "#define COMPUTE ${ARCH#sm_}" > $o/compute_select.opt
echo '#undef _DEBUG_'          > $o/dbg_select.opt
echo '#define FASTMATH 0'      > $o/fastmath_select.opt
echo '#define GIT_INFO_OUTPUT ""' > $o/git_info.opt
echo '#define GPUSPH_VERSION "reference-shim-build"' > $o/gpusph_version.opt
echo '#define USE_HDF5 0'      > $o/hdf5_select.opt
echo '#define USE_MPI 0'       > $o/mpi_select.opt
echo '#define MAKE_SHOW_OUTPUT "built by oracle/build_ref.sh\n"' > $o/make_show.opt
LIN="${LINEARIZATION:-yzx}"
printf '#define LINEARIZATION "%s"\n#define COORD1 %s\n#define COORD2 %s\n#define COORD3 %s\n' \
	"$LIN" "${LIN:0:1}" "${LIN:1:1}" "${LIN:2:1}" > $o/linearization_select.opt
awk -f "$REF/scripts/parse-debugflags.awk"    "$REF/src/debugflags.def" > "$WORK/src/parse-debugflags.h"
awk -f "$REF/scripts/describe-debugflags.awk" "$REF/src/debugflags.def" > "$WORK/src/describe-debugflags.h"

cd "$WORK"
INC="-Isrc -Isrc/adaptors -Isrc/cuda -Isrc/geometries -Isrc/integrators -Isrc/problem_api -Isrc/problems -Isrc/writers -Isrc/problems/user -Ioptions"
CPPFLAGS="-include cstdint -include climits -include cstring $INC -D__STDC_CONSTANT_MACROS -D__STDC_LIMIT_MACROS -D_GLIBCXX_USE_C99_MATH -DUSE_HDF5=0 -D__COMPUTE__=${ARCH#sm_}"
CXXFLAGS="-m64 -std=c++11 -O3 -w"
CUFLAGS="-arch=$ARCH --generate-line-info -std=c++11 --compiler-options -m64,-O3,-w -w"
CUDA_INC="-I/usr/local/cuda/include"

# host objects (shared by every problem)
# DisplayWriter needs VTK/Catalyst (the reference Makefile drops it when catalyst=0)
CCS=$(cd src && ls *.cc cuda/*.cc geometries/*.cc integrators/*.cc problem_api/*.cc writers/*.cc | grep -v DisplayWriter)
{
  objs=""
  for c in $CCS; do objs="$objs build/${c%.cc}.o"; done
  echo "OBJS=$objs"
  echo "all: \$(OBJS)"
  for c in $CCS; do
    ob="build/${c%.cc}.o"
    printf '%s: src/%s\n\t@mkdir -p $(dir $@)\n\t@echo CC %s\n\t@%s %s %s %s -c -o $@ $<\n' "$ob" "$c" "$c" "$CXX" "$CPPFLAGS" "$CUDA_INC" "$CXXFLAGS"
  done
} > build/host.mk
make -s -j"$JOBS" -f build/host.mk all
HOSTOBJS=$(for c in $CCS; do echo "build/${c%.cc}.o"; done)

for P in "${PROBLEMS[@]}"; do
  echo "== building reference problem $P ($ARCH)"
  sed -e "s/PROBLEM/$P/g" src/problem_gen.tpl > options/$P.gen.cc
  $CXX $CPPFLAGS $CUDA_INC $CXXFLAGS -c -o build/$P.gen.o options/$P.gen.cc &
  $NVCC $CPPFLAGS $CUFLAGS -c -o build/$P.o src/problems/$P.cu
  wait
  $NVCC -arch=$ARCH -o "$OUT/$P$REF_SUFFIX" $HOSTOBJS build/$P.gen.o build/$P.o -lpthread -lrt
  echo "   -> $OUT/$P$REF_SUFFIX"
done
