#!/usr/bin/env bash
# Drop-in demonstration: link the REFERENCE's host code (GPUSPH orchestrator, integrator, GPUWorker, problem API,
# writers) and its stock framework (visc / BC engines, other post-processing) with OUR engines: neibs, forces, integration,
# the SHEPARD / MLS filters and the TESTPOINTS post-process
# (gpusph_b200/host/b200_engines.h -> libb200sph.so). The only reference-side change is the 3-line patch of
# GPUWorker's constructor shown in INTEGRATION.md; the problem file (DamBreak3D.cu) is compiled unchanged.
# Needs the scratch tree and objects produced by oracle/build_ref.sh. Output: build/dropin/<Problem>_b200
# (git-ignored, travels to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
WORK="${TMPDIR:-/tmp}/gpusph_b200_refbuild"
P="${1:-DamBreak3D}"
[ -d "$WORK/build" ] || "$HERE/oracle/build_ref.sh" "$P"
OUT="$HERE/build/dropin"; mkdir -p "$OUT"
cd "$WORK"
cp src/GPUWorker.cc src/GPUWorker_b200.cc
# --- the maintainer's patch ---
sed -i 's|^using namespace std;|#include "b200_engines.h"\nstatic std::shared_ptr<b200::Contexts> b200_contexts() { static std::shared_ptr<b200::Contexts> c = std::make_shared<b200::Contexts>(); return c; }\nusing namespace std;|' src/GPUWorker_b200.cc
sed -i 's|neibsEngine(gdata->simframework->getNeibsEngine()),|neibsEngine(new b200::NeibsEngine(b200_contexts(), gdata->simframework->getNeibsEngine())),|' src/GPUWorker_b200.cc
sed -i 's|forcesEngine(gdata->simframework->getForcesEngine()),|forcesEngine(new b200::ForcesEngine(b200_contexts(), gdata->simframework->getForcesEngine())),|' src/GPUWorker_b200.cc
sed -i 's|integrationEngine(gdata->simframework->getIntegrationEngine()),|integrationEngine(new b200::IntegrationEngine(b200_contexts(), gdata->simframework->getIntegrationEngine())),|' src/GPUWorker_b200.cc
sed -i 's|filterEngines(gdata->simframework->getFilterEngines()),|filterEngines(b200::filters(b200_contexts(), gdata->simframework->getFilterEngines())),|' src/GPUWorker_b200.cc
sed -i 's|postProcEngines(gdata->simframework->getPostProcEngines()),|postProcEngines(b200::postprocess(b200_contexts(), gdata->simframework->getPostProcEngines())),|' src/GPUWorker_b200.cc
grep -c "b200::" src/GPUWorker_b200.cc
INC="-Isrc -Isrc/adaptors -Isrc/cuda -Isrc/geometries -Isrc/integrators -Isrc/problem_api -Isrc/problems -Isrc/writers -Isrc/problems/user -Ioptions"
g++ -include cstdint -include climits -include cstring $INC -I/usr/local/cuda/include -I"$HERE/include" -I"$HERE/gpusph_b200/host" \
    -D__STDC_CONSTANT_MACROS -D__STDC_LIMIT_MACROS -D_GLIBCXX_USE_C99_MATH -DUSE_HDF5=0 -D__COMPUTE__=100 \
    -m64 -std=c++11 -O3 -w -c -o build/GPUWorker_b200.o src/GPUWorker_b200.cc
OBJS=$(find build -name '*.o' ! -name 'GPUWorker.o' ! -name 'GPUWorker_b200.o' ! -name '*.gen.o' ! -path 'build/problems/*' ! -name 'DamBreak3D.o' ! -name 'Poiseuille.o' | tr '\n' ' ')
/usr/local/cuda/bin/nvcc -arch=sm_100 -o "$OUT/${P}_b200" $OBJS build/GPUWorker_b200.o build/$P.gen.o build/$P.o \
    -L"$HERE/gpusph_b200" -lb200sph -Xlinker -rpath -Xlinker '$ORIGIN/../../gpusph_b200' -lpthread -lrt
echo "-> $OUT/${P}_b200"
