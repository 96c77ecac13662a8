#!/bin/bash
# GPU box session: full GPU parity suite (all failures listed), then A/B of the cache-policy variants of the forces kernel
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -60) > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
tools/sweep_variants.sh "${VARS:-base na ef nael el}" "0" "${WLS:-dambreak2m}"
