#!/bin/bash
# round 2, GPU run FF (1 GPU): the benchmark-size list-layout test alone
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "benchmark_size or blocked" 2>&1 | tail -15) > gpurun_out/ff_pytest.log 2>&1; tail -3 gpurun_out/ff_pytest.log
