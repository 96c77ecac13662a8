#!/bin/bash
# what the driver does at round end, in one session: gpu tests, smoke(), default bench (ours + reference arm)
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --impl reference > gpurun_out/BENCH_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/BENCH_ref.json; echo
timeout 600 python bench.py > gpurun_out/BENCH.json 2> gpurun_out/bench.err; cat gpurun_out/BENCH.json; tail -3 gpurun_out/bench.err
