#!/bin/bash
# What the driver does at round end, in one session, plus the profiles for profiles/:
# gpu tests, smoke(), default bench (reference arm first, then ours), ncu launch list of the default bench command,
# full ncu captures of the two heaviest kernels.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -5) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/BENCH_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/BENCH_ref.json; echo
timeout 600 python bench.py > gpurun_out/BENCH.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/BENCH.json; echo; tail -2 gpurun_out/bench.err
timeout 600 python bench.py --workload dambreak8m --steps 10 --warmup 10 --no-cpu-baseline > gpurun_out/BENCH_8m.json 2> gpurun_out/bench8m.err
timeout 600 python bench.py --impl reference --workload dambreak8m --steps 10 --warmup 3 > gpurun_out/BENCH_ref_8m.json 2>> gpurun_out/bench8m.err
python - <<'PY'
import json
for a, b in (("BENCH", "BENCH_ref"), ("BENCH_8m", "BENCH_ref_8m")):
    try:
        d = json.load(open(f"gpurun_out/{a}.json")); r = json.load(open(f"gpurun_out/{b}.json"))
        print(a, "ours ms/step", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "| ref ms/step", r.get("ms_per_step"),
              "| value ratio", round(d["value"] / r["value"], 2), "e2e ratio", round(d["e2e"]["value"] / r["value"], 2), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(a, "failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 11 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
B200SPH_FUSED_EULER=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"forces_gather_kernel" -s 4 -c 1 -f -o gpurun_out/prof_forces python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1; tail -1 gpurun_out/ncu_full1.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"build_neibs_kernel" -c 1 -f -o gpurun_out/prof_buildneibs python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_8m.csv python bench.py --workload dambreak8m --steps 11 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench8m.log 2>&1; tail -1 gpurun_out/ncu_bench8m.log | cut -c1-120
