#!/bin/bash
# one GPU-box session: parity tests, drop-in check, benches (ours + reference arm)
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python tools/dropin_check.py > gpurun_out/dropin_check.log 2>&1; python - <<'PY'
import json
d=json.load(open("gpurun_out/dropin_report.json"))
for k,v in d.items():
    if "timing" in k: print(k, {t: (round(v[t]["ms_per_step"],3) if isinstance(v.get(t),dict) and "ms_per_step" in v[t] else v.get(t)) for t in ("reference","dropin")}, "speedup", v.get("speedup"))
    else: print(k, {kk: v[kk] for kk in ("same_sorted_slot_fraction","max_vel_err_rel","max_rho_err") if kk in v} or v)
PY
for wl in ${WORKLOADS:-dambreak2m}; do
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 10 2>gpurun_out/bench_err.log > gpurun_out/bench_${wl}.json
  timeout 600 python bench.py --impl reference --workload $wl --steps 20 --warmup 10 2>>gpurun_out/bench_err.log > gpurun_out/bench_ref_${wl}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${wl}.json")); r=json.load(open("gpurun_out/bench_ref_${wl}.json"))
    print("$wl ours ms/step", round(d["ms_per_step"],3), "MIPS", round(d["value"]), "upd/s", round(d["particle_updates_per_s"]/1e6,1), "forces ms", round(d["roofline"]["kernel_ms"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3), "| ref ms/step", r.get("ms_per_step"), "MIPS", r.get("value"), "| ratio", d["value"]/r["value"] if r.get("value") else None, "e2e ratio", d["e2e"]["value"]/r["value"] if r.get("value") else None)
except Exception as e: print("$wl failed", e); print(open("gpurun_out/bench_err.log").read()[-1500:])
PY
done
