for wl in dambreak2m lattice2m; do
for pad in 0 9000 16000 24000 42000 62000 99000; do
  B200SPH_FORCES_SMEM_PAD=$pad timeout 300 python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline 2>/dev/null > /tmp/o.json
  python - <<PY
import json
try:
    d=json.load(open('/tmp/o.json')); print("$wl", $pad, round(d["ms_per_step"],3), round(d["roofline"]["kernel_ms"],3))
except Exception as e: print("$wl", $pad, "failed", e)
PY
done; done
