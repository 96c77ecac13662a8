for wl in dambreak2m lattice2m; do for pf in 1 2 4; do
B200SPH_GATHER_PF=$pf timeout 300 python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline 2>/dev/null > /tmp/o.json
python - <<PY
import json
try:
    d=json.load(open("/tmp/o.json")); print("$wl pf=$pf ms/step", round(d["ms_per_step"],3), "forces ms", round(d["roofline"]["kernel_ms"],4))
except Exception as e: print("$wl pf=$pf failed", e)
PY
done; done
