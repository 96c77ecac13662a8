#!/usr/bin/env python
"""Run the stock reference binary and the drop-in binary (reference host code + our three hot engines, see
tools/build_dropin.sh) on the same DamBreak3D configuration; compare their final HotFile states and their own
per-command timers. GPU box only. Writes gpurun_out/dropin_report.json."""
import glob
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpusph_b200.hotfile import particle_arrays, read_hotfile  # noqa: E402


def run(binp, dp, maxiter, rhodiff, save, obstacles=0):
    d = tempfile.mkdtemp(prefix="dropin_")
    cmd = [binp, "--deltap", str(dp), "--maxiter", str(maxiter), "--dir", d, "--num_obstacles", str(obstacles),
           "--density-diffusion", str(rhodiff), "--debug", "benchmark_command_runtimes"]
    cmd += ["--checkpoint-every", "1000", "--checkpoints", "0"] if save else ["--nosave"]
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=d)
    out = p.stdout + p.stderr
    times = {}
    for ln in out.splitlines():
        if ln.startswith("CMDTIMES:") and not ln.startswith("CMDTIMES:COMMAND"):
            f = ln[len("CMDTIMES:"):].split("\t")
            times[f[0]] = float(f[4])
    state = None
    if save:
        files = sorted(glob.glob(os.path.join(d, "data", "hot_*.bin")))
        if files:
            hf = read_hotfile(files[-1])
            state = (hf["iterations"], hf["t"], particle_arrays(hf))
    m = re.findall(r"iteration=[\d,]+, dt=[0-9.eE+-]+s, ([\d,]+) parts", out)
    n = int(m[-1].replace(",", "")) if m else None
    return p.returncode, out, times, state, n


def ids_of(info):
    return (info[:, 3].astype(np.int64) << 16) | info[:, 2]


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "DamBreak3D")
    ours = os.path.join(ROOT, "build", "dropin", "DamBreak3D_b200")
    report = {}
    # 1. parity: 100 iterations of a small case, final states compared particle by particle
    for name, dp, rhodiff, obst in (("dp0.02_colagrossi", 0.02, 2, 0), ("dp0.02_ferrari", 0.02, 1, 0),
                                    ("dp0.02_default_obstacle_colagrossi", 0.02, 2, 1)):
        rc_r, out_r, _, st_r, n = run(ref, dp, 100, rhodiff, True, obst)
        rc_o, out_o, _, st_o, _ = run(ours, dp, 100, rhodiff, True, obst)
        if rc_r or rc_o or st_r is None or st_o is None:
            report[name] = {"error": f"rc {rc_r}/{rc_o}", "ours_tail": out_o[-600:], "ref_tail": out_r[-300:]}
            continue
        (it_r, t_r, (pr, vr, ir, hr)), (it_o, t_o, (po, vo, io, ho)) = st_r, st_o
        a, b = np.argsort(ids_of(ir)), np.argsort(ids_of(io))
        live = (ir[a, 0] & 7) != 3
        vs = np.abs(vr[:, :3]).max()
        report[name] = {
            "particles": n, "iterations": [int(it_r), int(it_o)], "t": [t_r, t_o],
            "same_cell_fraction": float((hr[a] == ho[b]).mean()),
            "same_sorted_slot_fraction": float((ids_of(ir) == ids_of(io)).mean()),
            "max_vel_err_rel": float(np.abs(vr[a][live, :3] - vo[b][live, :3]).max() / vs),
            "max_rho_err": float(np.abs(vr[a][live, 3] - vo[b][live, 3]).max()),
            "max_localpos_err_same_cell": float(np.abs(pr[a][hr[a] == ho[b]] - po[b][hr[a] == ho[b]])[:, :3].max()),
        }
    # 2. timing: ~2M particles, the reference's own per-command timers, K = 100 steps after 10 of warm-up
    for name, dp in (("timing_dp0.0043_ferrari", 0.0043),):
        res = {}
        for tag, binp in (("reference", ref), ("dropin", ours)):
            _, _, t_w, _, n = run(binp, dp, 10, 1, False)
            rc, out, t_k, _, n = run(binp, dp, 110, 1, False)
            if rc or not t_k:
                res[tag] = {"error": out[-500:]}
                continue
            ph = {k: (t_k[k] - t_w.get(k, 0.0)) / 100 for k in t_k}
            res[tag] = {"particles": n, "ms_per_step": sum(ph.values()),
                        "phases_ms_per_step": dict(sorted(ph.items(), key=lambda kv: -kv[1])[:8])}
        if "ms_per_step" in res.get("reference", {}) and "ms_per_step" in res.get("dropin", {}):
            res["speedup"] = res["reference"]["ms_per_step"] / res["dropin"]["ms_per_step"]
        report[name] = res
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "dropin_report.json"), "w"), indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
