#!/usr/bin/env python
"""Poiseuille validation in the manner of the reference's scripts/validate-poiseuille.py:32-37,95-121 — L-inf / L1 / L2
error of the stream-wise velocity of every fluid particle against the analytic plane-Poiseuille profile
    v_x(z) = F / (2 nu) ((lz/2)^2 - z^2)                         (src/problems/Poiseuille.inc:187-232)
— for the STOCK reference binary and for the DROP-IN binary (the same unmodified problem file on our engines) side by
side. The reference script reads the last VTU file through ParaView (absent here); this one reads the HotFile
checkpoint the run writes at its last iteration (same particle data) and rebuilds z from the cell hash with the grid
the binary prints at start-up (src/GPUSPH.cc:223-226).

    python tools/validate_poiseuille.py [--ppH 16 32 64 100] [--iters 2000] [--out profiles/r02_poiseuille_validation.json]

Runs start ON the analytic profile (--steady-init 1, Poiseuille.inc:166-182): the error after N iterations measures how
well the discretisation holds the steady state (the reference's own script runs to t = 100 s, i.e. ~5e5 iterations at
ppH 100 — out of reach of a test budget). GPU box only.
"""
import argparse
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

LZ, KINVISC, FORCE = 1.0, 0.1, 0.05           # Poiseuille.inc defaults (lz, --kinvisc, --driving-force)


def analytic(z):
    return np.where(np.abs(z) > LZ / 2, 0.0, FORCE / (2 * KINVISC) * ((LZ / 2) ** 2 - z ** 2))


def run(binp, ppH, iters, extra=(), timeout=3000):
    d = tempfile.mkdtemp(prefix="poiseuille_")
    cmd = [binp, "--ppH", str(ppH), "--steady-init", "1", "--maxiter", str(iters), "--dir", d,
           "--checkpoint-every", "1000", "--checkpoints", "0", "--debug", "benchmark_command_runtimes", *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=d, timeout=timeout)
    log = p.stdout + p.stderr
    if p.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed:\n{log[-2000:]}")
    files = sorted(glob.glob(os.path.join(d, "data", "hot_*.bin")))
    hf = read_hotfile(files[-1])
    num = r"([-+0-9.eE]+)"
    o = re.search(r"World origin:\s*" + num + r"\s*,\s*" + num + r"\s*,\s*" + num, log)
    w = re.search(r"World size:\s*" + num + r"\s*x\s*" + num + r"\s*x\s*" + num, log)
    g = re.search(r"Grid size:\s*(\d+) x (\d+) x (\d+)", log)
    origin = np.array([float(x) for x in o.groups()])
    size = np.array([float(x) for x in w.groups()])
    grid = np.array([int(x) for x in g.groups()])
    cyc = re.search(r"Elapsed time of simulation cycle:\s*([0-9.eE+-]+)s", log)
    return hf, origin, size, grid, (float(cyc.group(1)) if cyc else None)


def profile_error(hf, origin, size, grid):
    pos, vel, info, hashv = particle_arrays(hf)
    fluid = (info[:, 0] & 7) == 0
    cs = size / grid
    # default linearisation yzx: hash = x * (Gz * Gy) + z * Gy + y   (src/cuda/cellgrid.cuh:101-106)
    gz = ((hashv & 0x3FFFFFFF) // grid[1]) % grid[2]
    z = origin[2] + (gz + 0.5) * cs[2] + pos[:, 2]
    err = np.abs(vel[fluid, 0].astype(np.float64) - analytic(z[fluid]))
    return {"points": int(fluid.sum()), "linf": float(err.max()), "l1": float(err.mean()), "l2": float(np.sqrt((err ** 2).mean())),
            "max_vel": float(vel[fluid, 0].max()), "max_abs_vy_vz": float(np.abs(vel[fluid, 1:3]).max())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ppH", type=int, nargs="+", default=[16, 32, 64, 100])
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "poiseuille_validation.json"))
    args = ap.parse_args()
    bins = {"reference": os.path.join(ROOT, "oracle", "_ref", "Poiseuille"),
            "dropin": os.path.join(ROOT, "build", "dropin", "Poiseuille_b200")}
    rows = []
    for ppH in args.ppH:
        row = {"ppH": ppH, "iterations": args.iters, "analytic_max_vel": float(analytic(np.zeros(1))[0])}
        for tag, binp in bins.items():
            hf, origin, size, grid, cyc = run(binp, ppH, args.iters)
            row[tag] = profile_error(hf, origin, size, grid)
            row[tag].update(t=hf["t"], particles=hf["particle_count"], cycle_seconds=cyc)
        row["linf_ratio_dropin_over_reference"] = row["dropin"]["linf"] / max(row["reference"]["linf"], 1e-30)
        rows.append(row)
        print(json.dumps(row))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
