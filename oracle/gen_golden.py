#!/usr/bin/env python
"""Generate golden fixtures from the REAL reference (run on the GPU box; the reference needs a GPU).

    gpurun -- python oracle/gen_golden.py          # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/        # then commit them

Runs the shim-built reference binary oracle/_ref/DamBreak3D (see oracle/build_ref.sh) on small
DamBreak3D configurations with a HotFile checkpoint at EVERY iteration (--checkpoint-every 0
--checkpoints 0) and packs the states at iterations 0,1,10,11,20,21 into one compressed .npz per
configuration. These are the reference's own outputs on its own problem: the oracle
(oracle/sph_oracle.c) and the CUDA engines are both pinned against them (tests/test_golden.py).
"""
import glob
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpusph_b200.hotfile import particle_arrays, read_hotfile  # noqa: E402  (file-format reader only: no engine is loaded)

CONFIGS = {
    # name: (deltap, density-diffusion enum: 0 none, 1 Ferrari, 2 Colagrossi, 3 Brezzi, extra DamBreak3D options
    #        {mls, num_obstacles, use_planes} - src/problems/DamBreak3D.cu:41-71)
    "dambreak_dp040_colagrossi": (0.04, 2, {}),
    "dambreak_dp050_none": (0.05, 0, {}),
    "dambreak_dp045_ferrari": (0.045, 1, {}),
    # the options whose oracle restatement was pinned by formulas only in round 1
    "dambreak_dp050_mls10": (0.05, 0, {"mls": 10}),
    "dambreak_dp050_brezzi": (0.05, 3, {}),
    "dambreak_dp050_planes": (0.05, 2, {"use_planes": 1}),
    "dambreak_dp050_obstacle": (0.05, 2, {"num_obstacles": 1}),
}
KEEP = (0, 1, 10, 11, 20, 21)


def main():
    binp = os.path.join(ROOT, "oracle", "_ref", "DamBreak3D")
    outdir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    only = set(sys.argv[1:])
    for name, (dp, rhodiff, extra) in CONFIGS.items():
        if only and name not in only:
            continue
        d = tempfile.mkdtemp(prefix="golden_")
        opts = {"mls": 0, "num_obstacles": 0, "use_planes": 0}
        opts.update(extra)
        cmd = [binp, "--deltap", str(dp), "--maxiter", "21", "--dir", d, "--checkpoint-every", "0",
               "--checkpoints", "0", "--num_obstacles", str(opts["num_obstacles"]), "--density-diffusion", str(rhodiff),
               "--mls", str(opts["mls"]), "--use_planes", str(opts["use_planes"])]
        p = subprocess.run(cmd, capture_output=True, text=True, cwd=d)
        log = p.stdout + p.stderr
        open(os.path.join(outdir, name + ".log"), "w").write(" ".join(cmd) + "\n" + log)
        if p.returncode != 0:
            print(f"{name}: reference failed rc={p.returncode}", log[-500:])
            continue
        files = sorted(glob.glob(os.path.join(d, "data", "hot_*.bin")))
        states = {}
        for f in files:
            hf = read_hotfile(f)
            if hf["iterations"] in KEEP and hf["iterations"] not in states:
                states[hf["iterations"]] = hf
        pack = {"deltap": np.float64(dp), "rhodiff": np.int32(rhodiff), "iterations": np.array(sorted(states), dtype=np.int64),
                "mls": np.int32(opts["mls"]), "num_obstacles": np.int32(opts["num_obstacles"]), "use_planes": np.int32(opts["use_planes"])}
        for it, hf in states.items():
            pos, vel, info, hashv = particle_arrays(hf)
            pack[f"pos_{it}"], pack[f"vel_{it}"], pack[f"info_{it}"], pack[f"hash_{it}"] = pos, vel, info, hashv
            pack[f"t_{it}"] = np.float64(hf["t"])
            pack[f"dt_{it}"] = np.float32(hf["dt"])
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **pack)
        print(f"{name}: {len(files)} hotfiles, kept iterations {sorted(states)}, {hf['particle_count']} particles, "
              f"buffers {list(hf['buffers'])}")


if __name__ == "__main__":
    main()
