"""HotFile compatibility with the REAL reference binary (oracle/_ref/DamBreak3D, GPU box only):
* a HotFile the reference wrote, decoded and re-encoded by gpusph_b200/hotfile.py, is byte-identical;
* Worker.from_hotfile continues the reference's run (reference 0..10, ours 10..20 against the reference's own 20);
* Worker.save_hotfile writes a file the reference `--resume`s from (relay: reference 0..10, ours 10..20, reference
  20..30, against the reference's own uninterrupted 0..30).

The hand-over to the reference must happen at a multiple of buildneibsfreq: resumed at any other iteration the
reference skips its NEIBS_LIST phase (needs_new_neibs, src/Integrator.cc:85-92) and its forces kernel reads a
neighbour list that was never built (observed: resuming at iteration 15, from our file, it loads the file, prints
"Restarting from t=0.0133841, iteration=15" and dies in FORCES_SYNC with an illegal memory access).
"""
import glob
import os
import subprocess

import numpy as np
import pytest

import test_golden as tg
from gpusph_b200 import hotfile as hfmod
from gpusph_b200.hotfile import particle_arrays, read_hotfile, write_hotfile
from gpusph_b200.problems import ParticleArrays, make_params

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "DamBreak3D")
DP, RHODIFF = 0.05, 1          # Ferrari


def run_ref(d, maxiter, *extra):
    os.makedirs(d, exist_ok=True)
    cmd = [REF, "--deltap", str(DP), "--maxiter", str(maxiter), "--dir", d, "--checkpoint-every", "0", "--checkpoints", "0",
           "--num_obstacles", "0", "--density-diffusion", str(RHODIFF), "--mls", "0", *extra]
    p = subprocess.run(cmd, capture_output=True, text=True, cwd=d, timeout=240)
    log = p.stdout + p.stderr
    assert p.returncode == 0, f"reference failed: {' '.join(cmd)}\n{log[-1500:]}"
    hot = {}
    for f in sorted(glob.glob(os.path.join(d, "data", "hot_*.bin"))):
        hot.setdefault(read_hotfile(f)["iterations"], f)
    return hot, log


def params_for(n):
    return make_params(origin=(0, 0, 0), size=(1.6, 0.67, 0.6), deltap=DP, allocated_particles=n,
                       densitydiffusion=RHODIFF, density_diff_coeff=0.1)


@pytest.fixture(scope="module")
def reference_run(tmp_path_factory):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/DamBreak3D not built")
    hot, _ = run_ref(str(tmp_path_factory.mktemp("full")), 30)
    assert {10, 20, 30} <= set(hot), f"checkpoints at {sorted(hot)}"
    return hot


@pytest.mark.timeout(600)
def test_reencoded_hotfile_is_byte_identical_and_our_run_continues_the_reference(reference_run, tmp_path):
    from gpusph_b200.simulation import Worker
    hot = reference_run
    hf = read_hotfile(hot[10])
    assert hf["body_count"] == 0 and sorted(hf["buffers"]) == ["Hash", "Info", "Position", "Velocity"]
    layout = dict(buffer_count=hf["buffer_count"], num_open_boundaries=hf["num_open_boundaries"], order=list(hf["buffers"]))
    pos, vel, info, hashv = particle_arrays(hf)
    again = str(tmp_path / "reencoded.bin")
    write_hotfile(again, pos, vel, info, hashv, iterations=hf["iterations"], t=hf["t"], dt=hf["dt"], **layout)
    assert open(again, "rb").read() == open(hot[10], "rb").read()

    params = params_for(pos.shape[0])
    w = Worker.from_hotfile(params, hot[10], 0, clobber=True)
    assert w.iterations == 10 and w.t == pytest.approx(hf["t"]) and w.dt == pytest.approx(hf["dt"])
    for _ in range(10):
        w.step()
    mine = str(tmp_path / "hot_ours_00020.bin")
    w.save_hotfile(mine, **layout)
    h20 = read_hotfile(mine)
    assert h20["iterations"] == 20 and h20["t"] > hf["t"] and h20["buffer_count"] == hf["buffer_count"]
    # the reference checkpoints at its neighbour-rebuild iterations (and at the end): its own state at 20
    r20 = read_hotfile(hot[20])
    assert h20["t"] == pytest.approx(r20["t"], rel=5e-5)
    tg.compare(params, w.download(), ParticleArrays(*particle_arrays(r20)), pos_tol_dp=1e-4, vel_tol=1e-3, exact_order=False, rho_tol=5e-5)


@pytest.mark.timeout(600)
def test_reference_resumes_from_our_hotfile(reference_run, tmp_path):
    from gpusph_b200.simulation import Worker
    hot = reference_run
    hf = read_hotfile(hot[10])
    # the defaults of the writer (what save_hotfile() writes without overrides) are the reference's layout for this run
    assert hf["buffer_count"] == hfmod.PLAIN_BUFFER_COUNT and list(hf["buffers"]) == [b[0] for b in hfmod.PLAIN_BUFFERS]
    params = params_for(hf["particle_count"])
    w = Worker.from_hotfile(params, hot[10], 0, clobber=True)
    for _ in range(10):
        w.step()
    mine = str(tmp_path / "hot_ours_00020.bin")
    w.save_hotfile(mine)
    hot2, log = run_ref(str(tmp_path / "resumed"), 30, "--resume", mine)
    assert "Restarting from t=" in log, log[-1500:]
    assert 30 in hot2, f"resumed run wrote checkpoints at {sorted(hot2)}"
    got = ParticleArrays(*particle_arrays(read_hotfile(hot2[30])))
    exp = ParticleArrays(*particle_arrays(read_hotfile(hot[30])))
    tg.compare(params, got, exp, pos_tol_dp=2e-4, vel_tol=2e-3, exact_order=False, rho_tol=1e-4)
    assert read_hotfile(hot2[30])["t"] == pytest.approx(read_hotfile(hot[30])["t"], rel=5e-5)
