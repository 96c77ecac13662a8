"""Parity against the REAL reference: fixtures under tests/golden/ are HotFile checkpoints written by
the shim-built, otherwise unmodified GPUSPH binary (oracle/_ref/DamBreak3D, see oracle/build_ref.sh)
on a B200, packed by oracle/gen_golden.py. Each fixture holds the reference's particle state at
iterations 0, 10, 20 and 21 of a small DamBreak3D run.

* CPU (not gpu): the oracle restatement reproduces the reference — pins the oracle.
* GPU: the CUDA engines, through the C ABI, reproduce the reference.

Bar: hash and sorted particle order (ids) bit-exact after a neighbour rebuild + one step from the
reference's own state (iteration 20 -> 21); positions within 2e-6 dp and velocities within 2e-5 of the
velocity scale (2e-7 on rho/rho0 - 1) after that step; after 10 steps (0 -> 10, 10 -> 20) within 1e-4 dp /
1e-3 / 5e-5 (the CPU cannot reproduce the GPU's approximate __powf bit for bit; the GPU engines use it
like the reference).
"""
import glob
import os

import numpy as np
import pytest

import oracle_binding as ob
from gpusph_b200 import capi
from gpusph_b200.problems import ParticleArrays, global_positions, initial_dt, make_params, universe_box_planes

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def load(path):
    d = np.load(path)
    dp = float(d["deltap"])
    states = {}
    for it in d["iterations"]:
        it = int(it)
        states[it] = (ParticleArrays(d[f"pos_{it}"], d[f"vel_{it}"], d[f"info_{it}"], d[f"hash_{it}"]),
                      float(d[f"t_{it}"]), float(d[f"dt_{it}"]))
    n = states[0][0].n
    # DamBreak3D: origin 0, size 1.6 x 0.67 x 0.6 (src/problems/DamBreak3D.cu:107-122), neiblist 128,
    # artificial viscosity, c0 = 20, gamma = 7, diffusion coefficient 0.1 whatever the model (:46,:95), H = 0.4 (:82)
    flags = capi.ENABLE_DTADAPT | capi.ENABLE_REPACKING | (capi.ENABLE_PLANES if _opt(d, "use_planes") else 0)
    params = make_params(origin=(0, 0, 0), size=(1.6, 0.67, 0.6), deltap=dp, allocated_particles=n,
                         densitydiffusion=int(d["rhodiff"]), density_diff_coeff=0.1 if int(d["rhodiff"]) else None,
                         simflags=flags, maxfall=0.4)
    return params, states


def _opt(d, key):
    return int(d[key]) if key in d.files else 0


def single_step_rho_tol(path):
    """rho~ tolerance of the 20 -> 21 step. With --mls N iteration 20 applies the MLS filter first: a 4 x 4 moment matrix
    inverted in float per particle (src/cuda/forces_kernel.cu:509-721), whose result moves by ~1e-6 with the order of
    the float operations (FMA contraction of the GPU build vs the CPU oracle): 2e-6 there, 2e-7 otherwise."""
    return 2e-6 if _opt(np.load(path), "mls") else None


def options(path, params):
    """Worker / OracleWorker keyword arguments of a fixture's DamBreak3D options: --mls N adds the MLS filter every N
    iterations (src/problems/DamBreak3D.cu:63-71), --use_planes 1 replaces the boundary box by the six planes of the
    universe box (:127-130). The obstacle of --num_obstacles 1 is a force-feedback body that never moves (no callback,
    :169-180): its particles carry FG_MOVING_BOUNDARY | FG_COMPUTE_FORCE and are integrated like fixed DYN boundaries."""
    d = np.load(path)
    kw_gpu, kw_cpu = {}, {}
    if _opt(d, "mls"):
        from gpusph_b200.engines import MLS_FILTER
        kw_gpu["filters"] = {MLS_FILTER: _opt(d, "mls")}
        kw_cpu["filters"] = {"MLS_FILTER": _opt(d, "mls")}
    if _opt(d, "use_planes"):
        kw_gpu["planes"] = kw_cpu["planes"] = universe_box_planes(params, (0, 0, 0), (1.6, 0.67, 0.6))
    return kw_gpu, kw_cpu


def ids_of(info):
    return (info[:, 3].astype(np.int64) << 16) | info[:, 2]


def compare(params, got: ParticleArrays, exp: ParticleArrays, pos_tol_dp, vel_tol, exact_order, rho_tol=None):
    assert got.n == exp.n
    if exact_order:
        assert np.array_equal(got.hash, exp.hash), "cell hash differs from the reference"
        assert np.array_equal(got.info, exp.info), "sorted particle order differs from the reference"
    og, oe = np.argsort(ids_of(got.info)), np.argsort(ids_of(exp.info))
    assert np.array_equal(ids_of(got.info)[og], ids_of(exp.info)[oe])
    live = (exp.info[oe, 0] & 7) != capi.PT_TESTPOINT      # the reference's TESTPOINTS post-process rewrites their vel
    gp = global_positions(params, got.pos, got.hash)[og]
    ep = global_positions(params, exp.pos, exp.hash)[oe]
    perr = np.abs(gp - ep).max() / float(params.deltap)
    vs = max(np.abs(exp.vel[:, :3]).max(), 1e-3)
    verr = np.abs(got.vel[og][live, :3] - exp.vel[oe][live, :3]).max() / vs
    rerr = np.abs(got.vel[og][live, 3] - exp.vel[oe][live, 3]).max()
    assert perr < pos_tol_dp, f"position error {perr:.2e} dp"
    assert verr < vel_tol, f"velocity error {verr:.2e}"
    assert rerr < (rho_tol if rho_tol is not None else vel_tol * 1e-2), f"density error {rerr:.2e}"
    assert np.array_equal(got.pos[og, 3], exp.pos[oe, 3])   # masses untouched


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_fixture_matches_host_setup(path):
    """The parameters we derive (grid, cell size, initial dt) are the reference's: its initial state hashes
    consistently with our grid and its first dt equals ProblemCore::check_dt restated in problems.py."""
    params, states = load(path)
    p0, t0, dt0 = states[0]
    assert dt0 == pytest.approx(initial_dt(params), rel=1e-7)
    assert int(p0.hash.max()) < params.num_cells
    cs = np.array([params.cell_size[a] for a in range(3)])
    assert (np.abs(p0.pos[:, :3]) <= 0.5 * cs * (1 + 1e-5)).all()
    gp = global_positions(params, p0.pos, p0.hash)
    assert gp.min() >= -1e-6 and (gp.max(axis=0) <= np.array([1.6, 0.67, 0.6]) + 1e-6).all()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_dambreak_problem_is_the_reference_particle_set(path):
    """gpusph_b200.problems.dambreak_problem (what bench.py and the parity tests run) fills the tank exactly like the
    reference's DamBreak3D (--num_obstacles 0): same particle count per type, same positions (to the last bit of the
    cell-local floats), same cells, same masses, hydrostatic rho~ to float rounding. Ids / order are ours."""
    from scipy.spatial import cKDTree
    from gpusph_b200.problems import dambreak_problem
    d = np.load(path)
    if int(d["num_obstacles"]) if "num_obstacles" in d else 0:
        pytest.skip("fixture with the obstacle")
    if int(d["use_planes"]) if "use_planes" in d else 0:
        pytest.skip("fixture with planes instead of the boundary box")
    params, states = load(path)
    ref = states[0][0]
    mp, mine = dambreak_problem(float(d["deltap"]), densitydiffusion=int(d["rhodiff"]), testpoints=3)
    assert mine.n == ref.n
    tr, tm = ref.info[:, 0] & 7, mine.info[:, 0] & 7
    gr, gm = global_positions(params, ref.pos, ref.hash), global_positions(mp, mine.pos, mine.hash)
    for k in (capi.PT_FLUID, capi.PT_BOUNDARY, capi.PT_TESTPOINT):
        assert (tr == k).sum() == (tm == k).sum()
        dist, idx = cKDTree(gm[tm == k]).query(gr[tr == k])
        assert dist.max() < 1e-7 and len(np.unique(idx)) == idx.size
        a, b = ref.pos[tr == k], mine.pos[tm == k][idx]
        assert np.abs(a[:, :3] - b[:, :3]).max() < 1e-8
        assert np.array_equal(a[:, 3], b[:, 3]), "masses differ"
        assert np.array_equal(ref.hash[tr == k], mine.hash[tm == k][idx]), "cells differ"
        assert np.abs(ref.vel[tr == k][:, 3] - mine.vel[tm == k][idx][:, 3]).max() < 2e-7


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference_single_step(path):
    params, states = load(path)
    s20, _, dt20 = states[20]
    s21, _, dt21 = states[21]
    w = ob.OracleWorker(params, s20, start_iteration=20, dt=dt20, **options(path, params)[1])
    w.step()
    compare(params, w.download(), s21, pos_tol_dp=2e-6, vel_tol=2e-5, exact_order=True, rho_tol=single_step_rho_tol(path))
    assert w.dt == pytest.approx(dt21, rel=1e-5)
    assert w.neibs_info.has_too_many_neibs == -1


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("start", [0, 10])
def test_oracle_reproduces_reference_ten_steps(path, start):
    params, states = load(path)
    s0, t0, dt0 = states[start]
    s1, t1, dt1 = states[start + 10]
    w = ob.OracleWorker(params, s0, start_iteration=start, dt=dt0, **options(path, params)[1])
    for _ in range(10):
        w.step()
    assert w.t == pytest.approx(t1 - t0, rel=1e-5)
    assert w.dt == pytest.approx(dt1, rel=1e-4)
    compare(params, w.download(), s1, pos_tol_dp=1e-4, vel_tol=1e-3, exact_order=False, rho_tol=5e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_gpu_reproduces_reference_single_step(path):
    from gpusph_b200.simulation import Worker
    params, states = load(path)
    s20, _, dt20 = states[20]
    s21, _, dt21 = states[21]
    w = Worker(params, s20, 0, start_iteration=20, dt=dt20, clobber=True, **options(path, params)[0])
    w.step()
    compare(params, w.download(), s21, pos_tol_dp=2e-6, vel_tol=2e-5, exact_order=True, rho_tol=single_step_rho_tol(path))
    assert w.dt == pytest.approx(dt21, rel=1e-5)
    assert w.last_neibs_info.has_too_many_neibs == -1


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
@pytest.mark.parametrize("start", [0, 10])
def test_gpu_reproduces_reference_ten_steps(path, start):
    from gpusph_b200.simulation import Worker
    params, states = load(path)
    s0, t0, dt0 = states[start]
    s1, t1, dt1 = states[start + 10]
    w = Worker(params, s0, 0, start_iteration=start, dt=dt0, clobber=True, **options(path, params)[0])
    for _ in range(10):
        w.step()
    assert w.t == pytest.approx(t1 - t0, rel=1e-5)
    assert w.dt == pytest.approx(dt1, rel=1e-4)
    compare(params, w.download(), s1, pos_tol_dp=1e-4, vel_tol=1e-3, exact_order=False, rho_tol=5e-5)
