"""GPU parity tests: every C-ABI entry point of the hot path against the CPU oracle on the same
seeded inputs (through the engine mirror, i.e. through the C ABI), plus size-independent
properties at full benchmark size. Bar: bit-exact for hash / sort permutation / cellStart/End /
neighbour list / counters; stated float tolerances for forces, dt and integrated fields."""
import numpy as np
import pytest
import torch

import oracle_binding as ob
from gpusph_b200 import capi
from gpusph_b200.engines import (BUFFER_CELLEND, BUFFER_CELLSTART, BUFFER_CFL, BUFFER_FORCES, BUFFER_HASH,
                                 BUFFER_INFO, BUFFER_NEIBSLIST, BUFFER_PARTINDEX, BUFFER_POS, BUFFER_VEL, BufferList,
                                 SimFramework)
from gpusph_b200.problems import dambreak_problem, global_positions, lattice_problem, poiseuille_problem
from gpusph_b200.simulation import Worker
from gpusph_b200.engines import neibs_list_blocked, neibs_list_rows

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.view(dtype)
    return t.to(DEV)


def host(t, dtype=None):
    a = t.cpu().numpy()
    return a if dtype is None else a.view(dtype)


def problems():
    rng = np.random.default_rng(42)
    out = {}
    out["lattice"] = lattice_problem(20, jitter=0.3, densitydiffusion=capi.RHODIFF_COLAGROSSI)
    out["dambreak"] = dambreak_problem(0.03, densitydiffusion=capi.RHODIFF_COLAGROSSI)
    out["dambreak_ferrari"] = dambreak_problem(0.04, densitydiffusion=capi.RHODIFF_FERRARI, density_diff_coeff=0.1)
    out["periodic"] = lattice_problem(16, jitter=0.3, periodic=capi.PERIODIC_X | capi.PERIODIC_Y)
    out["ragged"] = lattice_problem(7, ny=5, nz=3, jitter=0.2)      # fewer particles than one block row
    # Newtonian laminar (Morris) viscosity, periodic XY, DYN plates: the Poiseuille specialisations
    out["poiseuille"] = poiseuille_problem(12, viscavgop=capi.AVG_HARMONIC)
    out["poiseuille_geo"] = poiseuille_problem(10, viscavgop=capi.AVG_GEOMETRIC)
    out["laminar_artvisc"] = lattice_problem(12, jitter=0.3, rheology=capi.RHEOLOGY_NEWTONIAN, kinvisc=5e-3,
                                             densitydiffusion=capi.RHODIFF_FERRARI, density_diff_coeff=0.1)
    # two fluids (different rest density, EOS and viscosity) in one lattice: the MULTIFLUID kernel variants
    two = [dict(rho0=800.0, gamma=5.0, c0=25.0, kinvisc=2e-3)]
    for nm, kw in (("twofluid", dict(densitydiffusion=capi.RHODIFF_COLAGROSSI)),
                   ("twofluid_laminar", dict(rheology=capi.RHEOLOGY_NEWTONIAN, kinvisc=5e-3, viscavgop=capi.AVG_HARMONIC,
                                             densitydiffusion=capi.RHODIFF_FERRARI, density_diff_coeff=0.1))):
        p2, q2 = lattice_problem(12, jitter=0.3, fluids=two, **kw)
        upper = global_positions(p2, q2.pos, q2.hash)[:, 2] > 0.06
        q2.info[upper, 1] = np.uint16(1 << 12)                       # fluid number 1 (src/particleinfo.h:135-160)
        q2.pos[upper, 3] *= np.float32(0.8)
        out[nm] = (p2, q2)
    for params, parts in out.values():
        fl = (parts.info[:, 0] & 7) == 0
        parts.vel[:, :3] += rng.normal(0, 0.3, size=(parts.n, 3)).astype(np.float32) * fl[:, None]
        parts.vel[:, 3] += rng.normal(0, 1e-3, size=parts.n).astype(np.float32)
    return out


PROBLEMS = None


def get(name):
    global PROBLEMS
    if PROBLEMS is None:
        PROBLEMS = problems()
    params, parts = PROBLEMS[name]
    return params, parts


NAMES = ["lattice", "dambreak", "dambreak_ferrari", "periodic", "ragged", "poiseuille", "poiseuille_geo", "laminar_artvisc",
         "twofluid", "twofluid_laminar"]


class Pipeline:
    """Runs the neighbour pipeline on the GPU (through the engines) and on the oracle, keeping every
    intermediate so that each stage can be compared."""

    def __init__(self, name, move=0.004, first=False):
        params, parts = get(name)
        self.params, self.parts = params, parts
        n = parts.n
        self.n = n
        fw = SimFramework(params, 0)
        self.fw = fw
        rng = np.random.default_rng(9)
        pos0 = parts.pos.copy()
        if not first:
            pos0[:, :3] += rng.uniform(-move, move, size=(n, 3)).astype(np.float32)
        # --- oracle ---
        o = {}
        o["pos"], o["hash"], o["info"] = pos0.copy(), parts.hash.copy(), parts.info.copy()
        if first:
            o["pidx"] = ob.fix_hash(params, o["hash"], o["info"])
        else:
            o["pidx"] = ob.calc_hash(params, o["pos"], o["hash"], o["info"])
        o["hash_unsorted"], o["pos_unsorted"] = o["hash"].copy(), o["pos"].copy()
        ob.sort(o["hash"], o["info"], o["pidx"])
        o["cs"], o["ce"], _, o["spos"], o["svel"], o["newn"] = ob.reorder(params, o["pos"], parts.vel, o["info"], o["hash"], o["pidx"])
        o["nl"], o["ninfo"] = ob.build_neibs(params, o["spos"], o["info"], o["hash"], o["cs"], o["ce"])
        self.o = o
        # --- device, through the engine mirror / C ABI ---
        A = int(params.neiblist_stride)
        g = {}
        g["pos"] = dev(pos0)
        g["vel"] = dev(parts.vel)
        g["info"] = dev(parts.info.view(np.int16))
        g["hash"] = dev(parts.hash.view(np.int32))
        g["pidx"] = torch.zeros(n, dtype=torch.int32, device=DEV)
        b = BufferList({BUFFER_POS: g["pos"], BUFFER_VEL: g["vel"], BUFFER_INFO: g["info"], BUFFER_HASH: g["hash"],
                        BUFFER_PARTINDEX: g["pidx"]})
        if first:
            fw.neibsEngine.fixHash(b, b, n)
        else:
            fw.neibsEngine.calcHash(b, b, n)
        g["hash_unsorted"] = g["hash"].clone()
        g["pos_unsorted"] = g["pos"].clone()
        fw.neibsEngine.sort(b, b, n)
        g["cs"] = torch.full((params.num_cells,), -1, dtype=torch.int32, device=DEV)
        g["ce"] = torch.full((params.num_cells,), -1, dtype=torch.int32, device=DEV)
        g["spos"] = torch.zeros_like(g["pos"])
        g["svel"] = torch.zeros_like(g["vel"])
        g["newn"] = torch.zeros(1, dtype=torch.int32, device=DEV)
        srt = BufferList(b)
        srt.update({BUFFER_POS: g["spos"], BUFFER_VEL: g["svel"], BUFFER_CELLSTART: g["cs"], BUFFER_CELLEND: g["ce"]})
        fw.neibsEngine.reorderDataAndFindCellStart(None, srt, b, n, g["newn"])
        g["nl"] = torch.full((int(params.neiblistsize), A), -1, dtype=torch.int16, device=DEV)
        srt[BUFFER_NEIBSLIST] = g["nl"]
        fw.neibsEngine.resetinfo()
        fw.neibsEngine.buildNeibsList(srt, srt, n, n)
        g["ninfo"] = fw.neibsEngine.getinfo()
        self.g = g
        self.sorted = srt


@pytest.fixture(scope="module", params=NAMES)
def pipe(request):
    return Pipeline(request.param)


def test_calc_hash_bit_exact(pipe):
    assert np.array_equal(host(pipe.g["hash_unsorted"], np.uint32), pipe.o["hash_unsorted"])
    # positions are re-localised with one FMA: bitwise identical too
    a, b = host(pipe.g["pos_unsorted"]), pipe.o["pos_unsorted"]
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_sort_bit_exact(pipe):
    assert np.array_equal(host(pipe.g["hash"], np.uint32), pipe.o["hash"])
    assert np.array_equal(host(pipe.g["info"], np.uint16), pipe.o["info"])
    assert np.array_equal(host(pipe.g["pidx"], np.uint32), pipe.o["pidx"])


def test_reorder_and_cell_ranges_bit_exact(pipe):
    assert np.array_equal(host(pipe.g["cs"], np.uint32), pipe.o["cs"])
    used = pipe.o["cs"] != 0xFFFFFFFF
    assert np.array_equal(host(pipe.g["ce"], np.uint32)[used], pipe.o["ce"][used])
    assert int(pipe.g["newn"].item()) == pipe.o["newn"]
    assert np.array_equal(host(pipe.g["spos"]).view(np.uint32), pipe.o["spos"].view(np.uint32))
    assert np.array_equal(host(pipe.g["svel"]).view(np.uint32), pipe.o["svel"].view(np.uint32))


def test_reorder_gathers_extra_buffers():
    """b200sph_reorder's `extras`: every further per-particle buffer the caller wants re-ordered with POS / VEL (the
    reference's reorderDataAndFindCellStart sorts all the buffers it is handed, src/cuda/buildneibs.cu:220-330; the C++
    adapter passes the optional ones this way) - 4-, 8- and 16-byte elements, against the sorted particle index."""
    p = Pipeline("dambreak")
    g, n, fw = p.g, p.n, p.fw
    gen = torch.Generator(device="cpu").manual_seed(3)
    tke = torch.randn(n, generator=gen).to(DEV)                       # float
    vertpos = torch.randn((n, 2), generator=gen).to(DEV)              # float2
    eulervel = torch.randn((n, 4), generator=gen).to(DEV)             # float4
    uns = BufferList({BUFFER_POS: g["pos_unsorted"], BUFFER_VEL: g["vel"], BUFFER_INFO: g["info"], BUFFER_HASH: g["hash"],
                      BUFFER_PARTINDEX: g["pidx"], "BUFFER_TKE": tke, "BUFFER_VERTPOS": vertpos, "BUFFER_EULERVEL": eulervel})
    srt = BufferList(uns)
    out = {k: torch.zeros_like(v) for k, v in (("BUFFER_TKE", tke), ("BUFFER_VERTPOS", vertpos), ("BUFFER_EULERVEL", eulervel))}
    srt.update({BUFFER_POS: torch.zeros_like(g["pos"]), BUFFER_VEL: torch.zeros_like(g["vel"]),
                BUFFER_CELLSTART: torch.full_like(g["cs"], -1), BUFFER_CELLEND: torch.full_like(g["ce"], -1), **out})
    newn = torch.zeros(1, dtype=torch.int32, device=DEV)
    fw.neibsEngine.reorderDataAndFindCellStart(None, srt, uns, n, newn, extra_keys=tuple(out))
    torch.cuda.synchronize()
    idx = g["pidx"].long()
    assert torch.equal(srt[BUFFER_POS], g["spos"]) and torch.equal(srt[BUFFER_CELLSTART], g["cs"])
    assert torch.equal(out["BUFFER_TKE"], tke[idx])
    assert torch.equal(out["BUFFER_VERTPOS"], vertpos[idx])
    assert torch.equal(out["BUFFER_EULERVEL"], eulervel[idx])
    # an element size the gather has no kernel for is refused, not skipped
    bad = torch.zeros((n, 3), device=DEV)
    uns["BUFFER_BAD"], srt["BUFFER_BAD"] = bad, torch.zeros_like(bad)
    with pytest.raises(ValueError):
        fw.neibsEngine.reorderDataAndFindCellStart(None, srt, uns, n, newn, extra_keys=("BUFFER_BAD",))


def test_neighbour_list_bit_exact(pipe):
    got = host(neibs_list_rows(pipe.g["nl"]), np.uint16)       # blocked layout -> the reference's (and the oracle's)
    assert np.array_equal(got, pipe.o["nl"])
    gi, oi = pipe.g["ninfo"], pipe.o["ninfo"]
    assert gi.num_interactions == oi.num_interactions
    assert gi.max_fluid_boundary_neibs == oi.max_fluid_boundary_neibs
    assert gi.has_too_many_neibs == -1 and oi.has_too_many_neibs == -1


@pytest.mark.parametrize("block", [32, 128, 1024])
def test_blocked_neighbour_list_layout(block):
    """Params.neiblist_block (include/b200sph.h): the same list VALUES in blocks of `block` particles - against the oracle
    list, and the forces / filters reading it against the default layout, bitwise. dambreak has ~5 800 particles: 182 /
    46 / 6 blocks, the last one narrower."""
    params, parts = get("dambreak")
    pb = params.copy()
    pb.neiblist_block = block
    from gpusph_b200.engines import MLS_FILTER, SHEPARD_FILTER
    filters = {MLS_FILTER: 2} if block == 32 else ({SHEPARD_FILTER: 2} if block == 128 else None)   # they walk the list too
    a, b = Worker(params, parts, 0, filters=filters), Worker(pb, parts, 0, filters=filters)
    assert int(params.neiblist_stride) % block != 0, "the last block should be a narrow one"
    for _ in range(3):
        a.step(); b.step()
    n = a.numParticles
    la, lb = neibs_list_rows(a.neibslist), neibs_list_rows(b.neibslist, block)
    assert torch.equal(la[:, :n], lb[:, :n])
    assert not torch.equal(a.neibslist, b.neibslist), "the raw buffers differ: another layout"
    ref = ob.OracleWorker(params, parts)
    ref.build_neibs()
    w = Worker(pb, parts, 0)
    w.build_neibs()
    assert np.array_equal(host(neibs_list_rows(w.neibslist, block), np.uint16)[:, :n], ref.neibslist[:, :n])
    ga, gb = a.download(), b.download()
    assert np.array_equal(ga.pos.view(np.uint32), gb.pos.view(np.uint32))
    assert np.array_equal(ga.vel.view(np.uint32), gb.vel.view(np.uint32))
    assert a.dt == b.dt


def test_first_iteration_fixhash_path():
    p = Pipeline("dambreak", first=True)
    assert np.array_equal(host(p.g["hash"], np.uint32), p.o["hash"])
    assert np.array_equal(host(neibs_list_rows(p.g["nl"]), np.uint16), p.o["nl"])


def test_sort_with_ids_wider_than_30_bits():
    params, parts = get("lattice")
    n = parts.n
    info = parts.info.copy()
    rng = np.random.default_rng(1)
    ids = rng.permutation(n).astype(np.uint64) * 1000 + (1 << 31)      # need 32 bits
    info[:, 2] = (ids & 0xFFFF).astype(np.uint16)
    info[:, 3] = (ids >> 16).astype(np.uint16)
    hashv = parts.hash.copy()
    hashv[rng.choice(n, 50, replace=False)] = 0xFFFFFFFF                  # some inactive particles sort last
    fw = SimFramework(params, 0)
    g_hash, g_info = dev(hashv.view(np.int32)), dev(info.view(np.int16))
    g_pidx = dev(np.arange(n, dtype=np.int32))
    b = BufferList({BUFFER_HASH: g_hash, BUFFER_INFO: g_info, BUFFER_PARTINDEX: g_pidx})
    fw.neibsEngine.sort(b, b, n)
    pidx = np.arange(n, dtype=np.uint32)
    ob.sort(hashv, info, pidx)
    assert np.array_equal(host(g_hash, np.uint32), hashv)
    assert np.array_equal(host(g_info, np.uint16), info)
    assert np.array_equal(host(g_pidx, np.uint32), pidx)


def test_neighbour_list_overflow_reported_like_reference():
    params, parts = lattice_problem(10, jitter=0.1, neiblistsize=32)
    w = Worker(params, parts, 0, clobber=True)
    w.build_neibs()
    ref = ob.OracleWorker(params, parts)
    ref.build_neibs()
    gi, oi = w.last_neibs_info, ref.neibs_info
    assert gi.has_too_many_neibs >= 0 and oi.has_too_many_neibs >= 0
    assert gi.num_interactions == oi.num_interactions and gi.max_fluid_boundary_neibs == oi.max_fluid_boundary_neibs
    assert np.array_equal(host(neibs_list_rows(w.neibslist), np.uint16), ref.neibslist)     # truncated identically


def gpu_forces(pipe, from_=0, to=None):
    n = pipe.n
    to = n if to is None else to
    fw = pipe.fw
    f = torch.zeros((n, 4), dtype=torch.float32, device=DEV)
    cfl = torch.zeros(fw.forcesEngine.getFmaxElements(n) + 8, dtype=torch.float32, device=DEV)
    b = BufferList(pipe.sorted)
    b[BUFFER_FORCES] = f
    b[BUFFER_CFL] = cfl
    nb = fw.forcesEngine.basicstep(b, b, n, from_, to, 0)
    eos = torch.zeros((n, 2), dtype=torch.float32, device=DEV)
    fw.forcesEngine.eos_probe(b, eos, n)
    return f, cfl, nb, eos, b


def test_forces_parity(pipe):
    """forces + finalize + CFL vs the oracle. Tolerances, relative to the per-particle sum of |pair terms|
    (the natural scale of a cancelling float sum):
      * 5e-5 with the device's approximate-pow EOS values injected into the oracle (pure summation /
        contraction differences: ~75 float terms),
      * 5e-4 against the oracle's own powf EOS (the reference's __powf is ~1e-4 relative on P for
        rho~ ~ 1e-3, src/cuda/phys_core.cu:99-136)."""
    params, o = pipe.params, pipe.o
    f, cfl, nb, eos, b = gpu_forces(pipe)
    f, eos = host(f), host(eos)
    ptype = o["info"][:, 0] & 7
    for inject, tol in ((True, 5e-5), (False, 5e-4)):
        ep = np.ascontiguousarray(eos[:, 0]) if inject else None
        ec = np.ascontiguousarray(eos[:, 1]) if inject else None
        fo, cflo, ab = ob.forces(params, o["spos"], o["svel"], o["info"], o["hash"], o["cs"], o["nl"], ep, ec, want_abssum=True)
        sv = ab[:, 0] + 1e-3 * np.abs(fo[:, :3]).max() + 1e-12
        sw = ab[:, 3] / float(params.rho0[0]) + 1e-7 * np.abs(fo[:, 3]).max() + 1e-12
        ev = np.abs(f[:, :3] - fo[:, :3]).max(axis=1) / sv
        ew = np.abs(f[:, 3] - fo[:, 3]) / sw
        assert ev.max() < tol, f"inject={inject}: momentum error {ev.max():.3e}"
        assert ew.max() < tol, f"inject={inject}: continuity error {ew.max():.3e}"
        assert nb == cflo.shape[0]
        assert np.allclose(host(cfl)[:nb], cflo, rtol=1e-4 if inject else 1e-3)
    assert (f[ptype == 1, :3] == 0).all()
    # EOS probe itself vs exact powf: within the documented accuracy of __powf
    po, co = ob.eos(params, o["svel"], o["info"])
    assert np.allclose(eos[:, 1], co, rtol=1e-5)
    assert np.abs(eos[:, 0] - po).max() <= 3e-4 * np.abs(po).max() + 1e-9


def test_forces_partial_range_and_dtreduce(pipe):
    """basicstep on [from,to) touches only that range and numbers its CFL blocks from cflOffset
    (striping call pattern, src/GPUWorker.cc:2137-2148); dtreduce matches the oracle."""
    params, o, n = pipe.params, pipe.o, pipe.n
    if n < 600:
        pytest.skip("too small to split")
    f_all, cfl_all, nb_all, eos, b = gpu_forces(pipe)
    frm, to = 256, n - 100
    f = torch.zeros((n, 4), dtype=torch.float32, device=DEV)
    cfl = torch.full((pipe.fw.forcesEngine.getFmaxElements(n) + 8,), -1.0, dtype=torch.float32, device=DEV)
    b2 = BufferList(b)
    b2[BUFFER_FORCES], b2[BUFFER_CFL] = f, cfl
    nb = pipe.fw.forcesEngine.basicstep(b2, b2, n, frm, to, 4)
    assert nb == ((to - frm + 127) // 128 + 3) // 4 * 4
    fh = host(f)
    assert np.array_equal(fh[frm:to], host(f_all)[frm:to])
    assert (fh[:frm] == 0).all() and (fh[to:] == 0).all()
    c = host(cfl)
    assert (c[:4] == -1).all() and (c[4:4 + nb] >= 0).all() and (c[4 + nb:] == -1).all()
    assert np.array_equal(c[4:4 + 2], host(cfl_all)[2:4])        # blocks of 128 from particle 256 = blocks 2,3
    dt = pipe.fw.forcesEngine.dtreduce(b, b, nb_all)
    assert dt == pytest.approx(ob.dtreduce(params, host(cfl_all)[:nb_all]), rel=1e-6)


def test_euler_parity(pipe):
    params, o, n = pipe.params, pipe.o, pipe.n
    rng = np.random.default_rng(5)
    forces = rng.normal(0, 10, size=(n, 4)).astype(np.float32)
    dt = 1.3e-4
    for step, d in ((1, dt / 2), (2, dt)):
        po, vo = ob.euler(params, o["spos"], o["svel"], o["info"], o["hash"], forces, d, step)
        npos = torch.zeros((n, 4), dtype=torch.float32, device=DEV)
        nvel = torch.zeros((n, 4), dtype=torch.float32, device=DEV)
        rd = BufferList(pipe.sorted)
        rd[BUFFER_FORCES] = dev(forces)
        wr = BufferList({BUFFER_POS: npos, BUFFER_VEL: nvel})
        pipe.fw.integrationEngine.basicstep(rd, wr, n, n, d, step)
        # one or two FMAs per component: 2 ulp of the magnitudes involved
        assert np.allclose(host(npos), po, rtol=3e-7, atol=1e-10)
        assert np.allclose(host(nvel), vo, rtol=3e-7, atol=1e-9)
    with pytest.raises(ValueError):
        pipe.fw.integrationEngine.basicstep(rd, wr, n, n, dt, 3)          # reference throws too (euler.cu:361)


def test_missing_mandatory_buffer_raises(pipe):
    b = BufferList(pipe.sorted)
    del b[BUFFER_POS]
    with pytest.raises(ValueError):
        pipe.fw.neibsEngine.buildNeibsList(b, pipe.sorted, pipe.n, pipe.n)


@pytest.mark.parametrize("name", ["dambreak", "lattice"])
def test_time_stepping_tracks_oracle(name):
    """12 predictor-corrector steps (one neighbour rebuild in between) on GPU vs the oracle worker with the
    same dt sequence. Drift bound: 1e-4 of the velocity scale, 1e-5 dp on positions."""
    params, parts = get(name)
    w = Worker(params, parts, 0, clobber=True)
    ref = ob.OracleWorker(params, parts)
    for _ in range(12):
        dt = w.dt
        w.step()
        ref.step(dt=dt)
        assert w.dt == pytest.approx(ref.dt, rel=2e-3)
    got, exp = w.download(), ref.download()
    assert got.n == exp.n
    gp = global_positions(params, got.pos, got.hash)
    ep = global_positions(params, exp.pos, exp.hash)
    ids_g = (got.info[:, 3].astype(np.int64) << 16) | got.info[:, 2]
    ids_e = (exp.info[:, 3].astype(np.int64) << 16) | exp.info[:, 2]
    og, oe = np.argsort(ids_g), np.argsort(ids_e)
    assert np.array_equal(ids_g[og], ids_e[oe])
    assert np.abs(gp[og] - ep[oe]).max() < 1e-5 * float(params.deltap)
    vs = np.abs(exp.vel[:, :3]).max()
    assert np.abs(got.vel[og, :3] - exp.vel[oe, :3]).max() < 1e-4 * vs
    # rho~: the Molteni-Colagrossi term of the dam break is switched per pair on |P_i - P_j| < rho g dz with the raw
    # pressures (forces_kernel.def:1925-1928). P = B ((rho~ + 1)^gamma - 1) with B = 5.7e7 amplifies the last bits of the
    # power - the GPU's __powf vs the CPU's powf - to ~50 Pa, the size of the threshold itself: borderline pairs fall on
    # different sides in the two implementations (as they do between any two builds of the reference), which moves
    # rho~ of a few particles by ~1e-5 over 12 steps. The lattice case (no diffusion) keeps 1e-6.
    assert np.abs(got.vel[og, 3] - exp.vel[oe, 3]).max() < (2e-5 if name == "dambreak" else 1e-6)


def test_xzy_linearisation_tracks_oracle():
    """The cell linearisation bench.py uses on N > 1 GPUs (xzy: y slowest) on ONE GPU against the oracle: list bit-exact,
    12 steps within the drift bounds of test_time_stepping_tracks_oracle."""
    params, parts = dambreak_problem(0.03, densitydiffusion=capi.RHODIFF_FERRARI, density_diff_coeff=0.1, coord=(0, 2, 1))
    assert list(params.coord) == [0, 2, 1]
    w = Worker(params, parts, 0, clobber=True)
    ref = ob.OracleWorker(params, parts)
    w.build_neibs(); ref.build_neibs()
    n = w.numParticles
    assert np.array_equal(host(w.hash[:n], np.uint32), ref.download().hash)
    assert np.array_equal(host(neibs_list_rows(w.neibslist), np.uint16)[:, :n], ref.neibslist[:, :n])
    for _ in range(12):
        dt = w.dt
        w.step()
        ref.step(dt=dt)
    got, exp = w.download(), ref.download()
    gp, ep = global_positions(params, got.pos, got.hash), global_positions(params, exp.pos, exp.hash)
    ids_g = (got.info[:, 3].astype(np.int64) << 16) | got.info[:, 2]
    ids_e = (exp.info[:, 3].astype(np.int64) << 16) | exp.info[:, 2]
    og, oe = np.argsort(ids_g), np.argsort(ids_e)
    assert np.array_equal(ids_g[og], ids_e[oe])
    assert np.abs(gp[og] - ep[oe]).max() < 1e-5 * float(params.deltap)
    assert np.abs(got.vel[og, :3] - exp.vel[oe, :3]).max() < 1e-4 * np.abs(exp.vel[:, :3]).max()
    assert np.abs(got.vel[og, 3] - exp.vel[oe, 3]).max() < 2e-5


@pytest.mark.timeout(900)
def test_list_layout_independence_at_benchmark_size():
    """DamBreak3D at the north-star size (--deltap 0.0026: 7.87 M particles, BASELINE configs' headline): the default list layout (blocks
    of 2 M particles: three full blocks and a narrow one) against the reference's interleaved layout (one block): the same
    list values, and bitwise the same state after two steps (rebuild, four force evaluations through the list)."""
    params, parts = dambreak_problem(0.0026, densitydiffusion=capi.RHODIFF_FERRARI, density_diff_coeff=0.1)
    assert parts.n > 3 * capi.NEIBLIST_BLOCK
    pi = params.copy()
    pi.neiblist_block = 1 << 23                   # > allocated: the whole allocation is one block = the reference's layout
    a = Worker(params, parts, 0)
    a.build_neibs()
    rows_a = neibs_list_rows(a.neibslist)
    info_a = a.last_neibs_info
    for _ in range(2):
        a.step()
    got_a = (a.pos[a.cur][:a.numParticles].clone(), a.vel[a.cur][:a.numParticles].clone())
    n = a.numParticles
    del a
    torch.cuda.empty_cache()
    b = Worker(pi, parts, 0)
    b.build_neibs()
    assert b.last_neibs_info.num_interactions == info_a.num_interactions
    assert b.last_neibs_info.max_fluid_boundary_neibs == info_a.max_fluid_boundary_neibs
    assert torch.equal(rows_a, b.neibslist), "one block: the buffer IS the [rows, allocated] array of the reference"
    del rows_a
    for _ in range(2):
        b.step()
    assert torch.equal(got_a[0].view(torch.int32), b.pos[b.cur][:n].view(torch.int32))
    assert torch.equal(got_a[1].view(torch.int32), b.vel[b.cur][:n].view(torch.int32))


def test_full_size_properties():
    """Size-independent properties at benchmark scale (2M lattice; the oracle is too slow there):
    sortedness, cell ranges partition the particles, neighbour relation is symmetric, and pairwise
    antisymmetry of the momentum terms (sum m a = sum m g for a boundary-free periodic-less fluid block)."""
    params, parts = lattice_problem(126, jitter=0.05)
    n = parts.n
    w = Worker(params, parts, 0, clobber=False)
    w.build_neibs()
    hashv = w.hash[:n].cpu().numpy().view(np.uint32).astype(np.int64)
    info = w.info[:n].cpu().numpy().view(np.uint16)
    ids = (info[:, 3].astype(np.int64) << 16) | info[:, 2]
    key = (hashv << 32) | ids
    assert (np.diff(key) > 0).all()
    assert np.array_equal(np.sort(ids), np.arange(n))
    cs = w.cellstart.cpu().numpy().view(np.uint32)
    ce = w.cellend.cpu().numpy().view(np.uint32)
    used = cs != 0xFFFFFFFF
    counts = np.bincount(hashv, minlength=params.num_cells)
    assert np.array_equal((ce[used] - cs[used]).astype(np.int64), counts[used]) and (counts[~used] == 0).all()
    gi = w.last_neibs_info
    assert gi.has_too_many_neibs == -1 and gi.max_fluid_boundary_neibs < 127
    # list entry count == sum of per-particle counts; symmetric relation => even total
    nl = neibs_list_rows(w.neibslist)
    cnt = (nl[:, :n] != -1).to(torch.int32)
    first_end = torch.argmax((nl[:, :n] == -1).to(torch.int8), dim=0)
    assert int(first_end.sum().item()) == gi.num_interactions
    assert gi.num_interactions % 2 == 0
    # momentum: sum_i m_i (a_i - g) ~ 0 relative to sum_i m_i |a_i - g|
    w.forces.basicstep(w.state(w.cur), w.state(w.cur), n, 0, n, 0)
    f = w.forces_buf[:n].double()
    g = torch.tensor([params.gravity[a] for a in range(3)], dtype=torch.float64, device=DEV)
    a = f[:, :3] - g
    tot = a.sum(dim=0).abs().max().item()
    scale = a.abs().sum().item()
    assert tot < 1e-6 * scale
    # continuity is symmetric for equal masses: sum_i drho_i over an isolated block need not vanish, but it is finite
    assert torch.isfinite(f).all()


def test_poiseuille_steady_profile_is_preserved():
    """Physics validation restated from scripts/validate-poiseuille.py:32-37,95-121 (which needs ParaView): started from
    the analytic steady profile v_x(z) = F/(2 nu)((lz/2)^2 - z^2), the flow must stay on it. L-inf error of the fluid
    velocity against the analytic profile after 300 steps, relative to the peak velocity: < 2 % at ppH = 16."""
    params, parts = poiseuille_problem(16)
    n_fluid = int(((parts.info[:, 0] & 7) == 0).sum())
    w = Worker(params, parts, 0)
    for _ in range(300):
        w.step()
    out = w.download()
    gp = global_positions(params, out.pos, out.hash)
    fl = (out.info[:, 0] & 7) == 0
    assert fl.sum() == n_fluid
    F, nu, lz = float(params.gravity[0]), float(params.visccoeff[0]), 1.0
    exact = F / (2 * nu) * ((lz / 2) ** 2 - gp[fl, 2] ** 2)
    vmax = F / (2 * nu) * (lz / 2) ** 2
    err = np.abs(out.vel[fl, 0] - exact).max() / vmax
    assert err < 0.02, f"L-inf error {err:.3%}"
    assert np.abs(out.vel[fl, 1:3]).max() < 0.01 * vmax


def test_device_resident_dt_matches_host_dt_path_bitwise():
    """The b200sph_step_* entry points (dt kept on the device, nothing read back during a step) must reproduce the
    reference-style host-dt call sequence bit for bit, including the adaptive dt sequence and the simulated time."""
    params, parts = get("dambreak")
    a = Worker(params, parts, 0, device_dt=True)
    b = Worker(params, parts, 0, device_dt=False)
    for _ in range(13):
        a.step()
        b.step()
    assert a.dt == b.dt and a.t == pytest.approx(b.t, rel=1e-15)
    ga, gb = a.download(), b.download()
    assert np.array_equal(ga.hash, gb.hash) and np.array_equal(ga.info, gb.info)
    assert np.array_equal(ga.pos.view(np.uint32), gb.pos.view(np.uint32))
    assert np.array_equal(ga.vel.view(np.uint32), gb.vel.view(np.uint32))


def test_force_feedback_body_parity():
    """Row f1: forces with compute_object_forces (force x mass + torque scatter), reduceRbForces (segmented scan) and the
    rigid motion of moving-body particles in euler, against the oracle."""
    from test_oracle_cpu import body_setup, prepared
    from gpusph_b200.engines import BUFFER_RB_FORCES, BUFFER_RB_KEYS, BUFFER_RB_TORQUES
    from gpusph_b200.problems import ParticleArrays
    params, parts = dambreak_problem(0.03, obstacle=True, densitydiffusion=capi.RHODIFF_FERRARI, density_diff_coeff=0.1)
    rng = np.random.default_rng(3)
    fl = (parts.info[:, 0] & 7) == 0
    parts.vel[:, :3] += rng.normal(0, 0.5, size=(parts.n, 3)).astype(np.float32) * fl[:, None]
    parts.vel[:, 3] += rng.normal(0, 2e-3, size=parts.n).astype(np.float32)
    spos, svel, info, hashv, pidx, cs, ce, newn = prepared(params, parts)
    n = parts.n
    b, isb, cg, nbody, first_id = body_setup(params, ParticleArrays(spos, svel, info, hashv))
    nl, _ = ob.build_neibs(params, spos, info, hashv, cs, ce)
    fw = SimFramework(params, 0)
    cgg = [[b.cgGridPos[o][a] for a in range(3)] for o in range(2)]
    cgp = [[b.cgPos[o][a] for a in range(3)] for o in range(2)]
    fw.forcesEngine.setrbcg(cgg, cgp, 2)
    fw.forcesEngine.setrbstart([0, b.startIndex[1]], 2)
    g = BufferList({BUFFER_POS: dev(spos), BUFFER_VEL: dev(svel), BUFFER_INFO: dev(info.view(np.int16)),
                    BUFFER_HASH: dev(hashv.view(np.int32)), BUFFER_CELLSTART: dev(cs.view(np.int32)),
                    BUFFER_NEIBSLIST: neibs_list_blocked(dev(nl.view(np.int16))).contiguous(),   # oracle list -> the engines' layout
                    BUFFER_FORCES: torch.zeros((n, 4), dtype=torch.float32, device=DEV),
                    BUFFER_CFL: torch.zeros(fw.forcesEngine.getFmaxElements(n), dtype=torch.float32, device=DEV),
                    BUFFER_RB_FORCES: torch.zeros((nbody, 4), dtype=torch.float32, device=DEV),
                    BUFFER_RB_TORQUES: torch.zeros((nbody, 4), dtype=torch.float32, device=DEV),
                    BUFFER_RB_KEYS: torch.ones(nbody, dtype=torch.int32, device=DEV)})
    fw.forcesEngine.basicstep(g, g, n, 0, n, 0, compute_object_forces=True)
    eos = torch.zeros((n, 2), dtype=torch.float32, device=DEV)
    fw.forcesEngine.eos_probe(g, eos, n)
    e = host(eos)
    rbf = np.zeros((nbody, 4), dtype=np.float32)
    rbt = np.zeros((nbody, 4), dtype=np.float32)
    fo, _, ab = ob.forces(params, spos, svel, info, hashv, cs, nl, np.ascontiguousarray(e[:, 0]), np.ascontiguousarray(e[:, 1]),
                          want_abssum=True, bodies=b, rb_forces=rbf, rb_torques=rbt)
    ids = (info[:, 3].astype(np.int64) << 16) | info[:, 2]
    k = ids[isb] - first_id
    m = spos[isb, 3]
    scale = (ab[isb, 0] * m)[:, None] + 1e-3 * np.abs(rbf).max() + 1e-12
    got_f, got_t = host(g[BUFFER_RB_FORCES]), host(g[BUFFER_RB_TORQUES])
    assert (np.abs(got_f[k, :3] - rbf[k, :3]) < 5e-5 * scale).all()
    assert np.abs(got_t[k, :3] - rbt[k, :3]).max() < 5e-5 * (np.abs(rbt).max() + 1e-9) + 1e-4 * 0.3 * scale.max()
    assert np.allclose(host(g[BUFFER_FORCES])[isb, :3], got_f[k, :3])           # FORCES holds the scaled force too
    # segmented reduction: totals of body 1 (single key) = plain sums
    tot_f, tot_t = fw.forcesEngine.reduceRbForces(g, [nbody - 1], 1, nbody)
    assert np.allclose(tot_f[0], got_f[:, :3].astype(np.float64).sum(axis=0), rtol=1e-4, atol=1e-4 * np.abs(got_f).max())
    assert np.allclose(tot_t[0], got_t[:, :3].astype(np.float64).sum(axis=0), rtol=1e-4, atol=1e-4 * np.abs(got_t).max())
    # rigid motion in euler
    th = 0.01
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], dtype=np.float32)
    tr, lv, om = [1e-3, -2e-3, 5e-4], [0.1, 0.0, -0.05], [0.0, 0.2, 1.0]
    for i in range(9):
        b.steprot[1][i] = float(R.ravel()[i])
    for a in range(3):
        b.trans[1][a], b.linearvel[1][a], b.angularvel[1][a] = tr[a], lv[a], om[a]
    ie = fw.integrationEngine
    ie.setrbcg(cgg, cgp, 2)
    ie.setrbtrans([[0, 0, 0], tr], 2)
    ie.setrbsteprot([np.eye(3).ravel(), R.ravel()], 2)
    ie.setrblinearvel([[0, 0, 0], lv], 2)
    ie.setrbangularvel([[0, 0, 0], om], 2)
    po, vo = ob.euler(params, spos, svel, info, hashv, fo, 1e-4, 2, bodies=b)
    npos = torch.zeros((n, 4), dtype=torch.float32, device=DEV)
    nvel = torch.zeros((n, 4), dtype=torch.float32, device=DEV)
    rd = BufferList(g)
    rd[BUFFER_FORCES] = dev(fo)
    ie.basicstep(rd, BufferList({BUFFER_POS: npos, BUFFER_VEL: nvel}), n, n, 1e-4, 2)
    assert np.allclose(host(npos), po, rtol=3e-7, atol=2e-9)
    assert np.allclose(host(nvel), vo, rtol=3e-7, atol=2e-8)


def test_empty_and_single_particle_inputs():
    """Edge cases: zero particles is a no-op for every entry point; a single particle gets an empty neighbour list,
    only gravity as acceleration and moves ballistically."""
    params, parts = lattice_problem(1, jitter=0.0)
    fw = SimFramework(params, 0)
    e = BufferList({k: torch.zeros((1, 4), dtype=torch.float32, device=DEV) for k in (BUFFER_POS, BUFFER_VEL, BUFFER_FORCES)})
    e.update({BUFFER_INFO: torch.zeros((1, 4), dtype=torch.int16, device=DEV), BUFFER_HASH: torch.zeros(1, dtype=torch.int32, device=DEV),
              BUFFER_PARTINDEX: torch.zeros(1, dtype=torch.int32, device=DEV),
              BUFFER_CELLSTART: torch.full((params.num_cells,), -1, dtype=torch.int32, device=DEV),
              BUFFER_CELLEND: torch.full((params.num_cells,), -1, dtype=torch.int32, device=DEV),
              BUFFER_NEIBSLIST: torch.full((int(params.neiblistsize), 1), -1, dtype=torch.int16, device=DEV),
              BUFFER_CFL: torch.zeros(8, dtype=torch.float32, device=DEV)})
    fw.neibsEngine.calcHash(e, e, 0)
    fw.neibsEngine.sort(e, e, 0)
    fw.neibsEngine.buildNeibsList(e, e, 0, 0)
    assert fw.forcesEngine.basicstep(e, e, 0, 0, 0, 0) == 0
    fw.integrationEngine.basicstep(e, e, 0, 0, 1e-4, 1)
    w = Worker(params, parts, 0, clobber=True)
    w.step()
    assert w.last_neibs_info.num_interactions == 0
    f = w.forces_buf[:1].cpu().numpy()
    assert np.allclose(f[0, :3], [0, 0, -9.81]) and f[0, 3] == 0
    out = w.download()
    assert out.n == 1 and np.isfinite(out.pos).all()
    assert out.vel[0, 2] == pytest.approx(parts.vel[0, 2] - 9.81 * w.t, rel=1e-5)
