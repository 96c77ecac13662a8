"""Pin the CPU oracle (oracle/sph_oracle.c) against INDEPENDENT restatements: brute-force O(N^2)
neighbour search and the SPH formulas written directly in float64 numpy from the published
equations. The reference ships no golden vectors; these checks plus the
reference-generated fixtures (test_golden.py) are what the oracle is anchored to."""
import numpy as np
import pytest

import oracle_binding as ob
from gpusph_b200 import capi
from gpusph_b200.problems import (dambreak_problem, global_positions, initial_dt, lattice_problem,
                                  localpos_and_hash)

CELLMASK = 0x3FFFFFFF


def prepared(params, parts, move=0.0, seed=1):
    """Run the oracle's neighbour pipeline once; optionally displace particles first."""
    pos, vel, info, hashv = parts.pos.copy(), parts.vel.copy(), parts.info.copy(), parts.hash.copy()
    if move:
        rng = np.random.default_rng(seed)
        pos[:, :3] += rng.uniform(-move, move, size=(parts.n, 3)).astype(np.float32)
        pidx = ob.calc_hash(params, pos, hashv, info)
    else:
        pidx = ob.fix_hash(params, hashv, info)
    ob.sort(hashv, info, pidx)
    cs, ce, _, spos, svel, newn = ob.reorder(params, pos, vel, info, hashv, pidx)
    return spos, svel, info, hashv, pidx, cs, ce, newn


def decode_list(params, nl, hashv, cs, i):
    """neighbour indices of particle i in list order: (fluid section, boundary section)."""
    G = [int(params.grid_size[a]) for a in range(3)]
    c = [params.coord[a] for a in range(3)]
    cell = int(hashv[i] & CELLMASK)
    t = G[c[1]] * G[c[0]]
    g = [0, 0, 0]
    g[c[2]] = cell // t
    r = cell - g[c[2]] * t
    g[c[1]] = r // G[c[0]]
    g[c[0]] = r - g[c[1]] * G[c[0]]
    out = []
    for start, step in ((0, 1), (int(params.neibboundpos), -1)):
        sec, base, slot = [], 0, start
        while True:
            nd = int(nl[slot, i])
            if nd == 0xFFFF:
                break
            if nd >= 2048:
                cn = (nd >> 11) - 1
                nd &= 2047
                o = [cn % 3 - 1, (cn // 3) % 3 - 1, cn // 9 - 1]
                ng = [(g[a] + o[a]) % G[a] for a in range(3)]
                base = int(cs[ng[c[2]] * G[c[1]] * G[c[0]] + ng[c[1]] * G[c[0]] + ng[c[0]]])
            sec.append(base + nd)
            slot += step
        out.append(sec)
    return out


def test_calc_hash_matches_independent_double_computation():
    params, parts = lattice_problem(10, jitter=0.3)
    pos, hashv, info = parts.pos.copy(), parts.hash.copy(), parts.info.copy()
    rng = np.random.default_rng(0)
    pos[:, :3] += rng.uniform(-0.02, 0.02, size=(parts.n, 3)).astype(np.float32)   # up to ~0.77 cell
    gpos = global_positions(params, pos, hashv)
    ob.calc_hash(params, pos, hashv, info)
    # independent: hash of the displaced global position
    _, h2 = localpos_and_hash(params, gpos, pos[:, 3])
    cs = np.array([params.cell_size[a] for a in range(3)])
    frac = np.abs(np.abs(pos[:, :3] / cs) - 0.5)
    clear = (frac > 1e-4).all(axis=1)            # not within rounding distance of a cell face
    assert clear.mean() > 0.99
    assert np.array_equal(hashv[clear], h2[clear])
    # local positions stay inside the cell and reproduce the same global position
    assert (np.abs(pos[:, :3]) <= cs * 0.5 * (1 + 1e-6)).all()
    assert np.allclose(global_positions(params, pos, hashv), gpos, atol=2e-7)


def test_calc_hash_disables_particles_that_flew_too_far_and_handles_periodicity():
    params, parts = lattice_problem(6, jitter=0.0)
    pos, hashv, info = parts.pos.copy(), parts.hash.copy(), parts.info.copy()
    cs = params.cell_size[0]
    g0 = global_positions(params, pos, hashv)
    i0 = int(np.argmin(g0[:, 0]))      # a particle of the first fluid layer: cell x = 1 (one padding cell)
    pos[i0, 0] -= np.float32(1.0 * cs)  # -> edge cell x = 0
    pos[1, 0] += np.float32(1.2 * cs)   # plain migration by one cell
    h1 = hashv[1]
    ob.calc_hash(params, pos, hashv, info)
    assert hashv[1] != h1 and np.isfinite(pos[1, 3]) and np.isfinite(pos[i0, 3])
    # from the edge cell, flying out by more than one cell disables the particle:
    # mass = NaN, hash = CELL_HASH_MAX (buildneibs_kernel.cu:288-299, 754-760)
    pos[i0, 0] -= np.float32(2.5 * cs)
    ob.calc_hash(params, pos, hashv, info)
    assert np.isnan(pos[i0, 3]) and hashv[i0] == 0xFFFFFFFF
    # periodic in x: a particle leaving through the low face re-enters in the last cell
    pp, parts2 = lattice_problem(6, jitter=0.0, periodic=capi.PERIODIC_X)
    pos, hashv, info = parts2.pos.copy(), parts2.hash.copy(), parts2.info.copy()
    g0 = global_positions(pp, pos, hashv)
    i = int(np.argmin(g0[:, 0]))
    shift = (g0[i, 0] - pp.world_origin[0]) + 0.3 * pp.cell_size[0]
    pos[i, 0] -= np.float32(shift)
    ob.calc_hash(pp, pos, hashv, info)
    g1 = global_positions(pp, pos, hashv)
    assert np.isfinite(pos[i, 3])
    assert g1[i, 0] > pp.world_origin[0] + (pp.grid_size[0] - 1) * pp.cell_size[0]


def test_sort_is_the_reference_total_order():
    params, parts = dambreak_problem(0.05)
    rng = np.random.default_rng(3)
    perm = rng.permutation(parts.n)
    hashv, info = parts.hash[perm].copy(), parts.info[perm].copy()
    pidx = np.arange(parts.n, dtype=np.uint32)
    ob.sort(hashv, info, pidx)
    ptype = (info[:, 0] & 7).astype(np.int64)
    ids = info[:, 2].astype(np.int64) | (info[:, 3].astype(np.int64) << 16)
    key = (hashv.astype(np.int64) << 34) | (ptype << 32) | ids
    assert (np.diff(key) > 0).all()
    # and it is a permutation of the input carried consistently
    assert np.array_equal(parts.hash[perm][pidx], hashv)
    assert np.array_equal(parts.info[perm][pidx], info)


def test_cellstart_cellend_partition_the_sorted_particles():
    params, parts = dambreak_problem(0.05)
    spos, svel, info, hashv, pidx, cs, ce, newn = prepared(params, parts)
    assert newn == parts.n
    used = cs != 0xFFFFFFFF
    assert (ce[used] > cs[used]).all()
    counts = np.bincount(hashv & CELLMASK, minlength=params.num_cells)
    assert np.array_equal((ce[used] - cs[used]).astype(np.int64), counts[used])
    assert (counts[~used] == 0).all()
    for c in np.flatnonzero(used)[::7]:
        assert ((hashv[cs[c]:ce[c]] & CELLMASK) == c).all()
    assert np.array_equal(spos, parts.pos[pidx]) and np.array_equal(svel, parts.vel[pidx])


@pytest.mark.parametrize("problem", ["lattice", "dambreak", "periodic"])
def test_neighbour_list_against_brute_force(problem):
    if problem == "lattice":
        params, parts = lattice_problem(9, jitter=0.2)
    elif problem == "periodic":
        params, parts = lattice_problem(8, jitter=0.2, periodic=capi.PERIODIC_X | capi.PERIODIC_Y)
    else:
        params, parts = dambreak_problem(0.06)
    spos, svel, info, hashv, pidx, cs, ce, newn = prepared(params, parts, move=0.002)
    nl, ninfo = ob.build_neibs(params, spos, info, hashv, cs, ce)
    n = spos.shape[0]
    g = global_positions(params, spos, hashv)
    size = np.array([params.grid_size[a] * params.cell_size[a] for a in range(3)], dtype=np.float64)
    ptype = info[:, 0] & 7
    R2 = float(params.nl_sq_influence_radius)
    total = 0
    maxfb = 0
    rng = np.random.default_rng(5)
    sample = rng.choice(n, size=min(n, 400), replace=False)
    for i in range(n):
        fl, bd = decode_list(params, nl, hashv, cs, i) if i in set(sample.tolist()) else (None, None)
        if fl is None:
            continue
        d = g - g[i]
        for a in range(3):
            if params.periodic & (1 << a):
                d[:, a] -= size[a] * np.round(d[:, a] / size[a])
        r2 = (d * d).sum(axis=1)
        inside = r2 < R2 * (1 - 1e-5)
        outside = r2 > R2 * (1 + 1e-5)
        inside[i] = False
        if ptype[i] == 1:            # DYN: boundary-boundary pairs are skipped
            inside &= ptype != 1
        got = set(fl) | set(bd)
        must = set(np.flatnonzero(inside).tolist())
        mustnot = set(np.flatnonzero(outside).tolist())
        assert must <= got, f"particle {i}: missing neighbours"
        assert not (got & mustnot), f"particle {i}: spurious neighbours"
        assert all(ptype[j] == 0 for j in fl) and all(ptype[j] == 1 for j in bd)
        assert len(set(fl)) == len(fl) and len(set(bd)) == len(bd)
    # counters: recount from the list itself
    for i in range(n):
        nf = int(np.argmax(nl[:, i] == 0xFFFF))
        col = nl[::-1, i][params.neiblistsize - 1 - params.neibboundpos:]
        nb = int(np.argmax(col == 0xFFFF))
        total += nf + nb
        maxfb = max(maxfb, nf + nb)
    assert ninfo.num_interactions == total
    assert ninfo.max_fluid_boundary_neibs == maxfb
    assert ninfo.has_too_many_neibs == -1


def test_neighbour_list_overflow_is_truncated_and_reported():
    params, parts = lattice_problem(8, jitter=0.1, neiblistsize=32)
    spos, svel, info, hashv, pidx, cs, ce, newn = prepared(params, parts)
    nl, ninfo = ob.build_neibs(params, spos, info, hashv, cs, ce)
    assert ninfo.has_too_many_neibs >= 0
    assert ninfo.has_max_neibs[0] >= params.neibboundpos
    # every column still terminates inside the list
    assert ((nl == 0xFFFF).sum(axis=0) >= 1).all()
    assert ninfo.max_fluid_boundary_neibs > 31


def numpy_forces(params, pos, vel, info, hashv, diffusion):
    """Independent float64 evaluation of the WCSPH right-hand side (Monaghan 1992/1994 continuity +
    momentum with artificial viscosity; Wendland C2 kernel; Tait EOS; Molteni-Colagrossi / Ferrari
    density diffusion) with brute-force neighbours. Returns Dv/Dt (before gravity) and D(rho~)/Dt."""
    n = pos.shape[0]
    g = global_positions(params, pos, hashv)
    v = vel[:, :3].astype(np.float64)
    m = pos[:, 3].astype(np.float64)
    rho0, B, gam, c0 = (float(params.rho0[0]), float(params.bcoeff[0]), float(params.gammacoeff[0]), float(params.sscoeff[0]))
    rt = vel[:, 3].astype(np.float64)
    rho = (rt + 1) * rho0
    P = B * ((rt + 1) ** gam - 1)
    c = c0 * (rt + 1) ** ((gam - 1) / 2)
    h = float(params.slength)
    fc = 105.0 / (128.0 * np.pi * h ** 5)
    R = float(params.influenceradius)
    grav = np.array([params.gravity[a] for a in range(3)], dtype=np.float64)
    ptype = info[:, 0] & 7
    acc = np.zeros((n, 3))
    drho = np.zeros(n)
    for i in range(n):
        if ptype[i] > 1:
            continue
        d = g[i] - g
        r = np.sqrt((d * d).sum(axis=1))
        mask = (r < R) & (r > 0) & (ptype <= 1)
        if ptype[i] == 1:
            mask &= ptype == 0
        j = np.flatnonzero(mask)
        rij, rr = d[j], r[j]
        F = fc * (rr / h - 2) ** 3
        vij = v[i] - v[j]
        vr = (vij * rij).sum(axis=1)
        dr = m[j] * vr * F
        fl = ptype[j] == 0
        if diffusion == capi.RHODIFF_COLAGROSSI:
            cond = ~(np.abs(P[i] - P[j]) < np.abs((rij @ grav) * rho[i]))
            dr -= np.where(fl & cond, float(params.density_diff_coeff) * c0 * (rho[j] / rho[i] - 1) * F * m[j], 0.0)
        elif diffusion == capi.RHODIFF_FERRARI:
            gc = -(rij @ grav) * rho0 / c0 ** 2
            dr += np.where(fl, float(params.density_diff_coeff) * m[j] * np.maximum(c[i], c[j]) * (rho[i] - rho[j] + gc) / rho[i] / rr * rr * rr * F, 0.0)
        drho[i] = dr.sum() / rho0
        pterm = -(P[i] / rho[i] ** 2 + P[j] / rho[j] ** 2) * m[j] * F
        av = np.where(vr < 0, vr * h * float(params.artvisccoeff) * (c[i] + c[j]) / ((rr * rr + float(params.epsartvisc)) * (rho[i] + rho[j])), 0.0)
        a = ((pterm + av * m[j] * F)[:, None] * rij).sum(axis=0)
        if ptype[i] == 0:
            acc[i] = a + grav
    return acc, drho


@pytest.mark.parametrize("diffusion", [capi.RHODIFF_NONE, capi.RHODIFF_COLAGROSSI, capi.RHODIFF_FERRARI])
def test_forces_against_numpy_float64(diffusion):
    kw = dict(density_diff_coeff=0.1) if diffusion == capi.RHODIFF_FERRARI else {}
    params, parts = dambreak_problem(0.06, densitydiffusion=diffusion, **kw)
    # give the fluid some velocity and density variation so every term is exercised
    rng = np.random.default_rng(7)
    parts.vel[:, :3] += rng.normal(0, 0.5, size=(parts.n, 3)).astype(np.float32) * ((parts.info[:, 0] & 7) == 0)[:, None]
    parts.vel[:, 3] += rng.normal(0, 2e-3, size=parts.n).astype(np.float32)
    spos, svel, info, hashv, pidx, cs, ce, newn = prepared(params, parts)
    nl, _ = ob.build_neibs(params, spos, info, hashv, cs, ce)
    f, cfl, ab = ob.forces(params, spos, svel, info, hashv, cs, nl, want_abssum=True)
    acc, drho = numpy_forces(params, spos, svel, info, hashv, diffusion)
    ptype = info[:, 0] & 7
    scale_v = ab[:, 0] + 1e-3 * np.abs(f[:, :3]).max()
    scale_w = ab[:, 3] / float(params.rho0[0]) + 1e-6
    # float32 oracle vs float64 numpy: 5e-5 of the summed magnitudes (float rounding of ~70 terms and of
    # the cell-local -> global position conversion)
    fluid = ptype == 0
    assert (np.abs(f[fluid, :3] - acc[fluid]).max(axis=1) <= 5e-5 * scale_v[fluid]).all()
    assert (np.abs(f[:, 3] - drho) <= 5e-5 * scale_w).all()
    # boundary particles without force feedback get no acceleration
    assert (f[ptype == 1, :3] == 0).all()
    # CFL term: max over fluid particles per 128-block of max(|a|, c^2/h)
    rt = svel[:, 3].astype(np.float64)
    c = float(params.sscoeff[0]) * (rt + 1) ** 3
    term = np.where(fluid, np.maximum(np.linalg.norm(acc, axis=1), c * c / float(params.slength)), 0.0)
    nb = cfl.shape[0]
    padded = np.zeros(nb * 128)
    padded[:term.shape[0]] = term
    assert np.allclose(cfl, padded.reshape(nb, 128).max(axis=1), rtol=1e-4)
    dt = ob.dtreduce(params, cfl)
    expect = 0.3 * min(np.sqrt(float(params.slength) / cfl.max()), float(params.slength) / (1.1 * 20.0))
    assert dt == pytest.approx(expect, rel=1e-6)


def test_euler_predictor_corrector_formulas():
    params, parts = dambreak_problem(0.08)
    rng = np.random.default_rng(11)
    n = parts.n
    f = rng.normal(0, 5, size=(n, 4)).astype(np.float32)
    parts.vel[:, :3] = rng.normal(0, 1, size=(n, 3)).astype(np.float32)
    dt = np.float32(1e-4)
    p1, v1 = ob.euler(params, parts.pos, parts.vel, parts.info, parts.hash, f, float(dt / 2), 1)
    p2, v2 = ob.euler(params, parts.pos, parts.vel, parts.info, parts.hash, f, float(dt), 2)
    fl = (parts.info[:, 0] & 7) == 0
    bd = ~fl
    hdt = dt / 2
    assert np.allclose(p1[fl, :3], parts.pos[fl, :3] + parts.vel[fl, :3] * hdt, rtol=0, atol=1e-9)
    assert np.allclose(v1[fl], parts.vel[fl] + f[fl] * hdt, rtol=1e-6, atol=1e-9)
    velc = parts.vel[fl, :3] + f[fl, :3] * hdt
    assert np.allclose(p2[fl, :3], parts.pos[fl, :3] + velc * dt, rtol=0, atol=1e-9)
    assert np.allclose(v2[fl], parts.vel[fl] + f[fl] * dt, rtol=1e-6, atol=1e-9)
    # DYN boundary particles: only the density evolves
    assert np.array_equal(p2[bd], parts.pos[bd]) and np.array_equal(v2[bd, :3], parts.vel[bd, :3])
    assert np.allclose(v2[bd, 3], parts.vel[bd, 3] + dt * f[bd, 3], rtol=1e-6, atol=1e-9)
    # mass never changes
    assert np.array_equal(p2[:, 3], parts.pos[:, 3])


def test_oracle_worker_conserves_particles_and_stays_finite():
    params, parts = dambreak_problem(0.05)
    w = ob.OracleWorker(params, parts)
    assert w.dt == pytest.approx(initial_dt(params))
    for _ in range(12):        # crosses one neighbour-list rebuild
        w.step()
    out = w.download()
    assert out.n == parts.n
    assert np.isfinite(out.pos).all() and np.isfinite(out.vel).all()
    assert sorted(((out.info[:, 3].astype(np.int64) << 16) | out.info[:, 2]).tolist()) == list(range(parts.n))
    assert 0 < w.dt < 1e-3


def body_setup(params, parts):
    """Body record for the obstacle of dambreak_problem(obstacle=True): object 1, centre of gravity = mean position."""
    import oracle_binding as ob
    flags = parts.info[:, 0]
    isb = (flags & capi.FG_COMPUTE_FORCE) != 0
    ids = (parts.info[:, 3].astype(np.int64) << 16) | parts.info[:, 2]
    gp = global_positions(params, parts.pos, parts.hash)
    cg = gp[isb].mean(axis=0)
    cs = np.array([params.cell_size[a] for a in range(3)], dtype=np.float64)
    org = np.array([params.world_origin[a] for a in range(3)], dtype=np.float64)
    cgcell = np.floor((cg - org) / cs).astype(np.int32)
    cgloc = (cg - org - (cgcell + 0.5) * cs).astype(np.float32)
    b = ob.OracleBodies()
    for a in range(3):
        b.cgGridPos[1][a] = int(cgcell[a])
        b.cgPos[1][a] = float(cgloc[a])
    first_id = int(ids[isb].min())
    b.startIndex[1] = -first_id          # rbindex = id + startIndex -> 0 .. nbody-1
    return b, isb, cg, int(isb.sum()), first_id


def test_body_forces_torques_and_rigid_motion():
    """Force-feedback bodies: finalize scatters force x mass and torque about the centre of gravity
    (forces_kernel.def:4116-4141); euler moves body particles rigidly (euler_kernel.def:470-503). Checked against
    float64 numpy from global positions."""
    params, parts = dambreak_problem(0.04, obstacle=True)
    rng = np.random.default_rng(3)
    fl = (parts.info[:, 0] & 7) == 0
    parts.vel[:, :3] += rng.normal(0, 0.5, size=(parts.n, 3)).astype(np.float32) * fl[:, None]
    parts.vel[:, 3] += rng.normal(0, 2e-3, size=parts.n).astype(np.float32)
    # move the column next to the obstacle so that it feels a force
    spos, svel, info, hashv, pidx, cs, ce, newn = prepared(params, parts)
    sorted_parts = type(parts)(spos, svel, info, hashv)
    b, isb, cg, nbody, first_id = body_setup(params, sorted_parts)
    nl, _ = ob.build_neibs(params, spos, info, hashv, cs, ce)
    rbf = np.zeros((nbody, 4), dtype=np.float32)
    rbt = np.zeros((nbody, 4), dtype=np.float32)
    f, cfl, _ = ob.forces(params, spos, svel, info, hashv, cs, nl, bodies=b, rb_forces=rbf, rb_torques=rbt)
    f0, _, _ = ob.forces(params, spos, svel, info, hashv, cs, nl)
    ids = (info[:, 3].astype(np.int64) << 16) | info[:, 2]
    gp = global_positions(params, spos, hashv)
    k = ids[isb] - first_id
    m = spos[isb, 3:4]
    # force = acceleration x mass, also in the FORCES buffer; non-body particles untouched
    assert np.allclose(rbf[k, :3], f0[isb, :3] * m, rtol=1e-6, atol=1e-12)
    assert np.array_equal(f[~isb], f0[~isb]) and np.allclose(f[isb, :3], f0[isb, :3] * m, rtol=1e-6, atol=1e-12)
    arm = gp[isb] - cg
    tq = np.cross(arm, rbf[k, :3].astype(np.float64))
    scale = np.abs(tq).max() + 1e-12
    assert np.abs(rbt[k, :3] - tq).max() < 1e-4 * scale
    # rigid motion: rotate by 0.01 rad about z through the centre of gravity, translate, spin
    th = 0.01
    R = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    tr = np.array([1e-3, -2e-3, 5e-4])
    lv, om = np.array([0.1, 0.0, -0.05]), np.array([0.0, 0.2, 1.0])
    for i in range(9):
        b.steprot[1][i] = float(R.ravel()[i])
    for a in range(3):
        b.trans[1][a], b.linearvel[1][a], b.angularvel[1][a] = float(tr[a]), float(lv[a]), float(om[a])
    npos, nvel = ob.euler(params, spos, svel, info, hashv, f, 1e-4, 2, bodies=b)
    g1 = global_positions(params, npos, hashv)
    expect = cg + (R @ (gp[isb] - cg).T).T + tr
    assert np.abs(g1[isb] - expect).max() < 2e-7
    assert np.allclose(nvel[isb, :3], lv + np.cross(om, gp[isb] - cg), rtol=0, atol=2e-6)
    # walls (not moving) stay put
    wall = ((info[:, 0] & 7) == 1) & ~isb
    assert np.array_equal(npos[wall], spos[wall])
