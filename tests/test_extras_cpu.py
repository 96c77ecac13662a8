"""Pin the oracle's restatement of the "next" rows of SURVEY.md section 8f — density filters (Shepard, MLS), TESTPOINTS,
BREZZI diffusion, the MONAGHAN / ESPANOL_REVENGA viscous models, XSPH, geometric planes — against INDEPENDENT float64
numpy evaluations of the published formulas with brute-force neighbours (the reference ships no golden vectors for
these either)."""
import numpy as np
import pytest

import oracle_binding as ob
from gpusph_b200 import capi
from gpusph_b200.problems import dambreak_problem, global_positions, lattice_problem, poiseuille_problem
from test_oracle_cpu import prepared


def perturbed(params, parts, seed=3, dv=0.4, drho=2e-3):
    rng = np.random.default_rng(seed)
    fl = (parts.info[:, 0] & 7) == 0
    parts.vel[:, :3] += rng.normal(0, dv, size=(parts.n, 3)).astype(np.float32) * fl[:, None]
    parts.vel[:, 3] += rng.normal(0, drho, size=parts.n).astype(np.float32)
    spos, svel, info, hashv, pidx, cs, ce, newn = prepared(params, parts)
    nl, _ = ob.build_neibs(params, spos, info, hashv, cs, ce)
    return spos, svel, info, hashv, cs, nl


class Field:
    """float64 view of a particle state with brute-force neighbour queries."""

    def __init__(self, params, pos, vel, info, hashv):
        self.p = params
        self.g = global_positions(params, pos, hashv)
        self.v = vel[:, :3].astype(np.float64)
        self.m = pos[:, 3].astype(np.float64)
        self.rt = vel[:, 3].astype(np.float64)
        self.rho0 = float(params.rho0[0])
        self.rho = (self.rt + 1) * self.rho0
        self.P = float(params.bcoeff[0]) * ((self.rt + 1) ** float(params.gammacoeff[0]) - 1)
        self.ptype = info[:, 0] & 7
        self.h = float(params.slength)
        self.R = float(params.influenceradius)
        self.n = pos.shape[0]
        # periodic minimum-image extents
        self.L = np.array([float(params.cell_size[a]) * int(params.grid_size[a]) if params.periodic & (1 << a) else 0.0 for a in range(3)])

    def W(self, r):
        q = r / self.h
        return 21.0 / (16.0 * np.pi * self.h ** 3) * (1 - q / 2) ** 4 * (1 + 2 * q)

    def F(self, r):
        return 105.0 / (128.0 * np.pi * self.h ** 5) * (r / self.h - 2) ** 3

    def neibs(self, i, types=(0, 1)):
        d = self.g[i] - self.g
        for a in range(3):
            if self.L[a]:
                d[:, a] -= self.L[a] * np.round(d[:, a] / self.L[a])
        r = np.sqrt((d * d).sum(axis=1))
        mask = (r < self.R) & np.isin(self.ptype, types)
        mask[i] = False
        j = np.flatnonzero(mask)
        return j, d[j], r[j]


# ---------------------------------------------------------------------------------------------- filters
def test_shepard_filter_against_numpy():
    """rho_i = sum_j m_j W_ij / sum_j (m_j / rho_j) W_ij over fluid + DYN boundary neighbours and the particle itself."""
    params, parts = dambreak_problem(0.05)
    spos, svel, info, hashv, cs, nl = perturbed(params, parts)
    out = ob.shepard(params, spos, svel, info, hashv, cs, nl)
    fld = Field(params, spos, svel, info, hashv)
    for i in np.flatnonzero(fld.ptype == 0)[::7]:
        j, d, r = fld.neibs(i)
        w = np.concatenate([[fld.W(0.0) * fld.m[i]], fld.W(r) * fld.m[j]])
        rho = w.sum() / (w / np.concatenate([[fld.rho[i]], fld.rho[j]])).sum()
        assert out[i, 3] == pytest.approx(rho / fld.rho0 - 1, abs=3e-6)
    # velocities untouched; non-fluid particles copied through
    assert np.array_equal(out[:, :3], svel[:, :3])
    assert np.array_equal(out[fld.ptype != 0], svel[fld.ptype != 0])


def test_mls_filter_against_numpy_linear_solve():
    """MLS (Dilts 1999 / Colagrossi & Landrini 2003): rho_i = sum_j (b0 + b . r_ij) m_j W_ij with b = first row of the
    inverse of A = sum_j V_j W_ij [1, r_ij]^T [1, r_ij]. The oracle follows the reference's float determinant + conjugate
    residual solve; numpy solves the system in float64."""
    params, parts = dambreak_problem(0.05, densitydiffusion=capi.RHODIFF_NONE)
    spos, svel, info, hashv, cs, nl = perturbed(params, parts)
    out = ob.mls(params, spos, svel, info, hashv, cs, nl)
    fld = Field(params, spos, svel, info, hashv)
    checked = 0
    for i in np.flatnonzero(fld.ptype <= 1)[::5]:
        # a DYN boundary particle lists fluid neighbours only (src/cuda/buildneibs_kernel.cu:591-602)
        j, d, r = fld.neibs(i, types=(0, 1) if fld.ptype[i] == 0 else (0,))
        if j.shape[0] < 25:
            continue                      # ill-conditioned moments near free surfaces/corners: not a formula check
        X = np.concatenate([np.zeros((1, 3)), d / fld.h])
        jj = np.concatenate([[i], j])
        rr = np.concatenate([[0.0], r])
        w = fld.W(rr) * fld.m[jj] / fld.rho[jj]
        basis = np.concatenate([np.ones((X.shape[0], 1)), X], axis=1)
        A = (basis[:, :, None] * basis[:, None, :] * w[:, None, None]).sum(axis=0)
        if np.linalg.cond(A) > 1e4:
            continue
        beta = np.linalg.solve(A, np.array([1.0, 0, 0, 0]))
        rho = ((basis @ beta) * fld.W(rr) * fld.m[jj]).sum()
        assert out[i, 3] == pytest.approx(rho / fld.rho0 - 1, abs=2e-4)
        checked += 1
    assert checked > 50
    assert np.array_equal(out[:, :3], svel[:, :3])


def test_mls_reproduces_a_linear_density_field_exactly():
    """The defining property of first-order MLS: a density field that is linear in space is a fixed point (to float
    accuracy) for particles with full kernel support."""
    params, parts = lattice_problem(10, jitter=0.2)
    spos, svel, info, hashv, pidx, cs, ce, newn = prepared(params, parts)
    g = global_positions(params, spos, hashv)
    svel[:, 3] = (1e-3 * (g[:, 0] + 2 * g[:, 1] - g[:, 2])).astype(np.float32)
    # constant volumes m/rho so that the moments are those of the particle positions
    spos[:, 3] = ((svel[:, 3] + 1) * float(params.rho0[0]) * 0.01 ** 3).astype(np.float32)
    nl, _ = ob.build_neibs(params, spos, info, hashv, cs, ce)
    out = ob.mls(params, spos, svel, info, hashv, cs, nl)
    lo, hi = g.min(axis=0) + 0.03, g.max(axis=0) - 0.03
    inner = ((g > lo) & (g < hi)).all(axis=1)
    assert inner.sum() > 20
    assert np.abs(out[inner, 3] - svel[inner, 3]).max() < 5e-6


def test_testpoints_against_numpy():
    """Shepard-normalised velocity and pressure of the fluid neighbours, in place, only on test points."""
    params, parts = dambreak_problem(0.05, testpoints=3)
    spos, svel, info, hashv, cs, nl = perturbed(params, parts)
    tke = np.abs(np.random.default_rng(1).normal(0, 1, size=parts.n)).astype(np.float32)
    v, k, e = ob.testpoints(params, spos, svel, info, hashv, cs, nl, tke=tke)
    fld = Field(params, spos, svel, info, hashv)
    tps = np.flatnonzero(fld.ptype == 3)
    assert tps.shape[0] == 12
    wet = 0
    for i in tps:
        j, d, r = fld.neibs(i, types=(0,))
        w = fld.W(r) * fld.m[j] / fld.rho[j]
        if w.sum() > 1e-5:
            wet += 1
            assert np.allclose(v[i, :3], (w[:, None] * fld.v[j]).sum(axis=0) / w.sum(), rtol=2e-5, atol=1e-6)
            assert v[i, 3] == pytest.approx((w * fld.P[j]).sum() / w.sum(), rel=5e-4, abs=0.5)
            assert k[i] == pytest.approx((w * tke[j]).sum() / w.sum(), rel=2e-5)
        else:
            assert (v[i] == 0).all() and k[i] == 0
    assert 0 < wet < 12
    others = fld.ptype != 3
    assert np.array_equal(v[others], svel[others]) and np.array_equal(k[others], tke[others]) and e is None


# ---------------------------------------------------------------------------------------------- forces options
def numpy_terms(fld, params, i, *, brezzi_dt=None, viscmodel=None, nu=0.0, zeta=0.0, avg=capi.AVG_ARITHMETIC):
    """float64 pair sums for fluid particle i: (extra continuity term / rho0, viscous acceleration, XSPH mean velocity)."""
    grav = np.array([params.gravity[a] for a in range(3)], dtype=np.float64)
    eps = float(params.epsartvisc)
    j, d, r = fld.neibs(i)
    F = fld.F(r)
    vij = fld.v[i] - fld.v[j]
    vr = (vij * d).sum(axis=1)
    fl = fld.ptype[j] == 0
    dr = 0.0
    if brezzi_dt is not None:
        t = float(params.density_diff_coeff) * ((2.0 / (fld.rho[i] + fld.rho[j])) * (fld.P[i] - fld.P[j]) - d @ grav) \
            * fld.m[j] / fld.rho[j] * F * brezzi_dt * 2.0 * fld.rho[i]
        dr = np.where(fl, t, 0.0).sum() / fld.rho0
    visc = np.zeros(3)
    if viscmodel is not None:
        ri, rj = fld.rho[i], fld.rho[j]

        def average(a, b):
            return {capi.AVG_ARITHMETIC: (a + b) / 2, capi.AVG_HARMONIC: 2 * a * b / (a + b), capi.AVG_GEOMETRIC: np.sqrt(a * b)}[avg]
        if viscmodel == capi.VISCMODEL_ESPANOL_REVENGA:
            third = average(nu * ri, nu * rj) / 3
            bulk = average(zeta, zeta + 0 * rj)
            c = fld.m[j] / (ri * rj) * F
            visc = (c[:, None] * ((5 * third - bulk)[:, None] * vij + (5 * (third + bulk) * vr / (r * r + eps))[:, None] * d)).sum(axis=0)
        else:
            # constant kinematic viscosity, single fluid: nu * <density average>  (visc_avg.cu)
            dens = {capi.AVG_ARITHMETIC: fld.m[j] * (ri + rj) / (ri * rj), capi.AVG_HARMONIC: 4 * fld.m[j] / (ri + rj),
                    capi.AVG_GEOMETRIC: 2 * fld.m[j] / np.sqrt(ri * rj)}[avg]
            s = nu * dens * F
            if viscmodel == capi.VISCMODEL_MONAGHAN:
                coef = np.where(vr < 0, 10.0 * vr / (r * r + eps), 0.0)
                visc = ((s * coef)[:, None] * d).sum(axis=0)
            else:
                visc = (s[:, None] * vij).sum(axis=0)
    xs = 2 * (np.where(fl, -fld.m[j] * fld.W(r) / (fld.rho[i] + fld.rho[j]), 0.0)[:, None] * vij).sum(axis=0)
    return dr, visc, xs


def test_brezzi_diffusion_term():
    """BREZZI = the NONE continuity equation + coeff (2 (P_i - P_j)/(rho_i + rho_j) - g.r_ij) (m_j/rho_j) F 2 dt rho_i
    over fluid neighbours (Ferrand et al. 2017)."""
    base, parts = dambreak_problem(0.06, densitydiffusion=capi.RHODIFF_NONE)
    spos, svel, info, hashv, cs, nl = perturbed(base, parts)
    f0, _, ab = ob.forces(base, spos, svel, info, hashv, cs, nl, want_abssum=True)
    params = base.copy()
    params.densitydiffusiontype = capi.RHODIFF_BREZZI
    params.density_diff_coeff = 0.2
    dt = 2.5e-4
    f1, _, ab1 = ob.forces(params, spos, svel, info, hashv, cs, nl, want_abssum=True, opts=ob.forces_opts(dt=dt))
    fld = Field(params, spos, svel, info, hashv)
    assert np.array_equal(f0[:, :3], f1[:, :3])              # momentum untouched
    big = 0
    for i in np.flatnonzero(fld.ptype <= 1)[::9]:
        dr, _, _ = numpy_terms(fld, params, i, brezzi_dt=dt)
        scale = ab1[i, 3] / fld.rho0 + 1e-9
        assert abs((f1[i, 3] - f0[i, 3]) - dr) <= 1e-4 * scale
        big += abs(dr) > 1e-3 * scale
    assert big > 20                                          # the term is not negligible in this test


@pytest.mark.parametrize("viscmodel,avg", [(capi.VISCMODEL_MORRIS, capi.AVG_HARMONIC), (capi.VISCMODEL_MONAGHAN, capi.AVG_ARITHMETIC),
                                           (capi.VISCMODEL_MONAGHAN, capi.AVG_GEOMETRIC),
                                           (capi.VISCMODEL_ESPANOL_REVENGA, capi.AVG_ARITHMETIC),
                                           (capi.VISCMODEL_ESPANOL_REVENGA, capi.AVG_HARMONIC)])
def test_laminar_viscous_models(viscmodel, avg):
    """Newtonian laminar viscosity = inviscid right-hand side + the model's pair term: MORRIS along v_ij, MONAGHAN along
    r_ij (approaching pairs only, coefficient 2(d+2) = 10), ESPANOL_REVENGA both with shear and bulk viscosity."""
    nu, zeta = 0.05, 12.0
    inv, parts = poiseuille_problem(10, kinvisc=nu, viscavgop=avg)
    inv.rheologytype, inv.turbmodel, inv.max_kinvisc = capi.RHEOLOGY_INVISCID, capi.TURB_LAMINAR, 0.0
    spos, svel, info, hashv, cs, nl = perturbed(inv, parts, dv=0.05, drho=1e-3)
    f0, _, _ = ob.forces(inv, spos, svel, info, hashv, cs, nl)
    params, _ = poiseuille_problem(10, kinvisc=nu, viscavgop=avg, viscmodel=viscmodel, bulkvisc=zeta)
    f1, _, ab = ob.forces(params, spos, svel, info, hashv, cs, nl, want_abssum=True)
    fld = Field(params, spos, svel, info, hashv)
    assert np.array_equal(f0[:, 3], f1[:, 3])                # continuity untouched
    big = 0
    for i in np.flatnonzero(fld.ptype == 0)[::5]:
        _, visc, _ = numpy_terms(fld, params, i, viscmodel=viscmodel, nu=nu, zeta=zeta, avg=avg)
        scale = ab[i, 0] + 1e-9
        assert np.abs((f1[i, :3] - f0[i, :3]) - visc).max() <= 2e-4 * scale
        big += np.abs(visc).max() > 1e-2 * scale
    assert big > 20


def test_xsph_mean_velocity_and_corrected_euler():
    params, parts = dambreak_problem(0.06, simflags=capi.ENABLE_DTADAPT | capi.ENABLE_XSPH, epsxsph=0.5)
    spos, svel, info, hashv, cs, nl = perturbed(params, parts)
    xs = np.zeros((parts.n, 4), dtype=np.float32)
    f, _, _ = ob.forces(params, spos, svel, info, hashv, cs, nl, opts=ob.forces_opts(xsph=xs))
    noflag = params.copy()
    noflag.simflags = capi.ENABLE_DTADAPT
    f_plain, _, _ = ob.forces(noflag, spos, svel, info, hashv, cs, nl)
    assert np.array_equal(f, f_plain)                        # XSPH does not touch the forces themselves
    fld = Field(params, spos, svel, info, hashv)
    for i in np.flatnonzero(fld.ptype == 0)[::7]:
        _, _, mv = numpy_terms(fld, params, i)
        assert np.allclose(xs[i, :3], mv, rtol=1e-4, atol=1e-5 * np.abs(fld.v).max())
    assert (xs[fld.ptype != 0] == 0).all() and (xs[:, 3] == 0).all()
    # euler: fluid positions move with v + eps * xsph (+ f dt/2 on the corrector); velocities as without XSPH
    dt = 2e-4
    for step, d in ((1, dt / 2), (2, dt)):
        p1, v1 = ob.euler(params, spos, svel, info, hashv, f, d, step, xsph=xs)
        p0, v0 = ob.euler(noflag, spos, svel, info, hashv, f, d, step)
        assert np.array_equal(v1, v0)
        fl = fld.ptype == 0
        assert np.allclose(p1[fl, :3] - p0[fl, :3], 0.5 * xs[fl, :3] * d, rtol=1e-3, atol=2e-8)   # differences of float positions ~0.1: 1 ulp = 7e-9
        assert np.array_equal(p1[~fl], p0[~fl])


def test_plane_repulsion_and_wall_friction():
    """Lennard-Jones repulsion D ((r0/r)^p1 - (r0/r)^p2) / r^2 along the plane normal for fluid particles closer than r0,
    plus the friction -mu A/(m r) v_t when the fluid is viscous (Monaghan 1994 boundaries as GPUSPH implements them)."""
    nu = 0.02
    params, parts = lattice_problem(8, jitter=0.2, simflags=capi.ENABLE_DTADAPT | capi.ENABLE_PLANES,
                                    rheology=capi.RHEOLOGY_NEWTONIAN, kinvisc=nu)
    spos, svel, info, hashv, cs, nl = perturbed(params, parts)
    g = global_positions(params, spos, hashv)
    z0 = float(g[:, 2].min()) - 0.3 * float(params.r0)
    x1 = float(g[:, 0].max()) + 0.5 * float(params.r0)
    # plane reference points given in grid + local coordinates like plane_t: cell (0,0,0) / cell (Gx-1, 0, 0)
    cs3 = np.array([params.cell_size[a] for a in range(3)], dtype=np.float64)
    org = np.array([params.world_origin[a] for a in range(3)], dtype=np.float64)

    def plane(normal, point):
        gp = np.floor((point - org) / cs3).astype(int)
        return (normal, gp, (point - org - (gp + 0.5) * cs3).astype(np.float32))
    planes = [plane((0.0, 0.0, 1.0), np.array([0.02, 0.03, z0])), plane((-1.0, 0.0, 0.0), np.array([x1, 0.01, 0.04]))]
    f1, _, _ = ob.forces(params, spos, svel, info, hashv, cs, nl, opts=ob.forces_opts(planes=planes))
    f0, _, _ = ob.forces(params, spos, svel, info, hashv, cs, nl)
    fld = Field(params, spos, svel, info, hashv)
    r0, D, p1, p2 = float(params.r0), float(params.dcoeff), 12.0, 6.0
    expect = np.zeros((parts.n, 3))
    for normal, point in (((0, 0, 1.0), (0.02, 0.03, z0)), ((-1.0, 0, 0), (x1, 0.01, 0.04))):
        nrm = np.array(normal)
        r = np.abs((g - np.array(point)) @ nrm)
        near = (r < r0) & (fld.ptype == 0)
        lj = D * ((r0 / r) ** p1 - (r0 / r) ** p2) / (r * r)
        vt = fld.v - (fld.v @ nrm)[:, None] * nrm
        fric = -(nu * fld.rho) * (r0 * r0) / (fld.m * r)
        expect += np.where(near[:, None], (lj * r)[:, None] * nrm + fric[:, None] * vt, 0.0)
    assert (np.abs(expect).sum(axis=1) > 0).sum() > 30
    assert np.allclose(f1[:, :3] - f0[:, :3], expect, rtol=2e-3, atol=1e-3 * np.abs(expect).max())
    assert np.array_equal(f1[:, 3], f0[:, 3])


def test_worker_with_mls_filter_and_brezzi_runs_and_stays_finite():
    params, parts = dambreak_problem(0.06, densitydiffusion=capi.RHODIFF_BREZZI, density_diff_coeff=0.1, testpoints=2)
    w = ob.OracleWorker(params, parts, filters={"MLS_FILTER": 3})
    for _ in range(7):
        w.step()
    w.postprocess()
    assert np.isfinite(w.pos).all() and np.isfinite(w.vel).all() and w.dt > 0
    assert np.abs(w.vel[(w.info[:, 0] & 7) == 0, 3]).max() < 0.05
