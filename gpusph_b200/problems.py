"""Host-side problem setup: parameters and initial particle arrays.

Mirrors what GPUSPH's ProblemCore does on the host before the first step
(src/ProblemCore.cc:121-173 initialize, :1433-1496 set_grid_params, :1554-1583
calc_localpos_and_hash; defaults from src/simparams.h:280-300, src/physparams.h:385-400).
Pure numpy — nothing here touches the GPU.
"""
from __future__ import annotations

import dataclasses
import math

import os

import numpy as np

from . import capi

COORD_YZX = (1, 2, 0)   # reference default linearisation "yzx" (Makefile:517-523, src/linearization.h)


@dataclasses.dataclass
class ParticleArrays:
    """Host copies of the per-particle buffers in the reference's layouts."""
    pos: np.ndarray      # float32 [N,4] cell-local xyz + mass
    vel: np.ndarray      # float32 [N,4] velocity + relative density
    info: np.ndarray     # uint16  [N,4]
    hash: np.ndarray     # uint32  [N]

    @property
    def n(self) -> int:
        return int(self.pos.shape[0])


def make_info(ptype: int, flags: int = 0, fluid: int = 0, obj: int = 0, ids=None, n: int | None = None) -> np.ndarray:
    """particleinfo bit layout, src/particleinfo.h:295-330."""
    ids = np.arange(n, dtype=np.uint32) if ids is None else np.asarray(ids, dtype=np.uint32)
    info = np.zeros((ids.shape[0], 4), dtype=np.uint16)
    info[:, 0] = ptype | flags
    info[:, 1] = (fluid << 12) | obj
    info[:, 2] = ids & 0xFFFF
    info[:, 3] = ids >> 16
    return info


def make_params(*, origin, size, deltap, allocated_particles: int,
                sfactor: float = 1.3, kernelradius: float = 2.0, nlexpansionfactor: float = 1.0,
                neiblistsize: int = 128, periodic: int = 0, coord=COORD_YZX,
                rho0: float = 1000.0, gamma: float = 7.0, c0: float = 20.0,
                gravity=(0.0, 0.0, -9.81), densitydiffusion: int = capi.RHODIFF_NONE,
                density_diff_coeff: float | None = None,
                rheology: int = capi.RHEOLOGY_INVISCID, turbmodel: int = capi.TURB_ARTIFICIAL,
                kinvisc: float = 0.0, viscavgop: int = capi.AVG_ARITHMETIC,
                artvisccoeff: float = 0.3, dtadaptfactor: float = 0.3, fluids=None,
                viscmodel: int = capi.VISCMODEL_MORRIS, compvisc: int = capi.COMPVISC_KINEMATIC, bulkvisc: float = 0.0,
                simflags: int = capi.ENABLE_DTADAPT, epsxsph: float = 0.5, maxfall: float = 1.0) -> capi.Params:
    """Build b200sph_params the way ProblemCore + the engines' setconstants derive them."""
    p = capi.Params()
    p.abi_version = capi.ABI_VERSION
    slength = np.float32(sfactor * deltap)                      # set_deltap / set_smoothing, simparams.h:340-375
    influence = np.float32(slength * np.float32(kernelradius))
    nl_influence = float(nlexpansionfactor) * float(influence)
    cell_side = nl_influence                                     # ProblemCore.cc:1471
    for a in range(3):
        g = int(math.floor(size[a] / cell_side))                 # :1475-1477
        if g <= 0:
            raise ValueError("resolution too low: grid size would be 0 (ProblemCore.cc:1484-1489)")
        p.grid_size[a] = g
        p.cell_size[a] = np.float32(size[a] / g)                 # :1491-1493
        p.world_origin[a] = np.float32(origin[a])
        p.coord[a] = coord[a]
        p.gravity[a] = gravity[a]
    p.periodic = periodic
    p.neiblistsize = neiblistsize
    p.neibboundpos = neiblistsize - 1                            # non-SA: ProblemCore.h:346-353
    # (diagnostics: B200SPH_TEST_ALLOC_SCALE spreads the rows of the neighbour list further apart at a fixed particle count)
    allocated_particles = int(allocated_particles * float(os.environ.get("B200SPH_TEST_ALLOC_SCALE", "1")))
    p.neiblist_stride = allocated_particles
    p.nl_sq_influence_radius = np.float32(nl_influence * nl_influence)
    p.kerneltype = capi.KERNEL_WENDLAND
    p.sph_formulation = capi.SPH_F1
    p.densitydiffusiontype = densitydiffusion
    p.boundarytype = capi.DYN_BOUNDARY
    p.rheologytype = rheology
    p.turbmodel = turbmodel
    p.compvisc = compvisc
    p.viscmodel = viscmodel
    p.viscavgop = viscavgop
    p.is_const_visc = 1 if rheology == capi.RHEOLOGY_NEWTONIAN else 0   # single fluid, Newtonian, no k-eps (visc_spec.h:262)
    p.slength = slength
    p.influenceradius = influence
    p.deltap = np.float32(deltap)
    if density_diff_coeff is None:
        density_diff_coeff = 0.1 if densitydiffusion == capi.RHODIFF_COLAGROSSI else 0.0
    if densitydiffusion == capi.RHODIFF_COLAGROSSI:             # pre-multiplied by 2h, ProblemCore.cc:1411-1418
        density_diff_coeff = np.float32(np.float32(density_diff_coeff) * np.float32(2.0) * slength)
    p.density_diff_coeff = np.float32(density_diff_coeff)
    p.dtadaptfactor = dtadaptfactor
    p.num_fluids = 1
    p.rho0[0] = rho0
    p.bcoeff[0] = np.float32(rho0 * c0 * c0 / gamma)             # physparams.h:506-517
    p.gammacoeff[0] = gamma
    p.sscoeff[0] = c0
    p.sspowercoeff[0] = np.float32((gamma - 1.0) / 2.0)
    p.visccoeff[0] = kinvisc
    # additional fluids (multi-fluid runs): list of dicts with rho0, gamma, c0, kinvisc (physparams.h add_fluid)
    for f, fl in enumerate(fluids or [], start=1):
        p.num_fluids = f + 1
        p.rho0[f] = fl["rho0"]
        p.bcoeff[f] = np.float32(fl["rho0"] * fl["c0"] ** 2 / fl["gamma"])
        p.gammacoeff[f] = fl["gamma"]
        p.sscoeff[f] = fl["c0"]
        p.sspowercoeff[f] = np.float32((fl["gamma"] - 1.0) / 2.0)
        p.visccoeff[f] = fl.get("kinvisc", 0.0)
        c0 = max(c0, fl["c0"])
        kinvisc = max(kinvisc, fl.get("kinvisc", 0.0))
    if fluids:
        p.is_const_visc = 0                                     # IS_SINGLEFLUID is false (visc_spec.h:262)
    p.artvisccoeff = artvisccoeff
    p.epsartvisc = np.float32(0.01 * float(slength) * float(slength))   # ProblemCore.cc:160-161
    # float *= double literal: the product is evaluated in double and rounded once (GPUWorker.cc:3010-3011)
    p.max_sound_speed_cfl = np.float32(float(np.float32(c0)) * 1.1)
    # max kinematic viscosity for the viscous dt limit, scaled like GPUWorker::dt_reduce does (GPUWorker.cc:2011-2023)
    visc_for_dt = np.float32(kinvisc)
    if viscmodel == capi.VISCMODEL_MONAGHAN:
        visc_for_dt = np.float32(visc_for_dt * np.float32(10.0))
    elif viscmodel == capi.VISCMODEL_ESPANOL_REVENGA:
        visc_for_dt = np.float32(visc_for_dt * np.float32(5.0))
    p.max_kinvisc = visc_for_dt if rheology != capi.RHEOLOGY_INVISCID else 0.0
    p.dtadapt = 1 if simflags & capi.ENABLE_DTADAPT else 0
    if fluids:
        simflags |= capi.ENABLE_MULTIFLUID                      # required with more than one fluid (ProblemCore.cc:108-112)
    p.simflags = simflags
    p.epsxsph = epsxsph                                         # physparams.h:409
    p.monaghan_visc_coeff = 10.0                                # 2(d+2), physparams.h:396
    for f in range(p.num_fluids):
        p.visc2coeff[f] = (fluids[f - 1].get("bulkvisc", bulkvisc) if f else bulkvisc)
    # Lennard-Jones defaults of ProblemCore::initialize (src/ProblemCore.cc:121-140, physparams.h:398-403)
    p.r0 = np.float32(deltap)
    # dcoeff = 5 g H: H = 1 in ProblemCore, the problem's maximum fall height with the problem API (DamBreak3D: setMaxFall(0.4);
    # src/problem_api/ProblemAPI_1.cc:322-326)
    p.dcoeff = np.float32(np.float32(5.0) * np.float32(math.sqrt(sum(float(g) ** 2 for g in gravity))) * np.float32(maxfall))
    p.p1coeff, p.p2coeff = 12.0, 6.0
    p.partsurf = 0.0
    return p


def universe_box_planes(p: capi.Params, origin, size):
    """The six planes of ProblemAPI<1>::makeUniverseBox (src/problem_api/ProblemAPI_1.cc:1329-1365) as plane_t triples
    (unit normal, cell, in-cell position) the way ProblemCore::implicit_plane builds them (src/ProblemCore.cc:945-963): the
    reference point of a plane a x + b y + c z + d = 0 is the point of the plane closest to the centre of the domain."""
    origin = np.asarray(origin, dtype=np.float64)
    size = np.asarray(size, dtype=np.float64)
    lo, hi = origin, origin + size
    mid = origin + size / 2
    G = np.array([p.grid_size[a] for a in range(3)], dtype=np.int64)
    cs = size / G
    out = []
    for a in range(3):
        for sign, d in ((1.0, -lo[a]), (-1.0, hi[a])):
            n = np.zeros(3)
            n[a] = sign
            point = mid - (np.dot(mid, n) + d) * n
            g = np.minimum(np.maximum(np.floor((point - origin) / cs).astype(np.int64), 0), G - 1)
            local = (point - origin - (g + 0.5) * cs).astype(np.float32)
            out.append((tuple(n.astype(np.float32)), tuple(int(x) for x in g), tuple(float(x) for x in local)))
    return out


def initial_dt(p: capi.Params) -> float:
    """First-iteration dt: ProblemCore::check_dt, src/ProblemCore.cc:748-803 (float arithmetic)."""
    f32 = np.float32
    dt_ss = min(f32(p.slength) / f32(p.sscoeff[f]) for f in range(p.num_fluids)) * f32(p.dtadaptfactor)
    g = math.sqrt(sum(float(p.gravity[a]) ** 2 for a in range(3)))
    dt_g = f32(math.sqrt(float(p.slength) / g)) * f32(p.dtadaptfactor) if g > 0 else f32(np.inf)
    dt = min(f32(dt_ss), f32(dt_g))
    if p.rheologytype != capi.RHEOLOGY_INVISCID and p.max_kinvisc > 0:
        dt = min(dt, f32(f32(p.slength) * f32(p.slength) / f32(p.max_kinvisc)) * f32(0.125))
    return float(f32(dt))


def localpos_and_hash(p: capi.Params, gpos: np.ndarray, mass: np.ndarray, origin=None, size=None):
    """Global double positions -> (cell-local float4 + mass, cell hash); ProblemCore.cc:1554-1583. The reference does
    this in double with its double world origin and cell size (m_origin, m_cellsize = m_size / gridsize): pass
    origin / size to reproduce its choice of cell for particles lying on a cell face; without them the float copies
    the engines hold are used."""
    G = np.array([p.grid_size[a] for a in range(3)], dtype=np.int64)
    origin = np.array([p.world_origin[a] for a in range(3)], dtype=np.float64) if origin is None else np.asarray(origin, dtype=np.float64)
    cs = np.array([p.cell_size[a] for a in range(3)], dtype=np.float64) if size is None else np.asarray(size, dtype=np.float64) / G
    g = np.floor((gpos - origin) / cs).astype(np.int64)
    g = np.minimum(np.maximum(g, 0), G - 1)
    c = [p.coord[a] for a in range(3)]
    hashv = (g[:, c[2]] * G[c[1]] * G[c[0]] + g[:, c[1]] * G[c[0]] + g[:, c[0]]).astype(np.uint32)
    pos = np.empty((gpos.shape[0], 4), dtype=np.float32)
    pos[:, :3] = (gpos - origin - (g + 0.5) * cs).astype(np.float32)
    pos[:, 3] = mass
    return pos, hashv


def global_positions(p: capi.Params, pos: np.ndarray, hashv: np.ndarray) -> np.ndarray:
    """Inverse of localpos_and_hash (double): worldOrigin + (gridPos + 0.5)*cellSize + pos."""
    G = [int(p.grid_size[a]) for a in range(3)]
    c = [p.coord[a] for a in range(3)]
    cell = (hashv & 0x3FFFFFFF).astype(np.int64)
    g = np.empty((pos.shape[0], 3), dtype=np.int64)
    t = G[c[1]] * G[c[0]]
    g[:, c[2]] = cell // t
    r = cell - g[:, c[2]] * t
    g[:, c[1]] = r // G[c[0]]
    g[:, c[0]] = r - g[:, c[1]] * G[c[0]]
    origin = np.array([p.world_origin[a] for a in range(3)], dtype=np.float64)
    cs = np.array([p.cell_size[a] for a in range(3)], dtype=np.float64)
    return origin + (g + 0.5) * cs + pos[:, :3].astype(np.float64)


def lattice_problem(n: int, *, dp: float = 0.01, jitter: float = 0.05, seed: int = 12345,
                    densitydiffusion: int = capi.RHODIFF_NONE, rho0: float = 1000.0, c0: float = 20.0,
                    nz: int | None = None, ny: int | None = None, alloc_extra: float = 0.0, **kw):
    """Synthetic cubic fluid lattice of SURVEY.md section 8(d): n^3 (or n*ny*nz) fluid particles with
    spacing dp at (i+1/2)dp, domain padded by one cell, h = 1.3 dp, velocity
    0.1 c0 (sin, cos, sin)(2 pi x/L), optional uniform jitter of +-jitter*dp."""
    nx = n
    ny = n if ny is None else ny
    nz = n if nz is None else nz
    N = nx * ny * nz
    pad = 2.6 * dp
    L = np.array([nx, ny, nz], dtype=np.float64) * dp
    origin = -np.array([pad, pad, pad])
    size = L + 2 * pad
    params = make_params(origin=origin, size=size, deltap=dp, allocated_particles=int(N * (1 + alloc_extra)),
                         rho0=rho0, c0=c0, densitydiffusion=densitydiffusion, **kw)
    ii, jj, kk = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    g = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float64)
    gpos = (g + 0.5) * dp
    if jitter:
        rng = np.random.default_rng(seed)
        gpos = gpos + rng.uniform(-jitter * dp, jitter * dp, size=gpos.shape)
    mass = np.full(N, rho0 * dp ** 3, dtype=np.float32)
    pos, hashv = localpos_and_hash(params, gpos, mass)
    vel = np.zeros((N, 4), dtype=np.float32)
    ph = 2 * np.pi * gpos / L
    vel[:, 0] = 0.1 * c0 * np.sin(ph[:, 0])
    vel[:, 1] = 0.1 * c0 * np.cos(ph[:, 1])
    vel[:, 2] = 0.1 * c0 * np.sin(ph[:, 2])
    info = make_info(capi.PT_FLUID, n=N)
    return params, ParticleArrays(pos, vel, info, hashv)


def _box_shell(lo, hi, dp, layers):
    """Points of `layers` layers lining the inside of the box [lo,hi] (all six faces), laid out like the reference's
    Cube::FillIn (src/geometries/Cube.cc:564-640): n = (int)(length / dp) intervals per side (TRUNCATED, so the actual
    spacing is length / n >= dp), points at i / n of the side, i = 0..n."""
    n = np.maximum(np.floor((np.asarray(hi, dtype=np.float64) - np.asarray(lo, dtype=np.float64)) / dp).astype(int), 1)
    xs = [np.asarray(lo)[a] + (np.arange(n[a] + 1)) * ((np.asarray(hi)[a] - np.asarray(lo)[a]) / n[a]) for a in range(3)]
    I, J, K = np.meshgrid(np.arange(n[0] + 1), np.arange(n[1] + 1), np.arange(n[2] + 1), indexing="ij")
    shell = (I < layers) | (I > n[0] - layers) | (J < layers) | (J > n[1] - layers) | (K < layers) | (K > n[2] - layers)
    return np.stack([xs[0][I[shell]], xs[1][J[shell]], xs[2][K[shell]]], axis=1)


def dambreak_problem(dp: float = 0.015, *, densitydiffusion: int = capi.RHODIFF_COLAGROSSI, layers: int = 3,
                     alloc_extra: float = 0.0, width_scale: int = 1, obstacle: bool = False, testpoints: int = 0, **kw):
    """DamBreak3D-like setup (src/problems/DamBreak3D.cu:36-205 with --num_obstacles 0): a 1.6 x 0.67 x 0.6 m
    tank lined with `layers` layers of DYN boundary particles, a 0.4 m long, 0.4 m high water column,
    Wendland kernel, artificial viscosity, c0 = 20, gamma = 7. The particle SET is the reference's (same counts and
    positions as its Cube::FillIn / Cube::Fill, checked against the reference's own initial states in
    tests/test_golden.py); fill order/ids are ours (bit-level parity with the reference is tested from the
    reference's own initial state instead). Masses are the reference's too: volume-with-margin x density / particle
    count per geometry (Object::SetPartMass, src/geometries/Object.cc:45-52; Cube::Volume, Cube.cc:216-224), and every
    particle starts with the hydrostatic density of its depth below the water level (ProblemAPI_1.cc:1770-1810).
    width_scale = N widens the tank (and the water column) N times along y: the weak-scaling variant used by
    bench.py on N GPUs (same physics per unit width, N times the particles)."""
    dim = np.array([1.6, 0.67 * width_scale, 0.6])
    H = 0.4
    bd = dp * layers
    # non-fluid geometries are filled with r0 = the FLOAT copy of deltap (ProblemAPI<1>::preferredDeltaP,
    # src/problem_api/ProblemAPI_1.cc:109-118; PhysParams::r0 is a float), fluid ones with the double deltap
    dpb = float(np.float32(dp))
    wall = _box_shell(np.zeros(3), dim, dpb, layers)
    lo = np.array([bd, bd, bd])
    ext = np.array([0.4 - bd, dim[1] - 2 * bd, H - bd])
    nf = np.maximum(np.floor(ext / dp).astype(int), 1)             # Cube::Fill: (int)(l / dx), src/geometries/Cube.cc:405-440
    I, J, K = np.meshgrid(np.arange(nf[0] + 1), np.arange(nf[1] + 1), np.arange(nf[2] + 1), indexing="ij")
    fluid = lo + np.stack([I.ravel(), J.ravel(), K.ravel()], axis=1) * (ext / nf)
    nfl, nb = fluid.shape[0], wall.shape[0]
    body = np.zeros((0, 3))
    if obstacle:
        # the force-feedback obstacle of DamBreak3D (src/problems/DamBreak3D.cu:169-180): a 0.12 x 0.12 x 0.6 column
        # lined with DYN boundary particles, flagged FG_MOVING_BOUNDARY | FG_COMPUTE_FORCE, object number 1
        # (axis-aligned here; the reference rotates it by 45 degrees)
        side = 0.12
        blo = np.array([0.9 - side / 2, dim[1] / 2 - side / 2, bd])
        body = _box_shell(blo, blo + np.array([side, side, dim[2] - 2 * bd]), dpb, min(layers, 2))
    nob = body.shape[0]
    # test points (src/problems/DamBreak3D.cu:201-213): `testpoints` per probe column at 0.25, 0.4, 0.75, 0.9 of the length
    tp = np.zeros((0, 3))
    if testpoints > 0:
        dist = dim[2] / (testpoints + 1)
        tp = np.array([[fx * dim[0], dim[1] / 2.0, (t + 1) * dist / 2.0] for fx in (0.25, 0.4, 0.75, 0.9) for t in range(testpoints)])
    ntp = tp.shape[0]
    N = nfl + nb + nob + ntp
    rho0 = 1000.0
    c0 = 20.0
    params = make_params(origin=np.zeros(3), size=dim, deltap=dp, allocated_particles=int(N * (1 + alloc_extra)),
                         rho0=rho0, c0=c0, densitydiffusion=densitydiffusion, **kw)
    gpos = np.concatenate([fluid, wall, body, tp], axis=0)

    def part_mass(length, dx):
        # Object::SetPartMass(dx, rho): Volume(dx) * rho / (particles of a full Fill), Cube::Volume = prod(l + dx)
        length = np.asarray(length, dtype=np.float64)
        cnt = np.prod(np.maximum(np.floor(length / dx).astype(int), 1) + 1)
        return np.prod(length + dx) * rho0 / cnt
    mass = np.concatenate([np.full(nfl, part_mass(ext, dp)), np.full(nb, part_mass(dim, dpb)),
                           np.full(nob, part_mass([0.12, 0.12, dim[2]], dpb)), np.zeros(ntp)]).astype(np.float32)
    pos, hashv = localpos_and_hash(params, gpos, mass, origin=np.zeros(3), size=dim)
    vel = np.zeros((N, 4), dtype=np.float32)
    # hydrostatic initial density of EVERY particle below the water level (fluid, DYN boundary and test points alike,
    # ProblemAPI_1.cc:1770-1810; ProblemCore::hydrostatic_density, ProblemCore.cc:890-902, float arithmetic):
    # rho~ = (g rho0 h / B + 1)^(1/gamma) - 1
    f32 = np.float32
    gam = f32(7.0)
    B = f32(rho0 * c0 * c0 / 7.0)
    depth = (H - gpos[:, 2]).astype(np.float32)
    g_abs = f32(abs(kw.get("gravity", (0.0, 0.0, -9.81))[2]))
    dens = np.where(depth > 0, np.power((g_abs * f32(rho0) * depth / B + f32(1.0)).astype(np.float64), 1.0 / float(gam)) - 1.0, 0.0)
    vel[:, 3] = dens.astype(np.float32)
    info = np.concatenate([make_info(capi.PT_FLUID, ids=np.arange(nfl)),
                           make_info(capi.PT_BOUNDARY, ids=np.arange(nfl, nfl + nb)),
                           make_info(capi.PT_BOUNDARY, flags=capi.FG_MOVING_BOUNDARY | capi.FG_COMPUTE_FORCE, obj=1,
                                     ids=np.arange(nfl + nb, nfl + nb + nob)),
                           make_info(capi.PT_TESTPOINT, ids=np.arange(nfl + nb + nob, N))], axis=0)
    return params, ParticleArrays(pos, vel, info, hashv)


def poiseuille_problem(ppH: int = 32, *, lz: float = 1.0, kinvisc: float = 0.1, driving_force: float = 0.05,
                       rho0: float = 1.0, viscavgop: int = capi.AVG_ARITHMETIC, steady_init: bool = True,
                       layers: int = 3, alloc_extra: float = 0.0, ncell: int = 4, **kw):
    """Plane Poiseuille flow like src/problems/Poiseuille.inc:60-232: fluid between two DYN-boundary plates at
    z = +-lz/2, periodic in x and y, Newtonian laminar (Morris) viscosity with constant kinematic viscosity,
    driven by a body force along x. Analytic steady profile (Poiseuille.inc:187-232, scripts/validate-poiseuille.py:32-37):
        v_x(z) = F/(2 nu) ((lz/2)^2 - z^2).
    The flow is invariant in x and y: a thin periodic box of ncell x ncell cells (default 4) is enough for validation;
    the reference's lx = ly = lz cube (Poiseuille.inc:120-130; ~1 M particles at --ppH 100) is ncell ~ lz / (2.6 dp)."""
    dp = lz / ppH
    max_vel = driving_force / (2 * kinvisc) * (lz / 2) ** 2
    hydro = math.sqrt(2 * driving_force * lz)
    c0 = 20 * max(hydro, max_vel)                                  # Poiseuille.inc:147
    h = 1.3 * dp
    cell = 2 * h
    lx = ly = round(ncell * cell / dp) * dp                         # multiple of dp (periodicity, ProblemCore.cc:1436-1456)
    zpad = (layers - 0.5) * dp
    origin = np.array([-lx / 2, -ly / 2, -lz / 2 - zpad])
    size = np.array([lx, ly, lz + 2 * zpad])
    nx, ny = int(round(lx / dp)), int(round(ly / dp))
    xs = origin[0] + (np.arange(nx) + 0.5) * dp
    ys = origin[1] + (np.arange(ny) + 0.5) * dp
    # as in the reference: the first wall layer sits AT z = +-lz/2 (addRect at +-lz/2, further layers outwards) and the
    # fluid box is lz - 2 dp high (Poiseuille.inc:152-160)
    zf = -lz / 2 + (np.arange(ppH - 1) + 1.0) * dp
    zb = np.concatenate([-lz / 2 - np.arange(layers) * dp, lz / 2 + np.arange(layers) * dp])
    X, Y, Z = np.meshgrid(xs, ys, zf, indexing="ij")
    fluid = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    X, Y, Z = np.meshgrid(xs, ys, zb, indexing="ij")
    wall = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    nfl, nb = fluid.shape[0], wall.shape[0]
    N = nfl + nb
    params = make_params(origin=origin, size=size, deltap=dp, allocated_particles=int(N * (1 + alloc_extra)),
                         rho0=rho0, c0=c0, gravity=(driving_force, 0.0, 0.0), periodic=capi.PERIODIC_X | capi.PERIODIC_Y,
                         rheology=capi.RHEOLOGY_NEWTONIAN, turbmodel=capi.TURB_LAMINAR, kinvisc=kinvisc, viscavgop=viscavgop, **kw)
    gpos = np.concatenate([fluid, wall], axis=0)
    mass = np.full(N, rho0 * dp ** 3, dtype=np.float32)
    pos, hashv = localpos_and_hash(params, gpos, mass)
    vel = np.zeros((N, 4), dtype=np.float32)
    if steady_init:
        vel[:nfl, 0] = driving_force / (2 * kinvisc) * ((lz / 2) ** 2 - fluid[:, 2] ** 2)
    info = np.concatenate([make_info(capi.PT_FLUID, ids=np.arange(nfl)),
                           make_info(capi.PT_BOUNDARY, ids=np.arange(nfl, N))], axis=0)
    return params, ParticleArrays(pos, vel, info, hashv)
