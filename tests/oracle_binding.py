"""ctypes binding of oracle/liboracle.so (the CPU restatement) for tests / smoke / cpu_baseline.

TEST INFRASTRUCTURE: the product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys_path_lib = os.path.join(ROOT, "oracle", "liboracle.so")

from gpusph_b200 import capi  # noqa: E402  (Params struct only: the shared C header)
from gpusph_b200.problems import ParticleArrays, initial_dt  # noqa: E402

_lib = None


MAXB = 16


class OracleBodies(C.Structure):
    """oracle_bodies (oracle/sph_oracle.c): mirror of the reference's per-body __constant__ arrays."""
    _fields_ = [("cgGridPos", (C.c_int * 3) * MAXB), ("cgPos", (C.c_float * 3) * MAXB), ("startIndex", C.c_int * MAXB),
                ("trans", (C.c_float * 3) * MAXB), ("steprot", (C.c_float * 9) * MAXB),
                ("linearvel", (C.c_float * 3) * MAXB), ("angularvel", (C.c_float * 3) * MAXB)]


class OracleForcesOpts(C.Structure):
    """oracle_forces_opts (oracle/sph_oracle.c): arguments only some option combinations read."""
    _fields_ = [("dt", C.c_float), ("xsph", C.c_void_p), ("numplanes", C.c_int),
                ("plane_normal", (C.c_float * 3) * 8), ("plane_gridpos", (C.c_int * 3) * 8), ("plane_pos", (C.c_float * 3) * 8)]


def forces_opts(dt=0.0, xsph=None, planes=None):
    """planes: list of (normal[3], gridPos[3], pos[3]) like plane_t (src/planes.h:42-46)."""
    o = OracleForcesOpts()
    o.dt = dt
    o.xsph = None if xsph is None else xsph.ctypes.data
    o.numplanes = 0 if planes is None else len(planes)
    for k, (nrm, gp, pp) in enumerate(planes or []):
        for a in range(3):
            o.plane_normal[k][a] = nrm[a]
            o.plane_gridpos[k][a] = int(gp[a])
            o.plane_pos[k][a] = pp[a]
    return o


class OracleNeibsInfo(C.Structure):
    _fields_ = [("num_interactions", C.c_int32), ("max_fluid_boundary_neibs", C.c_int32),
                ("max_vertex_neibs", C.c_int32), ("has_too_many_neibs", C.c_int32),
                ("has_max_neibs", C.c_int32 * 3)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(sys_path_lib):
            import __graft_entry__ as g
            g.build_oracle()
        _lib = C.CDLL(sys_path_lib)
        _lib.oracle_forces.restype = C.c_uint32
        _lib.oracle_forces_ex.restype = C.c_uint32
        _lib.oracle_dtreduce.restype = C.c_float
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def calc_hash(params, pos, hashv, info, cdm=None):
    n = pos.shape[0]
    pidx = np.empty(n, dtype=np.uint32)
    lib().oracle_calc_hash(C.byref(params), _p(pos), _p(hashv), _p(pidx), _p(info), _p(cdm), C.c_uint32(n))
    return pidx


def fix_hash(params, hashv, info, cdm=None):
    n = hashv.shape[0]
    pidx = np.empty(n, dtype=np.uint32)
    lib().oracle_fix_hash(C.byref(params), _p(hashv), _p(pidx), _p(info), _p(cdm), C.c_uint32(n))
    return pidx


def sort(hashv, info, pidx):
    lib().oracle_sort(_p(hashv), _p(info), _p(pidx), C.c_uint32(hashv.shape[0]))


def reorder(params, pos, vel, info, hashv, pidx, segments=False):
    n = pos.shape[0]
    nc = params.num_cells
    cs = np.full(nc, 0xFFFFFFFF, dtype=np.uint32)
    ce = np.full(nc, 0xFFFFFFFF, dtype=np.uint32)
    seg = np.zeros(4, dtype=np.uint32) if segments else None
    spos = np.zeros_like(pos)
    svel = np.zeros_like(vel)
    newn = np.zeros(1, dtype=np.uint32)
    lib().oracle_reorder(_p(cs), _p(ce), _p(seg), _p(spos), _p(svel), _p(pos), _p(vel), _p(info), _p(hashv), _p(pidx),
                         C.c_uint32(n), _p(newn))
    return cs, ce, seg, spos, svel, int(newn[0])


def build_neibs(params, pos, info, hashv, cs, ce, range_end=None):
    n = pos.shape[0]
    range_end = n if range_end is None else range_end
    nl = np.full((int(params.neiblistsize), int(params.neiblist_stride)), 0xFFFF, dtype=np.uint16)
    out = OracleNeibsInfo()
    lib().oracle_build_neibs(C.byref(params), _p(pos), _p(info), _p(hashv), _p(cs), _p(ce), _p(nl),
                             C.c_uint32(n), C.c_uint32(range_end), C.byref(out))
    return nl, out


def forces(params, pos, vel, info, hashv, cs, nl, eos_p=None, eos_c=None, from_=0, to=None, want_abssum=False,
           bodies=None, rb_forces=None, rb_torques=None, opts=None):
    n = pos.shape[0]
    to = n if to is None else to
    f = np.zeros((n, 4), dtype=np.float32)
    nb = (to - from_ + 127) // 128
    nb = (nb + 3) // 4 * 4
    cfl = np.zeros(max(nb, 4), dtype=np.float32)
    ab = np.zeros((n, 4), dtype=np.float32) if want_abssum else None
    got = lib().oracle_forces_ex(C.byref(params), _p(pos), _p(vel), _p(info), _p(hashv), _p(cs), _p(nl),
                                 _p(eos_p), _p(eos_c), _p(f), _p(cfl), _p(ab),
                                 C.c_uint32(n), C.c_uint32(from_), C.c_uint32(to), C.c_uint32(0),
                                 C.byref(bodies) if bodies is not None else None, _p(rb_forces), _p(rb_torques),
                                 C.byref(opts) if opts is not None else None)
    assert got == nb
    return f, cfl[:nb], ab


def eos(params, vel, info):
    n = vel.shape[0]
    p = np.zeros(n, dtype=np.float32)
    c = np.zeros(n, dtype=np.float32)
    lib().oracle_eos(C.byref(params), _p(vel), _p(info), _p(p), _p(c), C.c_uint32(n))
    return p, c


def dtreduce(params, cfl):
    return float(lib().oracle_dtreduce(C.byref(params), _p(cfl), C.c_uint32(cfl.shape[0])))


def euler(params, old_pos, old_vel, info, hashv, f, dt, step, range_end=None, bodies=None, xsph=None):
    n = old_pos.shape[0]
    range_end = n if range_end is None else range_end
    npos = old_pos.copy()
    nvel = old_vel.copy()
    lib().oracle_euler_ex(C.byref(params), _p(old_pos), _p(old_vel), _p(info), _p(hashv), _p(f), _p(xsph), _p(npos), _p(nvel),
                          C.c_uint32(n), C.c_uint32(range_end), C.c_float(dt), C.c_int(step),
                          C.byref(bodies) if bodies is not None else None)
    return npos, nvel


def shepard(params, pos, vel, info, hashv, cs, nl, range_end=None):
    n = pos.shape[0]
    out = vel.copy()
    lib().oracle_shepard(C.byref(params), _p(pos), _p(vel), _p(out), _p(info), _p(hashv), _p(cs), _p(nl),
                         C.c_uint32(n if range_end is None else range_end))
    return out


def mls(params, pos, vel, info, hashv, cs, nl, range_end=None):
    n = pos.shape[0]
    out = vel.copy()
    lib().oracle_mls(C.byref(params), _p(pos), _p(vel), _p(out), _p(info), _p(hashv), _p(cs), _p(nl),
                     C.c_uint32(n if range_end is None else range_end))
    return out


def testpoints(params, pos, vel, info, hashv, cs, nl, tke=None, epsilon=None, range_end=None):
    """In place on copies: returns (vel, tke, epsilon)."""
    n = pos.shape[0]
    v = vel.copy()
    k = None if tke is None else tke.copy()
    e = None if epsilon is None else epsilon.copy()
    lib().oracle_testpoints(C.byref(params), _p(pos), _p(v), _p(k), _p(e), _p(info), _p(hashv), _p(cs), _p(nl),
                            C.c_uint32(n if range_end is None else range_end))
    return v, k, e


class OracleWorker:
    """CPU twin of gpusph_b200.simulation.Worker built from the oracle functions (same call order)."""

    def __init__(self, params, particles: ParticleArrays, buildneibsfreq: int = 10, fixed_dt=None, start_iteration: int = 0, dt=None,
                 filters=None, planes=None):
        self.params = params
        self.pos = particles.pos.copy()
        self.vel = particles.vel.copy()
        self.info = particles.info.copy()
        self.hash = particles.hash.copy()
        self.n = particles.n
        self.buildneibsfreq = buildneibsfreq
        self.iterations = start_iteration
        self.t = 0.0
        self.fixed_dt = fixed_dt
        self.dt = fixed_dt if fixed_dt is not None else (dt if dt is not None else initial_dt(params))
        self.neibslist = None
        self.neibs_info = None
        self.filters = [(k, int(v)) for k, v in (filters or {}).items() if v > 0]   # {"SHEPARD_FILTER"|"MLS_FILTER": frequency}
        self.planes = planes
        self.xsph = None

    def _opts(self, dt):
        if self.params.simflags & capi.ENABLE_XSPH:
            self.xsph = np.zeros((self.n, 4), dtype=np.float32)
        return forces_opts(dt=dt, xsph=self.xsph, planes=self.planes)

    def run_filters(self):
        if self.iterations == 0:
            return
        for kind, freq in self.filters:
            if self.iterations % freq:
                continue
            fn = shepard if kind == "SHEPARD_FILTER" else mls
            self.vel = fn(self.params, self.pos, self.vel, self.info, self.hash, self.cs, self.neibslist)

    def postprocess(self):
        self.vel, _, _ = testpoints(self.params, self.pos, self.vel, self.info, self.hash, self.cs, self.neibslist)

    def build_neibs(self):
        if self.iterations == 0:
            pidx = fix_hash(self.params, self.hash, self.info)
        else:
            pidx = calc_hash(self.params, self.pos, self.hash, self.info)
        sort(self.hash, self.info, pidx)
        self.cs, self.ce, _, self.pos, self.vel, newn = reorder(self.params, self.pos, self.vel, self.info, self.hash, pidx)
        if newn != self.n:
            self.n = newn
            self.pos, self.vel, self.info, self.hash = self.pos[:newn], self.vel[:newn], self.info[:newn], self.hash[:newn]
        self.neibslist, self.neibs_info = build_neibs(self.params, self.pos, self.info, self.hash, self.cs, self.ce)

    def step(self, dt=None):
        if self.iterations % self.buildneibsfreq == 0 or self.neibslist is None:
            self.build_neibs()
        if self.filters:
            self.run_filters()
        dt = self.dt if dt is None else dt
        P = self.params
        f1, cfl1, _ = forces(P, self.pos, self.vel, self.info, self.hash, self.cs, self.neibslist, opts=self._opts(dt / 2))
        pos_s, vel_s = euler(P, self.pos, self.vel, self.info, self.hash, f1, dt / 2, 1, xsph=self.xsph)
        f2, cfl2, _ = forces(P, pos_s, vel_s, self.info, self.hash, self.cs, self.neibslist, opts=self._opts(dt))
        self.pos, self.vel = euler(P, self.pos, self.vel, self.info, self.hash, f2, dt, 2, xsph=self.xsph)
        self.iterations += 1
        self.t += dt
        if self.fixed_dt is None:
            self.dt = min(dtreduce(P, cfl1), dtreduce(P, cfl2))

    def download(self):
        return ParticleArrays(self.pos.copy(), self.vel.copy(), self.info.copy(), self.hash.copy())
