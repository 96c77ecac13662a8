"""CPU compute backend for gpusph_b200.multigpu.SlabWorker built on the oracle (TEST INFRASTRUCTURE).
Lets the slab decomposition / halo exchange host logic run at world_size 2 on gloo without a GPU."""
import ctypes as C

import numpy as np
import torch

import oracle_binding as ob


def _np(t, dtype=None):
    a = t.numpy()
    return a if dtype is None else a.view(dtype)


class OracleBackend:
    def __init__(self, params):
        self.params = params
        self.device = torch.device("cpu")

    def hash_update(self, first, pos, hashv, pidx, info, cdm, n):
        h, i, c = _np(hashv, np.uint32)[:n], _np(info, np.uint16)[:n], _np(cdm, np.uint32)
        if first:
            p = ob.fix_hash(self.params, h, i, c)
        else:
            p = ob.calc_hash(self.params, _np(pos)[:n], h, i, c)
        _np(pidx, np.uint32)[:n] = p

    def sort(self, hashv, info, pidx, n):
        ob.sort(_np(hashv, np.uint32)[:n], _np(info, np.uint16)[:n], _np(pidx, np.uint32)[:n])

    def reorder(self, cs, ce, seg, spos, svel, pos, vel, info, hashv, pidx, n, newn):
        lib = ob.lib()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _np(ce, np.uint32)[:] = 0xFFFFFFFF
        lib.oracle_reorder(p(_np(cs, np.uint32)), p(_np(ce, np.uint32)), p(_np(seg, np.uint32)) if seg is not None else None,
                           p(_np(spos)), p(_np(svel)), p(_np(pos)), p(_np(vel)), p(_np(info, np.uint16)),
                           p(_np(hashv, np.uint32)), p(_np(pidx, np.uint32)), C.c_uint32(n), p(_np(newn, np.uint32)))

    def build_neibs(self, pos, info, hashv, cs, ce, nl, n, range_end):
        lib = ob.lib()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        out = ob.OracleNeibsInfo()
        lib.oracle_build_neibs(C.byref(self.params), p(_np(pos)), p(_np(info, np.uint16)), p(_np(hashv, np.uint32)),
                               p(_np(cs, np.uint32)), p(_np(ce, np.uint32)), p(_np(nl, np.uint16)),
                               C.c_uint32(n), C.c_uint32(range_end), C.byref(out))
        return out

    def forces(self, pos, vel, info, hashv, cs, nl, forces, cfl, n, frm, to, cfl_offset=0):
        lib = ob.lib()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        f = _np(forces)
        f[frm:to] = 0
        return int(lib.oracle_forces(C.byref(self.params), p(_np(pos)), p(_np(vel)), p(_np(info, np.uint16)),
                                     p(_np(hashv, np.uint32)), p(_np(cs, np.uint32)), p(_np(nl, np.uint16)), None, None,
                                     p(f), p(_np(cfl)), None, C.c_uint32(n), C.c_uint32(frm), C.c_uint32(to), C.c_uint32(cfl_offset), None, None, None))

    def dtreduce(self, cfl, nblocks):
        return ob.dtreduce(self.params, _np(cfl)[:nblocks])

    def euler(self, opos, ovel, info, hashv, forces, npos, nvel, n, range_end, dt, step):
        lib = ob.lib()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib.oracle_euler(C.byref(self.params), p(_np(opos)), p(_np(ovel)), p(_np(info, np.uint16)), p(_np(hashv, np.uint32)),
                         p(_np(forces)), p(_np(npos)), p(_np(nvel)), C.c_uint32(n), C.c_uint32(range_end),
                         C.c_float(dt), C.c_int(step), None)

    def fmax_elements(self, n):
        return ((n + 127) // 128 + 3) // 4 * 4


class OracleDeviceDtBackend(OracleBackend):
    """OracleBackend + a host emulation of the device-resident dt record (b200sph_step_* / b200sph_cflmax), so that the
    SlabWorker's device-dt control flow (deferred, once-per-step all-reduce of the CFL maxima) runs on gloo too."""

    def __init__(self, params):
        super().__init__(params)
        self.st = dict(t=0.0, it=0, dt=0.0, dt1=0.0, dt2=0.0)

    def step_set_dt(self, dt):
        self.st.update(dt=float(np.float32(dt)), dt1=float(np.float32(dt)), dt2=float(np.float32(dt)))

    def cflmax(self, cfl, nblocks, out):
        _np(out)[0] = _np(cfl)[:nblocks].max() if nblocks else 0.0

    def dtreduce_async(self, cfl, nblocks, which):
        self.st["dt1" if which == 1 else "dt2"] = ob.dtreduce(self.params, _np(cfl)[:nblocks])

    def euler_async(self, opos, ovel, info, hashv, forces, npos, nvel, n, range_end, step):
        dt = np.float32(self.st["dt"])
        self.euler(opos, ovel, info, hashv, forces, npos, nvel, n, range_end, float(dt / np.float32(2)) if step == 1 else float(dt), step)

    def step_end(self):
        self.st["t"] += self.st["dt"]
        self.st["it"] += 1
        self.st["dt"] = min(self.st["dt1"], self.st["dt2"])

    def step_query(self):
        return self.st["t"], self.st["dt"], self.st["it"]
