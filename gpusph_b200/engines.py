"""Host-side mirror of GPUSPH's three hot engines on top of the C ABI.

The class and method names, the argument meaning and the error behaviour follow the
reference's abstract interfaces so that tests read like the reference's call sites:

* ``NeibsEngine``       <-> AbstractNeibsEngine        (src/engine_neibs.h:45-107)
* ``ForcesEngine``      <-> AbstractForcesEngine       (src/engine_forces.h:42-179)
* ``IntegrationEngine`` <-> AbstractIntegrationEngine  (src/engine_integration.h:40-143)

Buffers are passed like the reference's ``BufferList`` (src/buffer.h:595-775): a mapping from
the reference's buffer keys (``BUFFER_POS`` ...; src/define_buffers.h:78-200) to device arrays —
here torch CUDA tensors, whose only role is to own device memory. ``bufread`` holds inputs,
``bufwrite`` outputs; a missing mandatory buffer raises like the reference does.

All compute happens in ``libb200sph.so`` (hand-written sm_100a CUDA); this file contains no
arithmetic on particle data and no fallback path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import capi

# buffer keys (src/define_buffers.h)
BUFFER_POS = "BUFFER_POS"
BUFFER_VEL = "BUFFER_VEL"
BUFFER_INFO = "BUFFER_INFO"
BUFFER_HASH = "BUFFER_HASH"
BUFFER_PARTINDEX = "BUFFER_PARTINDEX"
BUFFER_CELLSTART = "BUFFER_CELLSTART"
BUFFER_CELLEND = "BUFFER_CELLEND"
BUFFER_NEIBSLIST = "BUFFER_NEIBSLIST"
BUFFER_FORCES = "BUFFER_FORCES"
BUFFER_CFL = "BUFFER_CFL"
BUFFER_CFL_TEMP = "BUFFER_CFL_TEMP"
BUFFER_COMPACT_DEV_MAP = "BUFFER_COMPACT_DEV_MAP"
BUFFER_RB_FORCES = "BUFFER_RB_FORCES"
BUFFER_RB_TORQUES = "BUFFER_RB_TORQUES"
BUFFER_RB_KEYS = "BUFFER_RB_KEYS"
BUFFER_XSPH = "BUFFER_XSPH"
BUFFER_TKE = "BUFFER_TKE"
BUFFER_EPSILON = "BUFFER_EPSILON"

# filter / post-process types (src/particledefine.h: FilterType, PostProcessType)
SHEPARD_FILTER = "SHEPARD_FILTER"
MLS_FILTER = "MLS_FILTER"
TESTPOINTS = "TESTPOINTS"


class BufferList(dict):
    """Keyed device arrays; ``get_ptr`` returns 0 for a missing optional buffer (reference: NULL)."""

    def ptr(self, key: str, mandatory: bool = True) -> int:
        t = self.get(key)
        if t is None:
            if mandatory:
                raise ValueError(f"missing mandatory buffer {key}")   # reference: std::invalid_argument
            return 0
        if not t.is_cuda:
            raise ValueError(f"buffer {key} is not a device array")
        if not t.is_contiguous():
            raise ValueError(f"buffer {key} is not contiguous")
        return t.data_ptr()


class DeviceContext:
    """Per-device engine state = the result of the engines' ``setconstants`` calls."""

    def __init__(self, params: capi.Params, device: torch.device | int | None = None):
        self.lib = capi.load()
        if not torch.cuda.is_available() or self.lib.b200sph_device_count() <= 0:
            raise capi.B200Error("no CUDA device: the B200 engines have no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device if isinstance(device, int) else device.index or 0))
        torch.cuda.set_device(self.device)
        self.params = params.copy()
        h = C.c_void_p()
        capi.check(self.lib.b200sph_create(C.byref(self.params), C.byref(h)))
        self.handle = h
        self.use_stream(torch.cuda.current_stream(self.device))

    def use_stream(self, stream: torch.cuda.Stream) -> None:
        self.stream = stream
        capi.check(self.lib.b200sph_set_stream(self.handle, C.c_void_p(stream.cuda_stream)))

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.b200sph_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NeibsEngine:
    """AbstractNeibsEngine (src/engine_neibs.h:45-107)."""

    def __init__(self, ctx: DeviceContext):
        self.ctx = ctx
        self.lib = ctx.lib

    def getconstants(self) -> int:
        v = C.c_uint32()
        capi.check(self.lib.b200sph_get_neibboundpos(self.ctx.handle, C.byref(v)))
        return v.value

    def resetinfo(self) -> None:
        capi.check(self.lib.b200sph_neibs_resetinfo(self.ctx.handle))

    def getinfo(self) -> capi.NeibsInfo:
        out = capi.NeibsInfo()
        capi.check(self.lib.b200sph_neibs_getinfo(self.ctx.handle, C.byref(out)))
        return out

    def calcHash(self, bufread: BufferList, bufwrite: BufferList, numParticles: int) -> None:
        capi.check(self.lib.b200sph_calc_hash(
            self.ctx.handle, bufwrite.ptr(BUFFER_POS), bufwrite.ptr(BUFFER_HASH), bufwrite.ptr(BUFFER_PARTINDEX),
            bufread.ptr(BUFFER_INFO), bufread.ptr(BUFFER_COMPACT_DEV_MAP, False), numParticles))

    def fixHash(self, bufread: BufferList, bufwrite: BufferList, numParticles: int) -> None:
        capi.check(self.lib.b200sph_fix_hash(
            self.ctx.handle, bufwrite.ptr(BUFFER_HASH, False), bufwrite.ptr(BUFFER_PARTINDEX),
            bufread.ptr(BUFFER_INFO), bufread.ptr(BUFFER_COMPACT_DEV_MAP, False), numParticles))

    def sort(self, bufread: BufferList, bufwrite: BufferList, numParticles: int) -> None:
        capi.check(self.lib.b200sph_sort(
            self.ctx.handle, bufwrite.ptr(BUFFER_HASH), bufwrite.ptr(BUFFER_INFO), bufwrite.ptr(BUFFER_PARTINDEX),
            numParticles))

    def reorderDataAndFindCellStart(self, segmentStart, sorted_buffers: BufferList, unsorted_buffers: BufferList,
                                    numParticles: int, newNumParticles: torch.Tensor,
                                    extra_keys=()) -> None:
        extras = (capi.ReorderExtra * max(len(extra_keys), 1))()
        for i, k in enumerate(extra_keys):
            extras[i].unsorted = unsorted_buffers.ptr(k)
            extras[i].sorted = sorted_buffers.ptr(k)
            extras[i].elem_size = sorted_buffers[k].element_size() * (sorted_buffers[k].shape[1] if sorted_buffers[k].dim() > 1 else 1)
        capi.check(self.lib.b200sph_reorder(
            self.ctx.handle, sorted_buffers.ptr(BUFFER_CELLSTART), sorted_buffers.ptr(BUFFER_CELLEND),
            0 if segmentStart is None else segmentStart.data_ptr(),
            sorted_buffers.ptr(BUFFER_POS), sorted_buffers.ptr(BUFFER_VEL),
            unsorted_buffers.ptr(BUFFER_POS), unsorted_buffers.ptr(BUFFER_VEL),
            extras, len(extra_keys),
            sorted_buffers.ptr(BUFFER_INFO), sorted_buffers.ptr(BUFFER_HASH), sorted_buffers.ptr(BUFFER_PARTINDEX),
            numParticles, newNumParticles.data_ptr()))

    def buildNeibsList(self, bufread: BufferList, bufwrite: BufferList, numParticles: int, particleRangeEnd: int,
                       gridCells: int = 0, sqinfluenceradius: float = 0.0, boundNlSqInflRad: float = 0.0) -> None:
        capi.check(self.lib.b200sph_build_neibs(
            self.ctx.handle, bufread.ptr(BUFFER_POS), bufread.ptr(BUFFER_INFO), bufread.ptr(BUFFER_HASH),
            bufread.ptr(BUFFER_CELLSTART), bufread.ptr(BUFFER_CELLEND), bufwrite.ptr(BUFFER_NEIBSLIST),
            numParticles, particleRangeEnd))


class ForcesEngine:
    """AbstractForcesEngine (src/engine_forces.h:42-179) — the subset on the hot path."""

    def __init__(self, ctx: DeviceContext):
        self.ctx = ctx
        self.lib = ctx.lib

    def setgravity(self, gravity) -> None:
        g = (C.c_float * 3)(*gravity)
        capi.check(self.lib.b200sph_set_gravity(self.ctx.handle, g))

    # texture binding does not exist on this architecture; kept so call sequences match GPUWorker::pre_forces
    def bind_textures(self, bufread: BufferList, numParticles: int, run_mode=None) -> None:
        return None

    def unbind_textures(self, run_mode=None) -> None:
        return None

    def getFmaxElements(self, n: int) -> int:
        return int(self.lib.b200sph_fmax_elements(n))

    def getFmaxTempElements(self, n: int) -> int:
        return int(self.lib.b200sph_fmax_temp_elements(n))

    def round_particles(self, n: int) -> int:
        return int(self.lib.b200sph_round_particles(n))

    def setplanes(self, planes) -> None:
        """planes: sequence of (normal[3], gridPos[3], pos[3]) like plane_t (src/planes.h:42-46)."""
        import numpy as np
        n = len(planes)
        nrm = np.ascontiguousarray(np.asarray([p[0] for p in planes], dtype=np.float32).reshape(n, 3))
        gp = np.ascontiguousarray(np.asarray([p[1] for p in planes], dtype=np.int32).reshape(n, 3))
        pp = np.ascontiguousarray(np.asarray([p[2] for p in planes], dtype=np.float32).reshape(n, 3))
        capi.check(self.lib.b200sph_set_planes(self.ctx.handle, nrm.ctypes.data_as(C.POINTER(C.c_float)),
                                               gp.ctypes.data_as(C.POINTER(C.c_int)), pp.ctypes.data_as(C.POINTER(C.c_float)), n))

    def basicstep(self, bufread: BufferList, bufwrite: BufferList, numParticles: int, fromParticle: int,
                  toParticle: int, cflOffset: int = 0, compute_object_forces: bool = False,
                  step: int = 0, dt: float = 0.0, dt_from_device: bool = False, euler=None, packed=None,
                  new_packed=None) -> int:
        """step / dt = the command's integrator step and dt (src/GPUWorker.cc:1931-1932), read by BREZZI diffusion only;
        dt_from_device: take dt from the context's device-resident record instead.
        euler = (old: BufferList, new: BufferList, step, dt | None): also integrate the same particles
        (AbstractIntegrationEngine::basicstep) right behind their forces; dt None = from the device record.
        packed / new_packed (torch uint8 tensors of 32 bytes per particle, or None): the neighbour records of the state
        in bufread the caller vouches for / the records of the integrated state (b200sph_forces_args.packed,
        b200sph_fused_euler_args.new_packed)."""
        nblocks = C.c_uint32()
        a = capi.ForcesArgs()
        a.pos, a.vel, a.info = bufread.ptr(BUFFER_POS), bufread.ptr(BUFFER_VEL), bufread.ptr(BUFFER_INFO)
        a.hash, a.cell_start, a.neibs_list = bufread.ptr(BUFFER_HASH), bufread.ptr(BUFFER_CELLSTART), bufread.ptr(BUFFER_NEIBSLIST)
        a.forces, a.cfl = bufwrite.ptr(BUFFER_FORCES), bufwrite.ptr(BUFFER_CFL, False)
        a.rb_forces = bufwrite.ptr(BUFFER_RB_FORCES) if compute_object_forces else 0
        a.rb_torques = bufwrite.ptr(BUFFER_RB_TORQUES) if compute_object_forces else 0
        a.xsph = bufwrite.ptr(BUFFER_XSPH, False)
        a.num_particles, a.from_particle, a.to_particle, a.cfl_offset = numParticles, fromParticle, toParticle, cflOffset
        a.dt, a.step, a.dt_from_device = dt, step, 1 if dt_from_device else 0
        a.packed = packed.data_ptr() if packed is not None else 0
        if euler is None:
            capi.check(self.lib.b200sph_forces_ex(self.ctx.handle, C.byref(a), C.byref(nblocks)))
        else:
            # forces + integration of the same particles in one call (b200sph_forces_euler: fused epilogue)
            old, new, estep, edt = euler
            e = capi.FusedEulerArgs()
            e.old_pos, e.old_vel = old.ptr(BUFFER_POS), old.ptr(BUFFER_VEL)
            e.new_pos, e.new_vel = new.ptr(BUFFER_POS), new.ptr(BUFFER_VEL)
            e.step, e.dt, e.dt_from_device = estep, (0.0 if edt is None else edt), (1 if edt is None else 0)
            e.new_packed = new_packed.data_ptr() if new_packed is not None else 0
            capi.check(self.lib.b200sph_forces_euler(self.ctx.handle, C.byref(a), C.byref(e), C.byref(nblocks)))
        return nblocks.value

    def pack_state(self, bufread: BufferList, packed: torch.Tensor, fromParticle: int, toParticle: int) -> None:
        """Interleave POS / VEL of [from, to) into the pair kernel's 32-byte neighbour records (b200sph_pack_state)."""
        capi.check(self.lib.b200sph_pack_state(self.ctx.handle, bufread.ptr(BUFFER_POS), bufread.ptr(BUFFER_VEL),
                                               packed.data_ptr(), fromParticle, toParticle))

    # ---- moving / force-feedback bodies (src/engine_forces.h:62-74) ----
    @staticmethod
    def _farr(a, n, k):
        import numpy as np
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(n, k))
        return a, a.ctypes.data_as(C.POINTER(C.c_float))

    def setrbcg(self, cgGridPos, cgPos, numbodies: int) -> None:
        import numpy as np
        g = np.ascontiguousarray(np.asarray(cgGridPos, dtype=np.int32).reshape(numbodies, 3))
        c, cp = ForcesEngine._farr(cgPos, numbodies, 3)
        capi.check(self.lib.b200sph_set_rbcg(self.ctx.handle, g.ctypes.data_as(C.POINTER(C.c_int)), cp, numbodies))

    def setrbstart(self, rbfirstindex, numbodies: int) -> None:
        import numpy as np
        f = np.ascontiguousarray(np.asarray(rbfirstindex, dtype=np.int32))
        capi.check(self.lib.b200sph_set_rbstart(self.ctx.handle, f.ctypes.data_as(C.POINTER(C.c_int)), numbodies))

    def reduceRbForces(self, bufwrite: BufferList, lastindex, numbodies: int, numBodiesParticles: int):
        import numpy as np
        li = np.ascontiguousarray(np.asarray(lastindex, dtype=np.uint32))
        tf = np.zeros((numbodies, 3), dtype=np.float32)
        tt = np.zeros((numbodies, 3), dtype=np.float32)
        capi.check(self.lib.b200sph_reduce_rb_forces(
            self.ctx.handle, bufwrite.ptr(BUFFER_RB_FORCES), bufwrite.ptr(BUFFER_RB_TORQUES), bufwrite.ptr(BUFFER_RB_KEYS),
            li.ctypes.data_as(C.POINTER(C.c_uint32)), tf.ctypes.data_as(C.POINTER(C.c_float)),
            tt.ctypes.data_as(C.POINTER(C.c_float)), numbodies, numBodiesParticles))
        return tf, tt

    def eos_probe(self, bufread: BufferList, out: torch.Tensor, numParticles: int) -> None:
        """Diagnostic: per-particle {P/rho^2, sound speed} as the forces kernel evaluates them."""
        capi.check(self.lib.b200sph_eos_probe(self.ctx.handle, bufread.ptr(BUFFER_VEL), bufread.ptr(BUFFER_INFO),
                                              out.data_ptr(), numParticles))

    def dtreduce(self, bufread: BufferList, bufwrite: BufferList, numBlocks: int) -> float:
        dt = C.c_float()
        capi.check(self.lib.b200sph_dtreduce(
            self.ctx.handle, bufread.ptr(BUFFER_CFL), bufwrite.ptr(BUFFER_CFL_TEMP, False), numBlocks, C.byref(dt)))
        return dt.value


    # ---- device-resident dt (include/b200sph.h "device-resident time stepping") ----
    def dtreduce_async(self, bufread: BufferList, numBlocks: int, which: int) -> None:
        capi.check(self.lib.b200sph_dtreduce_async(self.ctx.handle, bufread.ptr(BUFFER_CFL), numBlocks, which))

    def step_set_dt(self, dt: float) -> None:
        capi.check(self.lib.b200sph_step_set_dt(self.ctx.handle, C.c_float(dt)))

    def step_end(self) -> None:
        capi.check(self.lib.b200sph_step_end(self.ctx.handle))

    def step_query(self):
        t, dt, it = C.c_double(), C.c_float(), C.c_uint64()
        capi.check(self.lib.b200sph_step_query(self.ctx.handle, C.byref(t), C.byref(dt), C.byref(it)))
        return t.value, dt.value, it.value

    def cflmax(self, bufread: BufferList, numBlocks: int, out: torch.Tensor) -> None:
        """max of the CFL blocks into a device scalar, no synchronisation (multi-GPU: all-reduced on the device)."""
        capi.check(self.lib.b200sph_cflmax(self.ctx.handle, bufread.ptr(BUFFER_CFL), numBlocks, out.data_ptr()))

    def dt_from_cfl(self, max_cfl: float) -> float:
        dt = C.c_float()
        capi.check(self.lib.b200sph_dt_from_cfl(self.ctx.handle, C.c_float(max_cfl), C.byref(dt)))
        return dt.value


class IntegrationEngine:
    """AbstractIntegrationEngine (src/engine_integration.h:40-143) — basicstep."""

    def __init__(self, ctx: DeviceContext):
        self.ctx = ctx
        self.lib = ctx.lib

    def basicstep(self, bufread: BufferList, bufwrite: BufferList, numParticles: int, particleRangeEnd: int,
                  dt: float, step: int, new_packed=None) -> None:
        """new_packed (torch uint8 tensor, 32 bytes per particle) also receives the integrated particles as the pair
        kernel's neighbour records (b200sph_euler_packed)."""
        capi.check(self.lib.b200sph_euler_packed(
            self.ctx.handle, bufread.ptr(BUFFER_POS), bufread.ptr(BUFFER_VEL), bufread.ptr(BUFFER_INFO),
            bufread.ptr(BUFFER_HASH, False), bufread.ptr(BUFFER_FORCES), bufread.ptr(BUFFER_XSPH, False),
            bufwrite.ptr(BUFFER_POS), bufwrite.ptr(BUFFER_VEL), new_packed.data_ptr() if new_packed is not None else 0,
            numParticles, particleRangeEnd, dt, step, 0))

    def unpack_state(self, packed: torch.Tensor, bufwrite: BufferList, fromParticle: int, toParticle: int) -> None:
        """Neighbour records of [from, to) back into POS / VEL (b200sph_unpack_state)."""
        capi.check(self.lib.b200sph_unpack_state(self.ctx.handle, packed.data_ptr(), bufwrite.ptr(BUFFER_POS),
                                                 bufwrite.ptr(BUFFER_VEL), fromParticle, toParticle))


    # ---- moving bodies (src/engine_integration.h:54-68) ----
    def setrbcg(self, cgGridPos, cgPos, numbodies: int) -> None:
        """The integration engine's OWN copy of the centres of gravity (the reference keeps two, see include/b200sph.h)."""
        import numpy as np
        g = np.ascontiguousarray(np.asarray(cgGridPos, dtype=np.int32).reshape(numbodies, 3))
        c, cp = ForcesEngine._farr(cgPos, numbodies, 3)
        capi.check(self.lib.b200sph_set_rbcg_euler(self.ctx.handle, g.ctypes.data_as(C.POINTER(C.c_int)), cp, numbodies))

    def _setf(self, fn, a, numbodies, k):
        arr, ptr = ForcesEngine._farr(a, numbodies, k)
        capi.check(fn(self.ctx.handle, ptr, numbodies))

    def setrbtrans(self, trans, numbodies: int) -> None:
        self._setf(self.lib.b200sph_set_rbtrans, trans, numbodies, 3)

    def setrbsteprot(self, rot, numbodies: int) -> None:
        self._setf(self.lib.b200sph_set_rbsteprot, rot, numbodies, 9)

    def setrblinearvel(self, v, numbodies: int) -> None:
        self._setf(self.lib.b200sph_set_rblinearvel, v, numbodies, 3)

    def setrbangularvel(self, v, numbodies: int) -> None:
        self._setf(self.lib.b200sph_set_rbangularvel, v, numbodies, 3)

    def basicstep_async(self, bufread: BufferList, bufwrite: BufferList, numParticles: int, particleRangeEnd: int,
                        step: int, new_packed=None) -> None:
        """basicstep with dt taken from the context's device-resident record (no host round trip)."""
        capi.check(self.lib.b200sph_euler_packed(
            self.ctx.handle, bufread.ptr(BUFFER_POS), bufread.ptr(BUFFER_VEL), bufread.ptr(BUFFER_INFO),
            bufread.ptr(BUFFER_HASH, False), bufread.ptr(BUFFER_FORCES), bufread.ptr(BUFFER_XSPH, False),
            bufwrite.ptr(BUFFER_POS), bufwrite.ptr(BUFFER_VEL), new_packed.data_ptr() if new_packed is not None else 0,
            numParticles, particleRangeEnd, 0.0, step, 1))


class FilterEngine:
    """AbstractFilterEngine (src/engine_filter.h:40-83) for SHEPARD_FILTER / MLS_FILTER
    (src/cuda/forces.cu:1026-1146): runs every `frequency` iterations, reads VEL of bufread, writes VEL of bufwrite."""

    def __init__(self, ctx: DeviceContext, filtertype: str, frequency: int):
        if filtertype not in (SHEPARD_FILTER, MLS_FILTER):
            raise ValueError(f"unknown filter type {filtertype}")          # reference: std::invalid_argument
        self.ctx, self.lib = ctx, ctx.lib
        self.filtertype = filtertype
        self._frequency = int(frequency)

    def set_frequency(self, frequency: int) -> None:
        self._frequency = int(frequency)

    def frequency(self) -> int:
        return self._frequency

    def process(self, bufread: BufferList, bufwrite: BufferList, numParticles: int, particleRangeEnd: int,
                slength: float = 0.0, influenceradius: float = 0.0) -> None:
        fn = self.lib.b200sph_filter_shepard if self.filtertype == SHEPARD_FILTER else self.lib.b200sph_filter_mls
        capi.check(fn(self.ctx.handle, bufread.ptr(BUFFER_POS), bufread.ptr(BUFFER_VEL), bufwrite.ptr(BUFFER_VEL),
                      bufread.ptr(BUFFER_INFO), bufread.ptr(BUFFER_HASH), bufread.ptr(BUFFER_CELLSTART),
                      bufread.ptr(BUFFER_NEIBSLIST), numParticles, particleRangeEnd))


class PostProcessEngine:
    """AbstractPostProcessEngine (src/engine_postprocess.h:47-116) for TESTPOINTS (src/cuda/post_process.cu:148-216):
    VEL (and TKE / EPSILON when present) of the test points are updated in place in bufwrite."""

    def __init__(self, ctx: DeviceContext, pptype: str = TESTPOINTS, options: int = 0):
        if pptype != TESTPOINTS:
            raise capi.B200Unsupported(f"post-processing {pptype} is out of scope (SURVEY.md section 8)")
        self.ctx, self.lib = ctx, ctx.lib
        self.options = options

    def get_options(self) -> int:
        return self.options

    def get_updated_buffers(self):
        return (BUFFER_VEL, BUFFER_TKE, BUFFER_EPSILON)

    def get_written_buffers(self):
        return ()

    def process(self, bufread: BufferList, bufwrite: BufferList, numParticles: int, particleRangeEnd: int,
                deviceIndex: int = 0, gdata=None) -> None:
        capi.check(self.lib.b200sph_testpoints(
            self.ctx.handle, bufread.ptr(BUFFER_POS), bufwrite.ptr(BUFFER_VEL), bufwrite.ptr(BUFFER_TKE, False),
            bufwrite.ptr(BUFFER_EPSILON, False), bufread.ptr(BUFFER_INFO), bufread.ptr(BUFFER_HASH),
            bufread.ptr(BUFFER_CELLSTART), bufread.ptr(BUFFER_NEIBSLIST), numParticles, particleRangeEnd))


class SimFramework:
    """The bag of engines a problem gets from SETUP_FRAMEWORK (src/simframework.h:65-134)."""

    def __init__(self, params: capi.Params, device=None):
        self.ctx = DeviceContext(params, device)
        self.neibsEngine = NeibsEngine(self.ctx)
        self.forcesEngine = ForcesEngine(self.ctx)
        self.integrationEngine = IntegrationEngine(self.ctx)

    def newFilterEngine(self, filtertype: str, frequency: int) -> FilterEngine:
        """src/cuda/cudasimframework.cu:236-245"""
        return FilterEngine(self.ctx, filtertype, frequency)

    def newPostProcessEngine(self, pptype: str, options: int = 0) -> PostProcessEngine:
        """src/cuda/cudasimframework.cu:210-234"""
        return PostProcessEngine(self.ctx, pptype, options)

    @property
    def params(self) -> capi.Params:
        return self.ctx.params


def neibs_list_rows(nl: torch.Tensor, block: int = 0) -> torch.Tensor:
    """The neighbour list buffer ([neiblistsize, allocated] int16, filled by NeibsEngine.buildNeibsList in the blocked
    layout of include/b200sph.h) re-arranged into the reference's interleaved layout: out[k, i] = entry k of particle i
    (src/cuda/neibs_iteration.cuh:60-75). For inspection and tests; the engines only ever use the blocked layout.
    block = Params.neiblist_block of the context that made the list (0: the default)."""
    rows, A = nl.shape
    B = block or capi.NEIBLIST_BLOCK
    nb, rem = A // B, A % B
    flat = nl.reshape(-1)
    full = flat[:nb * B * rows].view(nb, rows, B).permute(1, 0, 2).reshape(rows, nb * B)
    if rem:
        return torch.cat([full, flat[nb * B * rows:].view(rows, rem)], dim=1)
    return full


def neibs_list_blocked(rows_layout: torch.Tensor, block: int = 0) -> torch.Tensor:
    """Inverse of neibs_list_rows: a list in the reference's layout ([neiblistsize, allocated]) as the engines expect it."""
    rows, A = rows_layout.shape
    B = block or capi.NEIBLIST_BLOCK
    nb, rem = A // B, A % B
    full = rows_layout[:, :nb * B].reshape(rows, nb, B).permute(1, 0, 2).reshape(-1)
    if rem:
        full = torch.cat([full, rows_layout[:, nb * B:].reshape(-1)])
    return full.view(rows, A)
