"""Per-device worker: owns the particle buffers and drives the engines through one time step.

Host-side mirror of the part of GPUWorker / PredictorCorrector that fixes the call order and the
arguments of the hot-path engines (src/GPUWorker.cc:1779-2269 runCommand<CALCHASH|SORT|REORDER|
BUILDNEIBS|FORCES_SYNC|EULER>, src/integrators/PredictorCorrectorIntegrator.cc:387-680,
src/Integrator.cc:93-249, src/GPUSPH.cc:636-699). torch is used for device memory only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import capi
from .engines import (BUFFER_CELLEND, BUFFER_CELLSTART, BUFFER_CFL, BUFFER_CFL_TEMP, BUFFER_COMPACT_DEV_MAP,
                      BUFFER_FORCES, BUFFER_HASH, BUFFER_INFO, BUFFER_NEIBSLIST, BUFFER_PARTINDEX, BUFFER_POS,
                      BUFFER_VEL, BUFFER_XSPH, TESTPOINTS, BufferList, SimFramework)
from .problems import ParticleArrays, initial_dt


# buffers indexed by particle (sliceable to a particle range)
_PER_PARTICLE = (BUFFER_POS, BUFFER_VEL, BUFFER_INFO, BUFFER_HASH, BUFFER_FORCES, BUFFER_XSPH)


def _dev(a: np.ndarray, device, dtype=None) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.view(dtype)
    return t.to(device)


class Worker:
    """One GPU's particle system (states "step n" / "step n*", src/ParticleSystem.h) + engines."""

    def __init__(self, params: capi.Params, particles: ParticleArrays, device=None, *,
                 buildneibsfreq: int = 10, clobber: bool = False, fixed_dt: float | None = None,
                 compact_dev_map: np.ndarray | None = None, start_iteration: int = 0, dt: float | None = None,
                 device_dt: bool = True, filters: dict | None = None, planes=None, graphs: bool = False):
        """filters: {SHEPARD_FILTER | MLS_FILTER: frequency in iterations} (Problem::addFilter);
        planes: [(normal, gridPos, pos)] for ENABLE_PLANES (AbstractForcesEngine::setplanes);
        graphs: replay the command stream of a time step (static between neighbour rebuilds once dt lives on the
        device) as a CUDA graph instead of re-issuing its launches from the host."""
        self.framework = SimFramework(params, device)
        self.params = self.framework.params
        self.device = self.framework.ctx.device
        self.neibs = self.framework.neibsEngine
        self.forces = self.framework.forcesEngine
        self.integration = self.framework.integrationEngine
        self.buildneibsfreq = buildneibsfreq
        self.clobber = clobber
        self.fixed_dt = fixed_dt
        self.allocated = int(self.params.neiblist_stride)
        n = particles.n
        if n > self.allocated:
            raise ValueError("more particles than allocated")
        self.numParticles = n
        self.particleRangeEnd = n
        A, dev = self.allocated, self.device
        f4 = lambda: torch.zeros((A, 4), dtype=torch.float32, device=dev)
        # double-buffered properties (src/predcorr_alloc_policy.cc:41-53)
        self.pos = [f4(), f4()]
        self.vel = [f4(), f4()]
        self.cur = 0                      # index of state "step n"
        self.info = torch.zeros((A, 4), dtype=torch.int16, device=dev)       # ushort4 bits
        self.hash = torch.zeros(A, dtype=torch.int32, device=dev)            # uint32 bits
        self.partindex = torch.zeros(A, dtype=torch.int32, device=dev)
        self.forces_buf = f4()
        # neighbour records of the pair kernel (32 B per particle: pos and vel interleaved, one 256-bit gather each): [0]
        # mirrors state n, [1] the predicted state n*. Handed from launch to launch by the fused integration epilogue;
        # re-made by a streaming pre-pass whenever something else touched the state (state_modified()).
        self.packed = [torch.empty(A * 32, dtype=torch.uint8, device=dev) for _ in range(2)]
        self._packed_valid = False
        ncells = self.params.num_cells
        self.cellstart = torch.empty(ncells, dtype=torch.int32, device=dev)
        self.cellend = torch.empty(ncells, dtype=torch.int32, device=dev)
        self.neibslist = torch.empty((int(self.params.neiblistsize), A), dtype=torch.int16, device=dev)
        self.neibslist.fill_(-1)
        # stripes of the pipelined host-buffer step (step_host) ... of at least this many particles each
        self.host_stripes = min(int(os.environ.get("B200SPH_HOST_STRIPES", "8")), capi.MAX_STRIPES)
        self.host_stripe_min = int(os.environ.get("B200SPH_HOST_STRIPE_MIN", "200000"))
        # striped force evaluations round every stripe's CFL blocks up to a multiple of 4: room for that
        ncfl = 2 * (self.forces.getFmaxElements(A) + 4 * (capi.MAX_STRIPES + 1))
        self.cfl = torch.zeros(ncfl, dtype=torch.float32, device=dev)
        self.cfl_temp = torch.zeros(max(self.forces.getFmaxTempElements(ncfl), 4), dtype=torch.float32, device=dev)
        self.new_num = torch.zeros(1, dtype=torch.int32, device=dev)
        self.compact_dev_map = None if compact_dev_map is None else _dev(compact_dev_map.astype(np.uint32), dev, torch.int32)
        # upload (GPUWorker::uploadSubdomain, src/GPUWorker.cc:1162-1225)
        self.pos[0][:n].copy_(_dev(particles.pos, dev))
        self.vel[0][:n].copy_(_dev(particles.vel, dev))
        self.info[:n].copy_(_dev(particles.info.view(np.int16), dev))
        self.hash[:n].copy_(_dev(particles.hash.view(np.int32), dev))
        self.iterations = start_iteration  # > 0 when resuming from a checkpoint: the first rebuild then uses calcHash
        # adaptive dt lives on the device by default: no host round trip per force evaluation (the reference does two
        # blocking 4-byte readbacks per step); `dt` and `t` are fetched on demand
        self.device_dt = device_dt and fixed_dt is None
        self._t = 0.0
        self._t_offset = 0.0
        self._dt = float(fixed_dt) if fixed_dt is not None else (float(dt) if dt is not None else initial_dt(self.params))
        self._stale = False
        if self.device_dt:
            self.forces.step_set_dt(self._dt)
        # XSPH mean velocity (BUFFER_XSPH exists iff ENABLE_XSPH, src/GPUWorker.cc:135-136)
        self.xsph = f4() if self.params.simflags & capi.ENABLE_XSPH else None
        self.filters = [(self.framework.newFilterEngine(k, v), int(v)) for k, v in (filters or {}).items() if v > 0]
        self.postproc = self.framework.newPostProcessEngine(TESTPOINTS)
        if planes:
            self.forces.setplanes(planes)
        # integration fused into the forces kernel's epilogue (device-dt stepping without XSPH; B200SPH_FUSED_EULER=0: off)
        self.fused = self.device_dt and self.xsph is None and n == self.particleRangeEnd and \
            os.environ.get("B200SPH_FUSED_EULER", "1") != "0"
        self.graphs = bool(graphs) and self.device_dt
        self._graphs = {}                 # (state buffers, numParticles, range end) -> [eager runs so far, CUDAGraph | None]
        self.last_neibs_info = None
        self.total_interactions = 0       # sum over steps of list entries x 2 force evaluations
        self.launches = 0                 # hand-written kernels launched (CUB's sort passes not counted)

    def state_modified(self) -> None:
        """Tell the worker that POS / VEL of the current state were written behind its back (direct tensor writes): the
        pair kernel's neighbour records are re-made before the next force evaluation."""
        self._packed_valid = False

    def _sync_time(self) -> None:
        if self._stale:
            t_dev, self._dt, _ = self.forces.step_query()
            self._t = self._t_offset + t_dev       # the device record counts from 0; a resumed run starts later
            self._stale = False

    @property
    def dt(self) -> float:
        self._sync_time()
        return self._dt

    @dt.setter
    def dt(self, v: float) -> None:
        self._sync_time()
        self._dt = float(v)
        if self.device_dt:
            self.forces.step_set_dt(self._dt)

    @property
    def t(self) -> float:
        self._sync_time()
        return self._t

    @t.setter
    def t(self, v: float) -> None:
        if self.device_dt:
            t_dev = self.forces.step_query()[0]
            self._t_offset = float(v) - t_dev
        self._t = float(v)

    # ---- buffer lists ----
    def _common(self) -> dict:
        d = {BUFFER_INFO: self.info, BUFFER_HASH: self.hash, BUFFER_PARTINDEX: self.partindex,
             BUFFER_CELLSTART: self.cellstart, BUFFER_CELLEND: self.cellend, BUFFER_NEIBSLIST: self.neibslist,
             BUFFER_FORCES: self.forces_buf, BUFFER_CFL: self.cfl, BUFFER_CFL_TEMP: self.cfl_temp}
        if self.compact_dev_map is not None:
            d[BUFFER_COMPACT_DEV_MAP] = self.compact_dev_map
        if self.xsph is not None:
            d[BUFFER_XSPH] = self.xsph
        return d

    def state(self, which: int) -> BufferList:
        b = BufferList(self._common())
        b[BUFFER_POS] = self.pos[which]
        b[BUFFER_VEL] = self.vel[which]
        return b

    # ---- NEIBS_LIST phase (src/Integrator.cc:93-249) ----
    def build_neibs(self, _fenced: bool = False) -> None:
        if not _fenced:
            self.host_fence()
        self._packed_valid = False
        n = self.numParticles
        cur, oth = self.cur, 1 - self.cur
        s = self.state(cur)
        if self.iterations == 0:
            self.neibs.fixHash(s, s, n)            # src/GPUWorker.cc:1800-1807
        else:
            self.neibs.calcHash(s, s, n)           # :1779-1799
        self.neibs.sort(s, s, n)                   # :1811-1827
        self.cellstart.fill_(-1)                   # clobber CELLSTART, :1846
        if self.clobber:
            self.cellend.fill_(-1)
        srt = self.state(oth)
        self.neibs.reorderDataAndFindCellStart(None, srt, s, n, self.new_num)   # :1830-1863
        self.cur = oth
        new_n = int(self.new_num.item())           # DOWNLOAD_NEWNUMPARTS
        if new_n != n:
            self.numParticles = new_n
            self.particleRangeEnd = new_n
        self.neibs.resetinfo()                     # BUILDNEIBS :1866-1905
        if self.clobber:
            self.neibslist.fill_(-1)
        s = self.state(self.cur)
        self.neibs.buildNeibsList(s, s, self.numParticles, self.particleRangeEnd)
        self.last_neibs_info = self.neibs.getinfo()
        self.launches += 6                # calc/fixHash, make_keys, apply_sort, reorder, reset_counters, build_neibs

    # ---- FILTER phases (src/integrators/PredictorCorrectorIntegrator.cc:800-877, 1010-1040) ----
    def run_filters(self) -> None:
        """After NEIBS_LIST, for iterations > 0: every enabled filter whose frequency divides the iteration count reads
        VEL of step n, writes the filtered VEL into the scratch state, and the two VEL buffers are swapped."""
        if self.iterations == 0:
            return
        self._packed_valid = False
        cur, oth = self.cur, 1 - self.cur
        for eng, freq in self.filters:
            if self.iterations % freq:
                continue
            rd = self.state(cur)
            wr = BufferList({BUFFER_VEL: self.vel[oth]})
            eng.process(rd, wr, self.numParticles, self.particleRangeEnd, self.params.slength, self.params.influenceradius)
            self.vel[cur], self.vel[oth] = self.vel[oth], self.vel[cur]
            self.launches += 1

    def postprocess(self) -> None:
        """TESTPOINTS post-processing before a write (src/GPUWorker.cc:2545-2580): in place on the current state."""
        s = self.state(self.cur)
        self._packed_valid = False
        self.postproc.process(s, s, self.numParticles, self.particleRangeEnd)
        self.launches += 1

    # ---- one force evaluation + integration sub-step ----
    def _forces(self, which: int, step: int = 0, dt: float = 0.0) -> float:
        s = self.state(which)
        if self.clobber:
            self.forces_buf.zero_()                # pre_forces: clobber FORCES, src/GPUWorker.cc:1949
        if self.xsph is not None:
            self.xsph.zero_()                      # :1951-1952
        nblocks = self.forces.basicstep(s, s, self.numParticles, 0, self.particleRangeEnd, 0, step=step, dt=dt)
        self.launches += 1                # fused forces kernel
        if self.fixed_dt is not None:
            return self.fixed_dt
        self.launches += 1                # CFL max-reduce
        return self.forces.dtreduce(s, s, nblocks)   # post_forces :1994-2037

    def step(self) -> None:
        """One predictor-corrector time step (src/integrators/PredictorCorrectorIntegrator.cc:917-1068)."""
        self.host_fence()
        if self.iterations % self.buildneibsfreq == 0 or self.last_neibs_info is None:
            self.build_neibs()
        if self.filters:
            self.run_filters()
        n, end = self.numParticles, self.particleRangeEnd
        cur, oth = self.cur, 1 - self.cur
        rd, wr = self.state(cur), self.state(oth)
        if self.device_dt:
            fused = self.fused and n == end
            if fused and not self._packed_valid:
                self.forces.pack_state(rd, self.packed[0], 0, n)
                self.launches += 1
            if self.graphs:
                self._step_graph(cur, n, end)
            else:
                self._enqueue_step(rd, wr, n, end)
            self._packed_valid = fused        # the corrector's epilogue left the records of state n+1 in packed[0]
            self.launches += (2 * 2 if fused else 2 * 3) + 1
            self._stale = True
            if fused:
                oth = cur                 # state n+1 is in the buffers state n was in
        else:
            self._packed_valid = False
            dt = self._dt
            # predictor: forces(n) -> euler step 1 with dt/2 writes n*
            dt1 = self._forces(cur, 1, dt / 2)
            self.integration.basicstep(rd, wr, n, end, dt / 2, 1)
            # corrector: forces(n*) -> euler step 2 with dt, reading pos/vel of n, updating n* in place -> n+1
            dt2 = self._forces(oth, 2, dt)
            self.integration.basicstep(rd, wr, n, end, dt, 2)
            self.launches += 2                # two euler launches
            self._t += dt
            if self.fixed_dt is None:
                self._dt = min(dt1, dt2)      # src/GPUWorker.cc:2224-2229, src/GPUSPH.cc:650-657
        self.cur = oth
        self.iterations += 1
        if self.last_neibs_info is not None:
            self.total_interactions += 2 * int(self.last_neibs_info.num_interactions)

    def _enqueue_step(self, rd: BufferList, wr: BufferList, n: int, end: int) -> None:
        """Everything enqueued, nothing read back: forces(n) -> dt candidate 1 -> euler step 1 (dt/2) -> forces(n*) ->
        dt candidate 2 -> euler step 2 (dt) -> t += dt, dt = min(candidates).

        Fused variant (self.fused): each integration runs in the epilogue of its forces kernel (b200sph_forces_euler),
        the corrector IN PLACE into the state-n buffers (the pair loop still gathers n* from the other ones), so
        state n+1 ends up where state n was and the two states do not swap."""
        fused = self.fused and n == end
        for which, st in ((1, rd), (2, wr)):
            if self.clobber:
                self.forces_buf.zero_()
            if self.xsph is not None:
                self.xsph.zero_()
            if fused:
                # gathers from the records of the state being evaluated; the epilogue writes the records of the state
                # it integrates: n* (predictor) into packed[1], n+1 (corrector) back into packed[0]
                nblocks = self.forces.basicstep(st, st, n, 0, end, 0, step=which, dt_from_device=True,
                                                euler=(rd, wr if which == 1 else rd, which, None),
                                                packed=self.packed[which - 1], new_packed=self.packed[2 - which])
                self.forces.dtreduce_async(st, nblocks, which)
            else:
                nblocks = self.forces.basicstep(st, st, n, 0, end, 0, step=which, dt_from_device=True)
                self.forces.dtreduce_async(st, nblocks, which)
                self.integration.basicstep_async(rd, wr, n, end, which)
        self.forces.step_end()

    def _step_graph(self, cur: int, n: int, end: int) -> None:
        """One time step as a CUDA graph replay. The step's launches depend only on which of the two states is "step n"
        and on the particle counts, so one graph per (state, counts) is captured (after one eager run that fills the
        library's lazily sized scratch) and replayed until the next neighbour rebuild changes the counts."""
        # the buffers a graph was captured with are part of its identity (filters swap the two VEL buffers)
        key = (self.pos[cur].data_ptr(), self.vel[cur].data_ptr(), self.pos[1 - cur].data_ptr(), self.vel[1 - cur].data_ptr(), n, end)
        slot = self._graphs.setdefault(key, [0, None])
        rd, wr = self.state(cur), self.state(1 - cur)
        if slot[1] is None:
            if slot[0] < 1:
                slot[0] += 1
                self._enqueue_step(rd, wr, n, end)
                return
            if len(self._graphs) > 8:            # counts changed at rebuilds: drop graphs of stale shapes
                for k in [k for k in self._graphs if k[4:] != (n, end)]:
                    del self._graphs[k]
            ctx = self.framework.ctx
            main = ctx.stream
            side = torch.cuda.Stream(self.device)
            side.wait_stream(main)
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, stream=side):
                    ctx.use_stream(torch.cuda.current_stream(self.device))
                    self._enqueue_step(rd, wr, n, end)
            finally:
                ctx.use_stream(main)
            main.wait_stream(side)
            slot[1] = g
        slot[1].replay()

    # ---- stepping a state that lives in HOST memory (the e2e path of bench.py) ----
    def _stripes(self) -> list:
        """Particle ranges [a, b) made of whole COORD3 cell layers, so that every neighbour of a particle of stripe s
        lies in stripes s-1, s or s+1 (cells are sorted by hash and COORD3 is the slowest hash digit). Cached per
        neighbour rebuild (one small readback)."""
        key = (self.iterations // self.buildneibsfreq, self.numParticles)
        if getattr(self, "_stripes_key", None) == key:
            return self._stripes_val
        n = self.numParticles
        c = self.params.coord
        S = int(self.params.grid_size[c[0]]) * int(self.params.grid_size[c[1]])
        cs2 = self.cellstart.view(-1, S)
        big = torch.iinfo(torch.int32).max
        first = torch.where(cs2 != -1, cs2, big).min(dim=1).values.cpu().numpy().astype(np.int64)
        starts = sorted(set(int(x) for x in first if x != big and 0 < x < n))
        # small systems: the per-stripe launch / event overhead outweighs the overlap
        want = max(1, min(self.host_stripes, n // self.host_stripe_min))
        bounds = [0]
        for k in range(1, want):
            target = n * k // want
            nxt = next((x for x in starts if x >= target), None)
            if nxt is not None and nxt > bounds[-1]:
                bounds.append(nxt)
        bounds.append(n)
        self._stripes_key, self._stripes_val = key, [(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        return self._stripes_val

    def step_host(self, hpos: torch.Tensor, hvel: torch.Tensor) -> None:
        """One time step of a state owned by the HOST: hpos / hvel (pinned, [allocated, 4] float32, sorted order) hold
        state n on entry and state n+1 once the copies have landed (host_sync() or torch.cuda.synchronize()); nothing
        is synchronised here, and consecutive calls chain on each other stripe by stripe. Results are bitwise those of
        step(). The work is done by the library (b200sph_step_host, csrc/hoststep.cu): uploads, striped force
        evaluations, in-place corrector and downloads pipelined on three streams.

        A step that starts with a neighbour rebuild (1 in buildneibsfreq) needs the whole state before the sort: chained
        upload, rebuild, then the same pipelined step on the (re-sorted) resident state."""
        lib, ctx = self.framework.ctx.lib, self.framework.ctx
        n = self.numParticles
        if not (hpos.is_pinned() and hvel.is_pinned()):
            raise ValueError("step_host needs pinned host buffers")
        rebuild = self.iterations % self.buildneibsfreq == 0 or self.last_neibs_info is None
        # (periodicity along COORD3 makes the first and the last stripe neighbours: the library runs the stripes as a ring)
        if not self.device_dt or self.filters or self.particleRangeEnd != n:
            # configurations the pipelined entry point does not serve: plain upload / step / download
            self.host_fence()
            self.pos[self.cur][:n].copy_(hpos[:n], non_blocking=True)
            self.vel[self.cur][:n].copy_(hvel[:n], non_blocking=True)
            self.step()
            n = self.numParticles
            hpos[:n].copy_(self.pos[self.cur][:n], non_blocking=True)
            hvel[:n].copy_(self.vel[self.cur][:n], non_blocking=True)
            return
        if rebuild:
            capi.check(lib.b200sph_host_upload(ctx.handle, hpos.data_ptr(), hvel.data_ptr(), self.pos[self.cur].data_ptr(),
                                               self.vel[self.cur].data_ptr(), n))
            self.build_neibs(_fenced=True)
            n = self.numParticles
        stripes = self._stripes()
        cur, oth = self.cur, 1 - self.cur
        key = (self._stripes_key, cur, hpos.data_ptr(), hvel.data_ptr())
        if getattr(self, "_host_args_key", None) != key:
            bounds = (C.c_uint32 * (len(stripes) + 1))(*([a for a, _ in stripes] + [stripes[-1][1]]))
            a = capi.HostStepArgs()
            a.host_pos, a.host_vel = hpos.data_ptr(), hvel.data_ptr()
            a.pos, a.vel = self.pos[cur].data_ptr(), self.vel[cur].data_ptr()
            a.pos_star, a.vel_star = self.pos[oth].data_ptr(), self.vel[oth].data_ptr()
            a.info, a.hash = self.info.data_ptr(), self.hash.data_ptr()
            a.cell_start, a.neibs_list = self.cellstart.data_ptr(), self.neibslist.data_ptr()
            a.forces, a.cfl = self.forces_buf.data_ptr(), self.cfl.data_ptr()
            a.xsph = self.xsph.data_ptr() if self.xsph is not None else None
            a.cfl_elements, a.num_particles = self.cfl.numel(), n
            a.stripe_bounds, a.num_stripes = bounds, len(stripes)
            self._host_args_key, self._host_args, self._host_bounds = key, a, bounds
        a = self._host_args
        a.resident = 1 if rebuild else 0
        capi.check(lib.b200sph_step_host(ctx.handle, C.byref(a)))
        self._host_pending = True
        self._packed_valid = False
        self.launches += 2 * len(stripes) * 2 + 4
        self._stale = True
        self.iterations += 1              # state n+1 is in the SAME buffers: self.cur does not flip
        self.total_interactions += 2 * int(self.last_neibs_info.num_interactions)

    def host_fence(self) -> None:
        """The compute stream waits for the copies of earlier step_host calls (before anything else uses the state)."""
        if getattr(self, "_host_pending", False):
            capi.check(self.framework.ctx.lib.b200sph_host_fence(self.framework.ctx.handle))
            self._host_pending = False

    def host_sync(self) -> None:
        """Block until the state written by the last step_host has landed in the host buffers."""
        capi.check(self.framework.ctx.lib.b200sph_host_sync(self.framework.ctx.handle))

    def forces_once(self) -> None:
        """One force evaluation on the current state (bench.py roofline timing): the pair kernel alone, gathering from
        the neighbour records of the current state (made first if something invalidated them)."""
        s = self.state(self.cur)
        if not self._packed_valid:
            self.forces.pack_state(s, self.packed[0], 0, self.numParticles)
            self._packed_valid = True
        self.forces.basicstep(s, s, self.numParticles, 0, self.particleRangeEnd, 0, packed=self.packed[0])

    def euler_once(self) -> None:
        """One predictor sub-step into the scratch state (bench.py: streaming-kernel reference point)."""
        rd, wr = self.state(self.cur), self.state(1 - self.cur)
        self.integration.basicstep(rd, wr, self.numParticles, self.particleRangeEnd, 0.0, 1)

    # ---- checkpoints in the reference's own format (src/writers/HotFile.cc) ----
    def save_hotfile(self, path: str, **layout) -> None:
        """HotWriter equivalent: the current state, iteration count, t and dt; `DamBreak3D --resume <path>` and
        Worker.from_hotfile() both continue from it. layout: buffer_count / order / num_open_boundaries overrides
        (gpusph_b200.hotfile.write_hotfile) for simulations whose host buffer list differs from the plain one."""
        from .hotfile import write_hotfile
        st = self.download()
        write_hotfile(path, st.pos, st.vel, st.info, st.hash, iterations=self.iterations, t=self.t, dt=self.dt, **layout)

    @classmethod
    def from_hotfile(cls, params: capi.Params, path: str, device=None, **kw) -> "Worker":
        """Resume like GPUSPH --resume (src/GPUSPH.cc:393-453): particles, iteration count, t and dt from the file;
        the first step rebuilds the neighbour list with calcHash (iterations > 0)."""
        from .hotfile import particle_arrays, read_hotfile
        hf = read_hotfile(path)
        pos, vel, info, hashv = particle_arrays(hf)
        w = cls(params, ParticleArrays(pos, vel, info, hashv), device, start_iteration=int(hf["iterations"]),
                dt=float(hf["dt"]), **kw)
        w.t = float(hf["t"])
        return w

    # ---- host copies ----
    def download(self) -> ParticleArrays:
        self.host_fence()
        n = self.numParticles
        return ParticleArrays(
            pos=self.pos[self.cur][:n].cpu().numpy(),
            vel=self.vel[self.cur][:n].cpu().numpy(),
            info=self.info[:n].cpu().numpy().view(np.uint16),
            hash=self.hash[:n].cpu().numpy().view(np.uint32))
