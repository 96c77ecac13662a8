"""Multi-GPU: 1-D slab decomposition with one halo cell layer per side.

Host-side mirror of the reference's multi-device machinery, re-done for one process per GPU and
torch.distributed (NCCL over NVLink; gloo for the CPU tests):

* cells are owned by ranks in slabs of whole cell layers along the SLOWEST linearisation axis (COORD3;
  x for the default yzx), so a slab is one contiguous range of cell hashes and, after the sort, one
  contiguous range of particles (reference: ProblemCore::fillDeviceMapByAxis, src/ProblemCore.cc:1061-1116);
* every cell gets a 2-bit type — inner / inner-edge / outer-edge (halo) / outer — in the compact device
  map, OR-ed into hash bits 30-31 by calcHash/fixHash (src/multi_gpu_defines.h:58-72,
  src/GPUWorker.cc:1560-1634, src/cuda/buildneibs_kernel.cu:768-769), so the sort lays every rank's particles
  out as [inner | inner-edge | outer-edge] and reorder reports the segment starts;
* halo copies are integrated locally with the owner's forces, so a particle that crosses a slab face changes
  owner simply because its new hash falls into the neighbour's cells (reference: CROP + APPEND_EXTERNAL at
  every neighbour rebuild, src/Integrator.cc:216-221);
* per force evaluation the owner's FORCES of its inner-edge particles are sent to the neighbour's
  outer-edge range (reference: UPDATE_EXTERNAL(FORCES), PredictorCorrectorIntegrator.cc:514-519) — one
  contiguous range per side and buffer, no packing; dt is the MIN over ranks (src/GPUSPH.cc:650-657).

Because neighbour-list order is (cell, type, id) on every rank, the per-particle summation order does not
depend on the decomposition: N-rank results are bitwise equal to single-rank results.

The numerical work is delegated to a backend: ``CudaBackend`` (the C-ABI engines) in production;
tests plug in a CPU backend to run the same decomposition/exchange logic at world_size 2 on gloo.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import capi
from .problems import ParticleArrays, initial_dt

CELLTYPE_INNER, CELLTYPE_INNER_EDGE, CELLTYPE_OUTER_EDGE, CELLTYPE_OUTER = 0, 1, 2, 3
CELLMASK = 0x3FFFFFFF


def _i32(v: int) -> int:
    """uint32 bit pattern as a python int that fits torch.int32."""
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def slab_partition(params: capi.Params, hashes: np.ndarray, world: int, weights: np.ndarray | None = None) -> list[tuple[int, int]]:
    """Split the COORD3 cell layers into `world` contiguous slabs with balanced (weighted) particle counts
    (reference: fillDeviceMapByAxisBalanced, src/ProblemCore.cc:1119-1200). Every slab gets >= 2 layers.
    `weights` = estimated work per particle (a fluid particle has ~75 pair interactions per force evaluation,
    a DYN boundary particle far fewer)."""
    c3 = params.coord[2]
    G3 = int(params.grid_size[c3])
    S = int(params.grid_size[params.coord[0]]) * int(params.grid_size[params.coord[1]])
    if G3 < 2 * world:
        raise ValueError(f"cannot split {G3} cell layers over {world} devices (need >= 2 layers each)")
    layer = (hashes.astype(np.int64) & CELLMASK) // S
    counts = np.bincount(layer, weights=weights, minlength=G3).astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(counts)])
    total = cum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        x = int(np.searchsorted(cum, target, side="left"))
        x = max(x, bounds[-1] + 2)                    # at least 2 layers for the previous slab
        x = min(x, G3 - 2 * (world - r))              # leave room for the remaining slabs
        bounds.append(x)
    bounds.append(G3)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def compact_device_map(params: capi.Params, slab: tuple[int, int], rank: int, world: int) -> np.ndarray:
    """Per-cell type bits (already shifted to bits 30-31), reference createCompactDeviceMap
    (src/GPUWorker.cc:1560-1634). With periodicity along the split axis the first and the last slab are neighbours: the
    layer below the first one is the last layer of the grid and vice versa."""
    c3 = params.coord[2]
    periodic = bool(params.periodic & (1 << c3))
    G3 = int(params.grid_size[c3])
    S = int(params.grid_size[params.coord[0]]) * int(params.grid_size[params.coord[1]])
    xs, xe = slab
    if periodic and world > 1 and G3 < 2 * world:
        raise ValueError("too few cell layers for a periodic split")
    t = np.full(G3, CELLTYPE_OUTER, dtype=np.uint32)
    t[xs:xe] = CELLTYPE_INNER
    if rank > 0 or (periodic and world > 1):
        t[xs] = CELLTYPE_INNER_EDGE
        t[(xs - 1) % G3] = CELLTYPE_OUTER_EDGE
    if rank < world - 1 or (periodic and world > 1):
        t[xe - 1] = CELLTYPE_INNER_EDGE
        t[xe % G3] = CELLTYPE_OUTER_EDGE
    return np.repeat(t << 30, S).astype(np.uint32)


def state_checksum(info: torch.Tensor, hashv: torch.Tensor, pos: torch.Tensor, vel: torch.Tensor) -> tuple[int, int]:
    """Order-independent 57-bit checksum of a set of particles: sum over particles of a 32-bit mix of (id, cell, the bit
    patterns of pos and vel). The checksums of disjoint sets add up to the checksum of their union, so N ranks can
    all-reduce(SUM) theirs and compare with a single-domain run particle for particle without gathering the state
    (bench.py does that after the warm-up steps of a multi-GPU run). Returns (checksum, particle count). Integer
    arithmetic stays below 2^63: every product is (< 2^32) x (< 2^27)."""
    n = int(pos.shape[0])
    if n == 0:
        return 0, 0
    i16 = info.to(torch.int64) & 0xFFFF
    ids = i16[:, 2] | (i16[:, 3] << 16)
    cell = hashv.to(torch.int64) & CELLMASK
    p = pos.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    v = vel.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    h = (ids * 0x45D9F3B + 0x27D4EB2F) & 0xFFFFFFFF
    for word in (cell, p[:, 0], p[:, 1], p[:, 2], p[:, 3], v[:, 0], v[:, 1], v[:, 2], v[:, 3]):
        h = (((h ^ word) * 0x45D9F3B) & 0xFFFFFFFF)
        h = h ^ (h >> 15)
    return int(h.sum().item()), n


class CudaBackend:
    """Numerical work through the C ABI (gpusph_b200.engines)."""

    def __init__(self, params: capi.Params, device):
        from .engines import SimFramework
        self.fw = SimFramework(params, device)
        self.device = self.fw.ctx.device

    def _b(self, **kw):
        from . import engines as E
        m = {"pos": E.BUFFER_POS, "vel": E.BUFFER_VEL, "info": E.BUFFER_INFO, "hash": E.BUFFER_HASH,
             "pidx": E.BUFFER_PARTINDEX, "cs": E.BUFFER_CELLSTART, "ce": E.BUFFER_CELLEND, "nl": E.BUFFER_NEIBSLIST,
             "forces": E.BUFFER_FORCES, "cfl": E.BUFFER_CFL, "cdm": E.BUFFER_COMPACT_DEV_MAP}
        return E.BufferList({m[k]: v for k, v in kw.items() if v is not None})

    def hash_update(self, first, pos, hashv, pidx, info, cdm, n):
        b = self._b(pos=pos, hash=hashv, pidx=pidx, info=info, cdm=cdm)
        (self.fw.neibsEngine.fixHash if first else self.fw.neibsEngine.calcHash)(b, b, n)

    def sort(self, hashv, info, pidx, n):
        b = self._b(hash=hashv, info=info, pidx=pidx)
        self.fw.neibsEngine.sort(b, b, n)

    def reorder(self, cs, ce, seg, spos, svel, pos, vel, info, hashv, pidx, n, newn):
        srt = self._b(pos=spos, vel=svel, info=info, hash=hashv, pidx=pidx, cs=cs, ce=ce)
        uns = self._b(pos=pos, vel=vel)
        self.fw.neibsEngine.reorderDataAndFindCellStart(seg, srt, uns, n, newn)

    def build_neibs(self, pos, info, hashv, cs, ce, nl, n, range_end):
        b = self._b(pos=pos, info=info, hash=hashv, cs=cs, ce=ce, nl=nl)
        self.fw.neibsEngine.resetinfo()
        self.fw.neibsEngine.buildNeibsList(b, b, n, range_end)
        return self.fw.neibsEngine.getinfo()

    def forces(self, pos, vel, info, hashv, cs, nl, forces, cfl, n, frm, to, cfl_offset=0, packed=None, fused=None):
        """fused = (old_pos, old_vel, new_pos, new_vel, step, new_packed): integrate [frm, to) in the kernel's epilogue with
        the device-resident dt (b200sph_forces_euler)."""
        b = self._b(pos=pos, vel=vel, info=info, hash=hashv, cs=cs, nl=nl, forces=forces, cfl=cfl)
        if fused is None:
            return self.fw.forcesEngine.basicstep(b, b, n, frm, to, cfl_offset, packed=packed)
        opos, ovel, npos, nvel, step, new_packed = fused
        return self.fw.forcesEngine.basicstep(b, b, n, frm, to, cfl_offset, step=step, dt_from_device=True, packed=packed,
                                              euler=(self._b(pos=opos, vel=ovel), self._b(pos=npos, vel=nvel), step, None),
                                              new_packed=new_packed)

    def unpack_state(self, packed, pos, vel, frm, to):
        self.fw.integrationEngine.unpack_state(packed, self._b(pos=pos, vel=vel), frm, to)

    def pack_state(self, pos, vel, packed, frm, to):
        """pos / vel of [frm, to) interleaved into the pair kernel's neighbour records (b200sph_pack_state)."""
        self.fw.forcesEngine.pack_state(self._b(pos=pos, vel=vel), packed, frm, to)

    def dtreduce(self, cfl, nblocks):
        b = self._b(cfl=cfl)
        return self.fw.forcesEngine.dtreduce(b, b, nblocks)

    def cflmax(self, cfl, nblocks, out):
        self.fw.forcesEngine.cflmax(self._b(cfl=cfl), nblocks, out)

    def dt_from_cfl(self, m):
        return self.fw.forcesEngine.dt_from_cfl(m)

    # device-resident dt (no host round trip per force evaluation)
    def dtreduce_async(self, cfl, nblocks, which):
        self.fw.forcesEngine.dtreduce_async(self._b(cfl=cfl), nblocks, which)

    def euler_async(self, opos, ovel, info, hashv, forces, npos, nvel, n, range_end, step, new_packed=None):
        rd = self._b(pos=opos, vel=ovel, info=info, hash=hashv, forces=forces)
        wr = self._b(pos=npos, vel=nvel)
        self.fw.integrationEngine.basicstep_async(rd, wr, n, range_end, step, new_packed=new_packed)

    def step_set_dt(self, dt):
        self.fw.forcesEngine.step_set_dt(dt)

    def step_end(self):
        self.fw.forcesEngine.step_end()

    def step_query(self):
        return self.fw.forcesEngine.step_query()

    def euler(self, opos, ovel, info, hashv, forces, npos, nvel, n, range_end, dt, step):
        rd = self._b(pos=opos, vel=ovel, info=info, hash=hashv, forces=forces)
        wr = self._b(pos=npos, vel=nvel)
        self.fw.integrationEngine.basicstep(rd, wr, n, range_end, dt, step)

    def fmax_elements(self, n):
        return self.fw.forcesEngine.getFmaxElements(n)


class SlabWorker:
    """One rank's particle system in a slab decomposition; same stepping interface as simulation.Worker."""

    def __init__(self, params: capi.Params, particles: ParticleArrays, device=None, *, rank: int, world: int,
                 backend=None, buildneibsfreq: int = 10, fixed_dt: float | None = None, alloc_factor: float = 1.5,
                 group=None):
        self.rank, self.world, self.group = rank, world, group
        self.buildneibsfreq = buildneibsfreq
        self.fixed_dt = fixed_dt
        work = np.where((particles.info[:, 0] & 7) == capi.PT_FLUID, 1.0, 0.25)
        self.slabs = slab_partition(params, particles.hash, world, work)
        self.slab = self.slabs[rank]
        c3 = params.coord[2]
        S = int(params.grid_size[params.coord[0]]) * int(params.grid_size[params.coord[1]])
        self.S = S
        xs, xe = self.slab
        layer = (particles.hash.astype(np.int64) & CELLMASK) // S
        G3 = int(params.grid_size[c3])
        self.periodic_split = bool(params.periodic & (1 << c3)) and world > 1
        keep = np.zeros(G3, dtype=bool)                              # own cell layers + one halo layer per neighbour
        keep[xs:xe] = True
        if rank > 0 or self.periodic_split:
            keep[(xs - 1) % G3] = True
        if rank < world - 1 or self.periodic_split:
            keep[xe % G3] = True
        sel = np.flatnonzero(keep[layer])
        n = sel.shape[0]
        self.allocated = int(n * alloc_factor) + 1024
        p = params.copy()
        p.neiblist_stride = self.allocated
        self.params = p
        # backend: None -> the CUDA engines; or a factory called with this rank's params (stride = local allocation)
        self.backend = backend(p) if backend is not None else CudaBackend(p, device)
        dev = self.backend.device
        self.device = dev
        A = self.allocated
        f4 = lambda: torch.zeros((A, 4), dtype=torch.float32, device=dev)
        self.pos, self.vel, self.cur = [f4(), f4()], [f4(), f4()], 0
        self.info = torch.zeros((A, 4), dtype=torch.int16, device=dev)
        self.hash = torch.zeros(A, dtype=torch.int32, device=dev)
        self.partindex = torch.zeros(A, dtype=torch.int32, device=dev)
        self.forces_buf = f4()
        # neighbour records of the pair kernel (CUDA engines): made once per force evaluation for own + halo particles,
        # then gathered by both stripes' launches
        # one per state buffer: packed[i] mirrors (pos[i], vel[i]) of own + halo particles
        self.packed = [torch.empty(A * 32, dtype=torch.uint8, device=dev) for _ in range(2)] if hasattr(self.backend, "pack_state") else None
        self._pending_x = [[], []]        # halo updates in flight, per state buffer
        self._xops = {}                   # cached P2P op lists, per state buffer (rebuilt with the neighbour list)
        nc = p.num_cells
        self.cellstart = torch.empty(nc, dtype=torch.int32, device=dev)
        self.cellend = torch.empty(nc, dtype=torch.int32, device=dev)
        self.neibslist = torch.full((int(p.neiblistsize), A), -1, dtype=torch.int16, device=dev)
        self.cfl = torch.zeros(self.backend.fmax_elements(A), dtype=torch.float32, device=dev)
        self.segments = torch.zeros(4, dtype=torch.int32, device=dev)
        self.cfl_scalar = torch.zeros(1, dtype=torch.float32, device=dev)
        # device-dt path: local CFL maxima of the two force evaluations of a step, and the copy that is all-reduced
        # (asynchronously, once per step) while the next step's first force evaluation is already running
        self.cfl_local = torch.zeros(2, dtype=torch.float32, device=dev)
        self.cfl_global = torch.zeros(2, dtype=torch.float32, device=dev)
        self._pending_dt = None
        self.new_num = torch.zeros(1, dtype=torch.int32, device=dev)
        self.cdm = torch.from_numpy(compact_device_map(p, self.slab, rank, world).view(np.int32)).to(dev)
        self.pos[0][:n].copy_(torch.from_numpy(particles.pos[sel]).to(dev))
        self.vel[0][:n].copy_(torch.from_numpy(particles.vel[sel]).to(dev))
        self.info[:n].copy_(torch.from_numpy(particles.info[sel].view(np.int16)).to(dev))
        self.hash[:n].copy_(torch.from_numpy(particles.hash[sel].view(np.int32)).to(dev))
        self.numParticles = n          # own + halo
        self.numOwn = 0                # set by the first rebuild
        self.iterations = 0
        self._t = 0.0
        self._dt = float(fixed_dt) if fixed_dt is not None else initial_dt(p)
        self._stale = False
        # adaptive dt stays on the device when the backend supports it (CUDA engines): CFL maxima are reduced on the
        # device, all-reduced (MAX) over NCCL on the device and turned into dt on the device
        self.device_dt = fixed_dt is None and hasattr(self.backend, "dtreduce_async")
        if self.device_dt:
            self.backend.step_set_dt(self._dt)
        self.last_neibs_info = None
        self.total_interactions = 0
        self.launches = 0
        # ranges for the per-evaluation force exchange: (start, count)
        self.edge_left = self.edge_right = self.halo_left = self.halo_right = (0, 0)
        # CUDA engines only: edge stripe + its exchange on a high-priority stream next to the inner stripe
        # (measured at 2 GPUs: 1.47 -> 1.38 ms per step; B200SPH_SLAB_EDGE_STREAM=0 turns it off)
        import os
        self._edge_stream = None
        # both half-steps integrate in the pair kernel's epilogue (B200SPH_SLAB_FUSED_PREDICTOR=0: the predictor in a
        # separate streaming launch behind the pair kernels, so that the dt all-reduce overlaps with them)
        self.fused_predictor = os.environ.get("B200SPH_SLAB_FUSED_PREDICTOR", "1") != "0"
        if isinstance(self.backend, CudaBackend) and os.environ.get("B200SPH_SLAB_EDGE_STREAM", "1") != "0":
            self._edge_stream = torch.cuda.Stream(self.device, priority=-1)

    def _mark(self, label: str, stream=None) -> None:
        """Diagnostics (tools/diag_slab_*.py): with self._trace a list, record a timed event on `stream` (default: the
        current one) together with the host clock at enqueue time."""
        trace = getattr(self, "_trace", None)
        if trace is None:
            return
        import time
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream if stream is not None else torch.cuda.current_stream(self.device))
        trace.append((self.iterations, label, ev, time.perf_counter()))

    def _finish_dt(self):
        """Complete the step whose dt candidates are still being all-reduced: dt = f(global CFL maxima), t += dt."""
        if self._pending_dt is None:
            return
        if self._pending_dt is not True:
            self._pending_dt.wait()
        be = self.backend
        be.dtreduce_async(self.cfl_global[0:1], 1, 1)
        be.dtreduce_async(self.cfl_global[1:2], 1, 2)
        be.step_end()
        self.launches += 3
        self._pending_dt = None

    def _sync_time(self):
        self._finish_dt()
        if self._stale:
            self._t, self._dt, _ = self.backend.step_query()
            self._stale = False

    @property
    def dt(self):
        self._sync_time()
        return self._dt

    @property
    def t(self):
        self._sync_time()
        return self._t

    # ------------------------------------------------------------------ communication helpers
    def _left(self):
        if self.periodic_split:
            return (self.rank - 1) % self.world
        return self.rank - 1 if self.rank > 0 else None

    def _right(self):
        if self.periodic_split:
            return (self.rank + 1) % self.world
        return self.rank + 1 if self.rank < self.world - 1 else None

    def _halo_pairs(self, t, edge_left=None, edge_right=None, halo_left=None, halo_right=None):
        """(sends, recvs) moving the edge layers of array `t` (row = particle) into the neighbours' halo ranges. Sends
        are posted right neighbour first, receives left neighbour first: when both neighbours are the SAME rank (two
        slabs of a periodic split) messages between two ranks match in posting order, and my left halo is that rank's
        right edge."""
        el, er = edge_left or self.edge_left, edge_right or self.edge_right
        hl, hr = halo_left or self.halo_left, halo_right or self.halo_right
        sends, recvs = [], []
        if self._right() is not None:
            sends.append((t[er[0]:er[0] + er[1]], self._right()))
        if self._left() is not None:
            sends.append((t[el[0]:el[0] + el[1]], self._left()))
            recvs.append((t[hl[0]:hl[0] + hl[1]], self._left()))
        if self._right() is not None:
            recvs.append((t[hr[0]:hr[0] + hr[1]], self._right()))
        return sends, recvs

    def _exchange(self, sends, recvs):
        """sends/recvs: lists of (tensor, peer). One batched NCCL group (reference: transferBursts)."""
        # raw bytes: NCCL has no int16, and the payload is opaque to the transport anyway
        for w in self._exchange_start(sends, recvs):
            w.wait()

    def _exchange_start(self, sends, recvs):
        """Enqueue the transfers and return the work handles (NCCL: they run on the communicator's stream after the
        work already enqueued on the compute stream, so kernels launched next overlap with them)."""
        ops = [dist.P2POp(dist.isend, t.view(torch.uint8), peer, self.group) for t, peer in sends if t.numel()]
        ops += [dist.P2POp(dist.irecv, t.view(torch.uint8), peer, self.group) for t, peer in recvs if t.numel()]
        return dist.batch_isend_irecv(ops) if ops else []

    def _exchange_counts(self, n_to_left: int, n_to_right: int):
        dev = self.device
        out = {}
        sends, recvs = [], []
        sl = torch.tensor([n_to_left], dtype=torch.int64, device=dev)
        sr = torch.tensor([n_to_right], dtype=torch.int64, device=dev)
        rl = torch.zeros(1, dtype=torch.int64, device=dev)
        rr = torch.zeros(1, dtype=torch.int64, device=dev)
        if self._right() is not None:
            sends.append((sr, self._right()))
        if self._left() is not None:
            sends.append((sl, self._left())); recvs.append((rl, self._left()))
        if self._right() is not None:
            recvs.append((rr, self._right()))
        self._exchange(sends, recvs)
        out["from_left"] = int(rl.item()) if self._left() is not None else 0
        out["from_right"] = int(rr.item()) if self._right() is not None else 0
        return out

    # ------------------------------------------------------------------ neighbour rebuild
    def build_neibs(self) -> None:
        be = self.backend
        n = self.numParticles
        cur, oth = self.cur, 1 - self.cur
        # halo updates still in flight land first; with neighbour records the halo particles only live as records
        # between rebuilds: bring them back into pos / vel, where the hash update below looks for particles that
        # crossed the slab face (that is how ownership changes, src/Integrator.cc:216-221)
        self._mark("rebuild: begin")
        self._wait_halo(0)
        self._wait_halo(1)
        self._xops = {}
        records = self.packed is not None and self.device_dt
        if records and self.iterations > 0 and n > self.numOwn:
            be.unpack_state(self.packed[cur], self.pos[cur], self.vel[cur], self.numOwn, n)
        # CALCHASH (with the compact device map) + SORT + REORDER  (src/Integrator.cc:93-160)
        be.hash_update(self.iterations == 0, self.pos[cur], self.hash, self.partindex, self.info, self.cdm, n)
        be.sort(self.hash, self.info, self.partindex, n)
        self.cellstart.fill_(-1)
        be.reorder(self.cellstart, self.cellend, self.segments, self.pos[oth], self.vel[oth], self.pos[cur], self.vel[cur],
                   self.info, self.hash, self.partindex, n, self.new_num)
        self.cur = cur = oth
        oth = 1 - cur
        self._mark("rebuild: hash+sort+reorder")
        # one packed readback: segment starts, active count and the particle ranges of the two edge layers
        # (contiguous: the slab axis is the slowest hash digit, so a cell layer is one run of cells)
        xs, xe = self.slab
        cs2, ce2 = self.cellstart.view(-1, self.S), self.cellend.view(-1, self.S)
        big = torch.iinfo(torch.int32).max

        def layer_range_dev(x):
            lcs = cs2[x]
            valid = lcs != -1
            return torch.stack([torch.where(valid, lcs, big).min(), torch.where(valid, ce2[x], 0).max()])
        packed = torch.cat([self.segments, self.new_num, layer_range_dev(xs), layer_range_dev(xe - 1)]).cpu().numpy()
        seg = packed[:4].view(np.uint32)
        n_active = int(packed[4])
        # CROP: keep what this rank owns = [inner | inner-edge]; drop stale halo copies and outer particles
        first_ext = min([int(s_) for s_ in seg[2:4] if s_ != 0xFFFFFFFF] + [n_active])
        n_own = first_ext

        def rng(a, b_):
            a, b_ = int(a), int(b_)
            return (a, b_ - a) if b_ > a and a != big else (0, 0)
        self.edge_left = rng(packed[5], packed[6]) if self._left() is not None else (0, 0)
        self.edge_right = rng(packed[7], packed[8]) if self._right() is not None else (0, 0)
        # first inner-edge particle: [0, edge_start) is the inner stripe, [edge_start, n_own) the edge stripe
        self.edge_start = int(seg[1]) if seg[1] != 0xFFFFFFFF else n_own
        # APPEND_EXTERNAL: fresh halo copies from the owners (pos, vel, info, hash)
        self._mark("rebuild: readback")
        cnt = self._exchange_counts(self.edge_left[1], self.edge_right[1])
        nl_, nr_ = cnt["from_left"], cnt["from_right"]
        self._mark("rebuild: counts exchanged")
        if n_own + nl_ + nr_ > self.allocated:
            raise MemoryError(f"rank {self.rank}: {n_own + nl_ + nr_} particles exceed the allocation {self.allocated}")
        self.halo_left = (n_own, nl_)
        self.halo_right = (n_own + nl_, nr_)
        sends, recvs = [], []
        for buf in (self.pos[cur], self.vel[cur], self.info, self.hash):
            s_, r_ = self._halo_pairs(buf)
            sends += s_; recvs += r_
        self._exchange(sends, recvs)
        n = n_own + nl_ + nr_
        # received hashes carry the sender's INNER_EDGE bits: they are OUTER_EDGE here
        ext = self.hash[n_own:n]
        ext.copy_(torch.bitwise_or(torch.bitwise_and(ext, CELLMASK), _i32(CELLTYPE_OUTER_EDGE << 30)))
        self.numParticles, self.numOwn = n, n_own
        # cell ranges over own + halo (the concatenation is already sorted): identity permutation through reorder
        self.partindex[:n].copy_(torch.arange(n, dtype=torch.int32, device=self.device))
        self.cellstart.fill_(-1)
        be.reorder(self.cellstart, self.cellend, self.segments, self.pos[oth], self.vel[oth], self.pos[cur], self.vel[cur],
                   self.info, self.hash, self.partindex, n, self.new_num)
        self.cur = oth
        self._mark("rebuild: halo appended + cell ranges")
        # BUILDNEIBS for the particles this rank owns (their neighbours include the halo)
        self.last_neibs_info = be.build_neibs(self.pos[self.cur], self.info, self.hash, self.cellstart, self.cellend,
                                              self.neibslist, n, n_own)
        if records:
            be.pack_state(self.pos[self.cur], self.vel[self.cur], self.packed[self.cur], 0, n)
        self._mark("rebuild: list built")
        self.launches += 9

    # ------------------------------------------------------------------ time stepping
    # What crosses the slab faces is the STATE, not the forces. The reference sends the owner's FORCES of the edge layer
    # to the neighbour, which then integrates its halo copies itself (UPDATE_EXTERNAL, src/GPUWorker.cc:2086-2160);
    # SURVEY.md section 8e names the alternative used here: the owner integrates its particles - in the epilogue of the
    # pair kernel - and sends the integrated edge layer (one 32-byte {pos, vel} record per particle instead of a
    # 16-byte force) into the neighbour's halo range. Halo copies are never integrated locally, so there is no
    # integration launch over them and no forces exchange sitting between the two stripes of a force evaluation; the
    # values a rank sees for its halo are the owner's bits, so the run stays bitwise the single-GPU run.
    def _halo_tensors(self, which: int):
        """The arrays a halo update of state buffer `which` carries (row = one particle)."""
        if self.packed is not None:
            return [self.packed[which].view(-1, 32)]
        return [self.pos[which], self.vel[which]]

    def _start_halo_update(self, which: int):
        """Owner's edge layers -> neighbours' halo ranges for state buffer `which`: one batched NCCL group, enqueued
        behind the work already on the current stream. The op list is cached until the next rebuild."""
        ops = self._xops.get(which)
        if ops is None:
            sends, recvs = [], []
            for t in self._halo_tensors(which):
                s_, r_ = self._halo_pairs(t)
                sends += s_; recvs += r_
            ops = [dist.P2POp(dist.isend, t.view(torch.uint8), peer, self.group) for t, peer in sends if t.numel()]
            ops += [dist.P2POp(dist.irecv, t.view(torch.uint8), peer, self.group) for t, peer in recvs if t.numel()]
            self._xops[which] = ops
        return dist.batch_isend_irecv(ops) if ops else []

    def _wait_halo(self, which: int):
        for w_ in self._pending_x[which]:
            w_.wait()
        self._pending_x[which] = []

    def _cfl_candidate(self, nblocks: int, cand: int) -> float:
        """dt candidate of one force evaluation (local CFL maximum; combined over ranks by the caller's scheme)."""
        be, n_own = self.backend, self.numOwn
        if self.fixed_dt is not None:
            return self.fixed_dt
        if self.device_dt:
            # local maximum only: the two maxima of a step are all-reduced together at the end of the step (step())
            mine = self.cfl_local[cand - 1:cand]
            if n_own > 0:
                be.cflmax(self.cfl, nblocks, mine)
            else:
                mine.zero_()
            self.launches += 1
            return 0.0
        # dt = min over ranks (src/GPUSPH.cc:650-657) = dt(max over ranks of the CFL maxima): the block maxima are
        # reduced on the device, all-reduced(MAX) on the device, and read back once
        if hasattr(be, "cflmax"):
            if n_own > 0:
                be.cflmax(self.cfl, nblocks, self.cfl_scalar)
            else:
                self.cfl_scalar.zero_()
            dist.all_reduce(self.cfl_scalar, op=dist.ReduceOp.MAX, group=self.group)
            self.launches += 1
            return be.dt_from_cfl(float(self.cfl_scalar.item()))
        dt = be.dtreduce(self.cfl, nblocks) if n_own > 0 else float("inf")
        t = torch.tensor([dt], dtype=torch.float32, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return float(t.item())

    def _forces(self, which: int, cand: int = 0, fused=None) -> float:
        """One force evaluation of the particles this rank owns, on state buffer `which`.
        fused = (old, new, step): integrate them in the kernel's epilogue (device dt) from state buffer `old` into state
        buffer `new` (`new` == `old`: in place), the integrated edge layer leaving for the neighbours as soon as the edge
        stripe's launch is done."""
        be = self.backend
        n, n_own = self.numParticles, self.numOwn
        e0 = min(self.edge_start, n_own)
        args = (self.pos[which], self.vel[which], self.info, self.hash, self.cellstart, self.neibslist, self.forces_buf, self.cfl, n)
        kw = {}
        if self.packed is not None:
            kw["packed"] = self.packed[which]
        kwf = dict(kw)
        if fused is not None:
            old, new, step = fused
            kwf["fused"] = (self.pos[old], self.vel[old], self.pos[new], self.vel[new], step, self.packed[new])
        if self._edge_stream is not None and n_own > e0 > 0:
            # striping (reference: --striping): the EDGE stripe (one cell layer, a quarter of a wave of CTAs) runs on its own
            # high-priority stream next to the INNER stripe's grid; it is the only one that reads the halo, so it alone
            # waits for the halo update still in flight for this state buffer
            main, es = torch.cuda.current_stream(self.device), self._edge_stream
            ctx = be.fw.ctx
            es.wait_stream(main)
            with torch.cuda.stream(es):
                self._wait_halo(which)
            try:
                ctx.use_stream(es)
                nb_edge = be.forces(*args, e0, n_own, 0, **kwf)
            finally:
                ctx.use_stream(main)
            if fused is not None:
                with torch.cuda.stream(es):
                    self._pending_x[fused[1]] = self._start_halo_update(fused[1])
            nb_inner = be.forces(*args, 0, e0, nb_edge, **kwf)
            main.wait_stream(es)          # the edge stripe's kernel (the transfer runs on the communicator's stream)
            self.launches += 2
        else:
            self._wait_halo(which)
            if fused is not None and n_own > e0:
                nb_edge = be.forces(*args, e0, n_own, 0, **kwf)
                self._pending_x[fused[1]] = self._start_halo_update(fused[1])
                nb_inner = be.forces(*args, 0, e0, nb_edge, **kwf) if e0 > 0 else 0
                nblocks_ = nb_edge + nb_inner
                self.launches += 2
                return self._finish_forces(nblocks_, cand)
            nb_edge, nb_inner = (be.forces(*args, 0, n_own, 0, **kwf) if n_own > 0 else 0), 0
            if fused is not None:
                self._pending_x[fused[1]] = self._start_halo_update(fused[1])
            self.launches += 1
        return self._finish_forces(nb_edge + nb_inner, cand)

    def _finish_forces(self, nblocks: int, cand: int) -> float:
        return self._cfl_candidate(nblocks, cand)

    def step(self) -> None:
        if self.iterations % self.buildneibsfreq == 0 or self.last_neibs_info is None:
            self.build_neibs()
        be = self.backend
        n, n_own = self.numParticles, self.numOwn
        cur, oth = self.cur, 1 - self.cur
        # integration of the particles this rank OWNS (numParticles = n, range end = n_own)
        eargs = (self.pos[cur], self.vel[cur], self.info, self.hash, self.forces_buf, self.pos[oth], self.vel[oth], n, n_own)
        if self.device_dt and self.packed is not None:
            # production path (CUDA engines): neighbour records for both states, corrector fused, nothing read back.
            # dt of a step only depends on the CFL maxima of the PREVIOUS step (src/GPUSPH.cc:636-699) and nothing before
            # the first integration needs it: the all-reduce (one per step, both maxima) overlaps with the predictor's
            # pair kernels, which is why the predictor integrates in a separate (streaming) launch and the corrector,
            # whose dt is known by then, in the pair kernel's epilogue.
            self._mark("step: begin")
            if self.fused_predictor:
                # dt first (the all-reduce was started at the end of the previous step and has had the tail of that
                # step's corrector to finish), then both half-steps integrate in the pair kernel's epilogue
                self._finish_dt()
                self._forces(cur, 1, fused=(cur, oth, 1))
                self._mark("predictor forces+euler")
            else:
                self._forces(cur, 1)
                self._mark("predictor forces")
                self._finish_dt()
                be.euler_async(*eargs, 1, new_packed=self.packed[oth])
                self._mark("dt + predictor euler")
                self._pending_x[oth] = self._start_halo_update(oth)
            self._forces(oth, 2, fused=(cur, cur, 2))    # state n+1 lands IN PLACE in the buffers of state n
            self._mark("corrector forces+euler")
            self.cfl_global.copy_(self.cfl_local)
            self._pending_dt = dist.all_reduce(self.cfl_global, op=dist.ReduceOp.MAX, group=self.group, async_op=True) or True
            self._stale = True
            self._mark("step: end")
            self.launches += 1
            oth = cur
        elif self.device_dt:
            self._forces(cur, 1)
            self._finish_dt()
            be.euler_async(*eargs, 1)
            self._pending_x[oth] = self._start_halo_update(oth)
            self._forces(oth, 2)
            be.euler_async(*eargs, 2)
            self._pending_x[oth] = self._start_halo_update(oth)
            self.cfl_global.copy_(self.cfl_local)
            self._pending_dt = dist.all_reduce(self.cfl_global, op=dist.ReduceOp.MAX, group=self.group, async_op=True) or True
            self._stale = True
            self.launches += 2
        else:
            dt = self._dt
            if self.packed is not None:
                be.pack_state(self.pos[cur], self.vel[cur], self.packed[cur], 0, n)
            dt1 = self._forces(cur)
            be.euler(*eargs, dt / 2, 1)
            self._pending_x[oth] = self._start_halo_update_arrays(oth)
            self._wait_halo(oth)
            if self.packed is not None:
                be.pack_state(self.pos[oth], self.vel[oth], self.packed[oth], 0, n)
            dt2 = self._forces(oth)
            be.euler(*eargs, dt, 2)
            self._pending_x[oth] = self._start_halo_update_arrays(oth)
            self._wait_halo(oth)
            self._t += dt
            if self.fixed_dt is None:
                self._dt = min(dt1, dt2)
            self.launches += 2
        self.cur = oth
        self.iterations += 1
        self.total_interactions += 2 * int(self.last_neibs_info.num_interactions)

    def _start_halo_update_arrays(self, which: int):
        """Host-dt path with the CUDA engines: the halo update carries the pos / vel arrays themselves."""
        sends, recvs = [], []
        for t in (self.pos[which], self.vel[which]):
            s_, r_ = self._halo_pairs(t)
            sends += s_; recvs += r_
        return self._exchange_start(sends, recvs)

    def state_modified(self) -> None:
        """POS / VEL of the particles this rank owns were written behind the worker's back (a host upload): refresh their
        neighbour records. The halo records stay: they are the neighbours' current states, delivered by the exchange."""
        if self.packed is not None and self.device_dt and self.numOwn > 0:
            self.backend.pack_state(self.pos[self.cur], self.vel[self.cur], self.packed[self.cur], 0, self.numOwn)
            self.launches += 1

    # ---- stepping a state that lives in HOST memory (bench.py's e2e leg on N > 1 GPUs) ----
    def _inner_stripes(self, max_pieces: int = 24):
        """The inner stripe [0, edge_start) cut into ranges of whole cell layers (every neighbour of a particle of range k
        lies in ranges k-1 .. k+1, in the edge stripe or in the halo). Cached per neighbour rebuild (one small readback).
        A range holds about B200SPH_SLAB_HOST_PIECE particles (default 350 000, ~11 MB per copy direction): what the
        pipeline cannot overlap is one range's download + upload + pair kernels, so ranges are kept small, but large
        enough for a copy to run at the link's rate and for the pair kernel to fill the GPU (at most `max_pieces`)."""
        key = (self.iterations // self.buildneibsfreq, self.numOwn, self.edge_start)
        if getattr(self, "_istripes_key", None) == key:
            return self._istripes_val
        e0 = min(self.edge_start, self.numOwn)
        cs2 = self.cellstart.view(-1, self.S)
        big = torch.iinfo(torch.int32).max
        first = torch.where(cs2 != -1, cs2, big).min(dim=1).values.cpu().numpy().astype(np.int64)
        starts = sorted(set(int(x) for x in first if x != big and 0 < x < e0))
        piece = max(1, int(os.environ.get("B200SPH_SLAB_HOST_PIECE", "350000")))
        want = max(1, min(max_pieces, e0 // piece))
        bounds = [0]
        for k in range(1, want):
            target = e0 * k // want
            nxt = next((x for x in starts if x >= target), None)
            if nxt is not None and nxt > bounds[-1]:
                bounds.append(nxt)
        bounds.append(e0)
        self._istripes_key, self._istripes_val = key, [(a, b_) for a, b_ in zip(bounds[:-1], bounds[1:]) if b_ > a]
        return self._istripes_val

    def step_host(self, hpos: torch.Tensor, hvel: torch.Tensor, chunks: int = 8) -> None:
        """One time step of the particles this rank owns with the state held by the HOST: hpos / hvel (pinned, sorted
        order) hold state n of [0, numOwn) on entry and state n+1 once the copies have landed (host_fence()). Nothing is
        synchronised; consecutive calls chain piece by piece through events. Halo copies are not the host's business:
        they arrive from the neighbours over NVLink as usual. Results are bitwise those of step().

        Between neighbour rebuilds the step is pipelined with its copies: the owned particles are cut into the edge
        stripe and a few inner ranges of whole cell layers; the predictor's pair kernel of range k starts when range
        k+1 has been uploaded, the corrector (integration fused, in place) of range k is followed at once by its download,
        and the upload of a piece of step n+1 waits only for the download of the same piece of step n - both PCIe
        directions and the SMs are busy at the same time. A step that starts with a rebuild needs the whole state
        first: its copies run in `chunks` pieces around the ordinary step()."""
        if not (hpos.is_pinned() and hvel.is_pinned()):
            raise ValueError("step_host needs pinned host buffers")
        dev = self.device
        if getattr(self, "_up", None) is None:
            self._up, self._down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            self._upE, self._downE = torch.cuda.Stream(dev), torch.cuda.Stream(dev)   # the edge stripe's own copy lanes
            self._down_ev, self._down_key = {}, None
        rebuild = self.iterations % self.buildneibsfreq == 0 or self.last_neibs_info is None
        records = self.device_dt and self.packed is not None and self._edge_stream is not None
        if not records or self.numOwn == 0:
            return self._step_host_plain(hpos, hvel, chunks)
        mark = self._mark
        resident = False
        if rebuild:
            # the sort needs the whole state: upload it (behind every earlier download), rebuild, then run the same
            # pipelined step on the resident, re-sorted state (only its downloads are left to overlap)
            main0 = torch.cuda.current_stream(dev)
            n0, cur0 = self.numOwn, self.cur
            mark("rebuild: begin", main0)
            self._up.wait_stream(main0)
            self._up.wait_stream(self._down)
            self._up.wait_stream(self._downE)
            with torch.cuda.stream(self._up):
                self.pos[cur0][:n0].copy_(hpos[:n0], non_blocking=True)
                self.vel[cur0][:n0].copy_(hvel[:n0], non_blocking=True)
            main0.wait_stream(self._up)
            mark("rebuild: uploaded", main0)
            self.state_modified()
            self.build_neibs()
            mark("rebuild: done", main0)
            self._down_key = None
            resident = True
            if self.numOwn == 0:
                return self._step_host_plain(hpos, hvel, chunks)
        be = self.backend
        ctx = be.fw.ctx
        main, es = torch.cuda.current_stream(dev), self._edge_stream
        n, n_own = self.numParticles, self.numOwn
        e0 = min(self.edge_start, n_own)
        cur, oth = self.cur, 1 - self.cur
        P = self.packed
        inner = self._inner_stripes()
        K = len(inner)
        ranges = {k: ab for k, ab in enumerate(inner)}
        if n_own > e0:
            ranges["E"] = (e0, n_own)
        lane_up = lambda name: self._upE if name == "E" else self._up
        lane_down = lambda name: self._downE if name == "E" else self._down
        key = ("pipe", self.iterations // self.buildneibsfreq, n_own, e0, K)
        chained = self._down_key == key
        mark("step: begin", main)
        # ---- uploads: the inner ranges in order on one lane, the edge stripe on its own. Chained to the previous step
        # piece by piece: the upload of a piece waits only for the download of the same piece (which followed its
        # corrector), so the copies of step n+1 run under the corrector of step n.
        up_ev = {}
        if not resident:
            for name, (a, b_) in ranges.items():
                lu = lane_up(name)
                if chained:
                    lu.wait_event(self._down_ev[name])
                else:
                    lu.wait_stream(main)
                    lu.wait_stream(self._down)
                    lu.wait_stream(self._downE)
                with torch.cuda.stream(lu):
                    self.pos[cur][a:b_].copy_(hpos[a:b_], non_blocking=True)
                    self.vel[cur][a:b_].copy_(hvel[a:b_], non_blocking=True)
                    up_ev[name] = torch.cuda.Event()
                    up_ev[name].record(lu)
                    mark(f"up {name}", lu)
        else:
            # earlier downloads read the buffers the corrector is about to integrate in place
            main.wait_stream(self._down)
            main.wait_stream(self._downE)
        have = set()

        def need(*names):                      # the compute stream sees these pieces and their neighbour records
            for name in names:
                if resident or name in have or name not in ranges:
                    continue                   # (a rebuild left state and records of every particle on the device)
                have.add(name)
                main.wait_event(up_ev[name])
                be.pack_state(self.pos[cur], self.vel[cur], P[cur], *ranges[name])
        # ---- predictor: the pair kernel of an inner range runs as soon as the range and its two neighbours are there;
        # the ranges next to the edge stripe (first and last) go last. The integration waits for dt.
        self._wait_halo(cur)                   # the halo update of state n (sent at the end of the previous step) has landed:
        #                                        nothing still reads the edge records this step re-makes from the upload
        args = (self.pos[cur], self.vel[cur], self.info, self.hash, self.cellstart, self.neibslist, self.forces_buf, self.cfl, n)
        off = 0
        order = list(range(1, K - 1)) + ([0] if K >= 1 else []) + ([K - 1] if K >= 2 else [])
        for k in order:
            need(k - 1, k, k + 1)
            if k == 0 or k == K - 1:
                need("E")
            off += be.forces(*args, *ranges[k], off, packed=P[cur])
            mark(f"pred {k}", main)
        if n_own > e0:
            need("E", 0, K - 1)
            es.wait_stream(main)
            try:
                ctx.use_stream(es)
                off += be.forces(*args, e0, n_own, off, packed=P[cur])
            finally:
                ctx.use_stream(main)
            main.wait_stream(es)
        self._cfl_candidate(off, 1)
        self._finish_dt()
        be.euler_async(self.pos[cur], self.vel[cur], self.info, self.hash, self.forces_buf, self.pos[oth], self.vel[oth],
                       n, n_own, 1, new_packed=P[oth])
        self._pending_x[oth] = self._start_halo_update(oth)
        predicted = torch.cuda.Event()
        predicted.record(main)
        mark("predicted", main)
        # ---- corrector: integration fused, IN PLACE into the state-n buffers, every range followed by its download
        args2 = (self.pos[oth], self.vel[oth], self.info, self.hash, self.cellstart, self.neibslist, self.forces_buf, self.cfl, n)
        fused = (self.pos[cur], self.vel[cur], self.pos[cur], self.vel[cur], 2, P[cur])
        down_ev = {}

        def download(name, after):
            a, b_ = ranges[name]
            ld = lane_down(name)
            ld.wait_event(after)
            with torch.cuda.stream(ld):
                hpos[a:b_].copy_(self.pos[cur][a:b_], non_blocking=True)
                hvel[a:b_].copy_(self.vel[cur][a:b_], non_blocking=True)
                down_ev[name] = torch.cuda.Event()
                down_ev[name].record(ld)
                mark(f"down {name}", ld)
        off = 0
        for k in range(K):
            off += be.forces(*args2, *ranges[k], off, packed=P[oth], fused=fused)
            done = torch.cuda.Event()
            done.record(main)
            mark(f"corr {k}", main)
            download(k, done)
        if n_own > e0:
            es.wait_event(predicted)
            with torch.cuda.stream(es):
                self._wait_halo(oth)
            try:
                ctx.use_stream(es)
                off += be.forces(*args2, e0, n_own, off, packed=P[oth], fused=fused)
            finally:
                ctx.use_stream(main)
            with torch.cuda.stream(es):
                self._pending_x[cur] = self._start_halo_update(cur)
                done = torch.cuda.Event()
                done.record(es)
                mark("corr E", es)
            download("E", done)
            main.wait_stream(es)
        self._cfl_candidate(off, 2)
        self.cfl_global.copy_(self.cfl_local)
        self._pending_dt = dist.all_reduce(self.cfl_global, op=dist.ReduceOp.MAX, group=self.group, async_op=True) or True
        self._stale = True
        mark("step: end", main)
        self.launches += 2 * (K + 1) + len(ranges) + 2
        self.iterations += 1
        self.total_interactions += 2 * int(self.last_neibs_info.num_interactions)
        self._down_ev, self._down_key = down_ev, key

    def _step_host_plain(self, hpos, hvel, chunks):
        """step_host around the ordinary step(): the copies in pieces, the download of step n next to the upload of
        step n+1 (the force evaluations wait for the whole upload)."""
        dev = self.device
        main = torch.cuda.current_stream(dev)
        n = self.numOwn
        cur = self.cur
        if n > 0:
            self._up.wait_stream(main)                       # whatever still reads / writes the state on the compute stream
            self._up.wait_stream(self._down)                 # and every earlier download (the piece tables differ)
            self._up.wait_stream(self._downE)
            with torch.cuda.stream(self._up):
                self.pos[cur][:n].copy_(hpos[:n], non_blocking=True)
                self.vel[cur][:n].copy_(hvel[:n], non_blocking=True)
            main.wait_stream(self._up)
            self.state_modified()
        self.step()
        n, cur = self.numOwn, self.cur
        b = [n * c // chunks for c in range(chunks + 1)]
        self._down.wait_stream(main)
        with torch.cuda.stream(self._down):
            for c in range(chunks):
                hpos[b[c]:b[c + 1]].copy_(self.pos[cur][b[c]:b[c + 1]], non_blocking=True)
                hvel[b[c]:b[c + 1]].copy_(self.vel[cur][b[c]:b[c + 1]], non_blocking=True)
        self._down_ev, self._down_key = {}, None

    def host_fence(self) -> None:
        """The compute stream waits for the copies of earlier step_host calls."""
        if getattr(self, "_down", None) is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._down)
            torch.cuda.current_stream(self.device).wait_stream(self._downE)

    def download_own(self) -> ParticleArrays:
        n = self.numOwn
        return ParticleArrays(self.pos[self.cur][:n].cpu().numpy(), self.vel[self.cur][:n].cpu().numpy(),
                              self.info[:n].cpu().numpy().view(np.uint16), self.hash[:n].cpu().numpy().view(np.uint32))

    def euler_once(self) -> None:
        cur, oth = self.cur, 1 - self.cur
        self.backend.euler(self.pos[cur], self.vel[cur], self.info, self.hash, self.forces_buf, self.pos[oth], self.vel[oth],
                           self.numParticles, self.numParticles, 0.0, 1)

    def forces_once(self) -> None:
        """One force evaluation on the current state without exchange (bench.py roofline timing)."""
        kw = {"packed": self.packed[self.cur]} if (self.packed is not None and self.device_dt) else {}
        self._wait_halo(self.cur)
        self.backend.forces(self.pos[self.cur], self.vel[self.cur], self.info, self.hash, self.cellstart, self.neibslist,
                            self.forces_buf, self.cfl, self.numParticles, 0, self.numOwn, **kw)
