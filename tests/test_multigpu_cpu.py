"""Slab decomposition + halo exchange host logic (gpusph_b200/multigpu.py) at world_size 2 on gloo, with the
oracle as compute backend. The decomposed run must reproduce the single-domain run BITWISE (per-particle
summation order does not depend on the decomposition) including particles that change owner."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from gpusph_b200 import capi  # noqa: E402
from gpusph_b200.multigpu import compact_device_map, slab_partition  # noqa: E402
from gpusph_b200.problems import dambreak_problem, lattice_problem  # noqa: E402


def make_problem(periodic=False):
    if periodic == "xzy":
        # the linearisation bench.py uses on N > 1 GPUs (and the reference's `linearization=xzy` build): y is the slowest
        # hash digit, hence the slab axis
        params, parts = lattice_problem(8, ny=14, nz=8, jitter=0.2, densitydiffusion=capi.RHODIFF_COLAGROSSI, coord=(0, 2, 1))
        parts.vel[:, 1] += 6.0
        return params, parts
    if periodic:
        # periodic along x, the slab axis of the default yzx linearisation: the first and the last slab are neighbours.
        # The domain must be a whole number of lattice spacings long for the lattice to close on itself.
        params, parts = lattice_problem(16, ny=8, nz=8, jitter=0.2, densitydiffusion=capi.RHODIFF_COLAGROSSI, periodic=capi.PERIODIC_X)
    else:
        params, parts = lattice_problem(14, ny=8, nz=8, jitter=0.2, densitydiffusion=capi.RHODIFF_COLAGROSSI)
    # a strong flow along the split axis so that particles cross the slab face within a few steps
    parts.vel[:, 0] += 6.0
    return params, parts


def _rank_main(rank, world, port, steps, outdir, device_dt=False, periodic=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from gpusph_b200.multigpu import SlabWorker
    from oracle_backend import OracleBackend, OracleDeviceDtBackend
    params, parts = make_problem(periodic)
    w = SlabWorker(params, parts, None, rank=rank, world=world, backend=OracleDeviceDtBackend if device_dt else OracleBackend)
    assert w.device_dt == device_dt
    own0 = None
    dts = []
    for _ in range(steps):
        w.step()
        dts.append(w.dt)
        if own0 is None:
            own0 = w.numOwn
    out = w.download_own()
    # the piece table of the host-resident step (SlabWorker.step_host): pieces of ~40 particles for this small system
    os.environ["B200SPH_SLAB_HOST_PIECE"] = "40"
    pieces = np.array(w._inner_stripes(), dtype=np.int64).reshape(-1, 2)
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), pos=out.pos, vel=out.vel, info=out.info, hash=out.hash,
             dts=np.array(dts), own0=own0, slab=np.array(w.slab), inter=w.total_interactions,
             pieces=pieces, edge_start=w.edge_start, S=w.S)
    dist.destroy_process_group()


def ids_of(info):
    return (info[:, 3].astype(np.int64) << 16) | info[:, 2]


def test_partition_and_device_map():
    params, parts = dambreak_problem(0.05)
    slabs = slab_partition(params, parts.hash, 3)
    G3 = int(params.grid_size[params.coord[2]])
    assert slabs[0][0] == 0 and slabs[-1][1] == G3
    assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:])) and all(e - s >= 2 for s, e in slabs)
    S = int(params.grid_size[params.coord[0]]) * int(params.grid_size[params.coord[1]])
    cdm = compact_device_map(params, slabs[1], 1, 3)
    t = (cdm >> 30).reshape(G3, S)
    xs, xe = slabs[1]
    assert (t[xs] == 1).all() and (t[xe - 1] == 1).all() and (t[xs - 1] == 2).all() and (t[xe] == 2).all()
    assert (t[xs + 1:xe - 1] == 0).all() and (t[:xs - 1] == 3).all() and (t[xe + 1:] == 3).all()
    with pytest.raises(ValueError):
        slab_partition(params, parts.hash, G3)        # fewer than 2 layers per device
    # periodic along the slab axis: the layer "below" the first slab is the last layer of the grid
    pp, _ = lattice_problem(6, periodic=1 << params.coord[2])
    Gp = int(pp.grid_size[pp.coord[2]])
    Sp = int(pp.grid_size[pp.coord[0]]) * int(pp.grid_size[pp.coord[1]])
    tp = (compact_device_map(pp, (0, Gp // 2), 0, 2) >> 30).reshape(Gp, Sp)
    assert (tp[0] == 1).all() and (tp[Gp // 2 - 1] == 1).all() and (tp[Gp - 1] == 2).all() and (tp[Gp // 2] == 2).all()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_periodic_slab_axis_matches_single_domain_bitwise(world):
    """Periodicity ALONG the slab axis (the reference's device map handles it, src/GPUWorker.cc:1560-1634): the first and
    the last slab exchange halos across the periodic face; with two slabs both neighbours of a rank are the same rank."""
    import oracle_binding as ob
    steps = 12
    params, parts = make_problem(periodic=True)
    ref = ob.OracleWorker(params, parts)
    for _ in range(steps):
        ref.step()
    exp = ref.download()
    with tempfile.TemporaryDirectory() as d:
        port = 31500 + (os.getpid() % 2000) + world
        mp.spawn(_rank_main, args=(world, port, steps, d, True, True), nprocs=world, join=True)
        r = [np.load(os.path.join(d, f"rank{k}.npz")) for k in range(world)]
    ids = np.concatenate([ids_of(r[k]["info"]) for k in range(world)])
    assert np.array_equal(np.sort(ids), np.arange(parts.n))
    assert any(int(r[k]["own0"]) != r[k]["pos"].shape[0] for k in range(world)), "particles should change owner"
    pos = np.concatenate([r[k]["pos"] for k in range(world)])
    vel = np.concatenate([r[k]["vel"] for k in range(world)])
    hashv = np.concatenate([r[k]["hash"] for k in range(world)]) & 0x3FFFFFFF
    o, oe = np.argsort(ids), np.argsort(ids_of(exp.info))
    assert np.array_equal(hashv[o], exp.hash[oe])
    assert np.array_equal(pos[o].view(np.uint32), exp.pos[oe].view(np.uint32))
    assert np.array_equal(vel[o].view(np.uint32), exp.vel[oe].view(np.uint32))


@pytest.mark.timeout(600)
def test_xzy_linearisation_slabs_along_y_match_single_domain_bitwise():
    """bench.py --gpus N and the reference's own multi-GPU DamBreak3D split along Y; with the xzy cell linearisation y is
    the slowest hash digit. Same bitwise comparison as above on that linearisation, 3 ranks."""
    import oracle_binding as ob
    steps, world = 12, 3
    params, parts = make_problem("xzy")
    assert list(params.coord) == [0, 2, 1]
    ref = ob.OracleWorker(params, parts)
    for _ in range(steps):
        ref.step()
    exp = ref.download()
    with tempfile.TemporaryDirectory() as d:
        port = 27500 + (os.getpid() % 2000)
        mp.spawn(_rank_main, args=(world, port, steps, d, True, "xzy"), nprocs=world, join=True)
        r = [np.load(os.path.join(d, f"rank{k}.npz")) for k in range(world)]
    ids = np.concatenate([ids_of(r[k]["info"]) for k in range(world)])
    assert np.array_equal(np.sort(ids), np.arange(parts.n))
    assert any(int(r[k]["own0"]) != r[k]["pos"].shape[0] for k in range(world)), "particles should change owner"
    pos = np.concatenate([r[k]["pos"] for k in range(world)])
    vel = np.concatenate([r[k]["vel"] for k in range(world)])
    hashv = np.concatenate([r[k]["hash"] for k in range(world)]) & 0x3FFFFFFF
    o, oe = np.argsort(ids), np.argsort(ids_of(exp.info))
    assert np.array_equal(hashv[o], exp.hash[oe])
    assert np.array_equal(pos[o].view(np.uint32), exp.pos[oe].view(np.uint32))
    assert np.array_equal(vel[o].view(np.uint32), exp.vel[oe].view(np.uint32))


@pytest.mark.timeout(600)
@pytest.mark.parametrize("device_dt", [False, True], ids=["host_dt", "deferred_device_dt"])
def test_two_rank_gloo_run_matches_single_domain_bitwise(device_dt):
    """device_dt: the production control flow on GPUs — dt record owned by the backend, CFL maxima of both force
    evaluations all-reduced once per step, asynchronously, and consumed just before the next step's first euler."""
    import oracle_binding as ob
    steps = 12
    params, parts = make_problem()
    ref = ob.OracleWorker(params, parts)
    ref_dts = []
    for _ in range(steps):
        ref.step()
        ref_dts.append(ref.dt)
    exp = ref.download()
    with tempfile.TemporaryDirectory() as d:
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_rank_main, args=(2, port + int(device_dt), steps, d, device_dt), nprocs=2, join=True)
        r = [np.load(os.path.join(d, f"rank{k}.npz")) for k in range(2)]
    # same adaptive time-step sequence on both ranks and in the single-domain run
    assert np.array_equal(r[0]["dts"], r[1]["dts"])
    assert np.array_equal(r[0]["dts"].astype(np.float32), np.array(ref_dts, dtype=np.float32))
    # ownership is a partition of the particles, and some particles changed owner
    ids = np.concatenate([ids_of(r[k]["info"]) for k in range(2)])
    assert np.array_equal(np.sort(ids), np.arange(parts.n))
    assert int(r[0]["own0"]) != r[0]["pos"].shape[0], "test problem should move particles across the slab face"
    # interactions counted once
    assert int(r[0]["inter"]) + int(r[1]["inter"]) > 0
    # the pieces of the host-resident step: they tile the inner stripe [0, edge_start), and every boundary is the first
    # particle of a cell layer (so that the neighbours of a piece lie in the adjacent pieces, the edge stripe or the halo)
    for k in range(2):
        pc, e0 = r[k]["pieces"], int(r[k]["edge_start"])
        e0 = min(e0, r[k]["pos"].shape[0])
        assert pc.shape[0] >= 2 and pc[0, 0] == 0 and pc[-1, 1] == e0
        assert np.array_equal(pc[1:, 0], pc[:-1, 1]) and (pc[:, 1] > pc[:, 0]).all()
        layer = (r[k]["hash"].astype(np.int64) & 0x3FFFFFFF) // int(r[k]["S"])
        for b in pc[1:, 0]:
            assert layer[b] != layer[b - 1], "a piece must start with a cell layer"
    pos = np.concatenate([r[k]["pos"] for k in range(2)])
    vel = np.concatenate([r[k]["vel"] for k in range(2)])
    hashv = np.concatenate([r[k]["hash"] for k in range(2)]) & 0x3FFFFFFF
    o, oe = np.argsort(ids), np.argsort(ids_of(exp.info))
    assert np.array_equal(hashv[o], exp.hash[oe])
    assert np.array_equal(pos[o].view(np.uint32), exp.pos[oe].view(np.uint32))
    assert np.array_equal(vel[o].view(np.uint32), exp.vel[oe].view(np.uint32))


def test_state_checksum_is_additive_order_independent_and_sensitive():
    """multigpu.state_checksum (what bench.py all-reduces after the warm-up of a multi-GPU run to compare with a
    single-GPU run of the same steps)."""
    import torch
    from gpusph_b200.multigpu import state_checksum
    params, parts = make_problem()
    n = parts.n
    info = torch.from_numpy(parts.info.view(np.int16).copy())
    hashv = torch.from_numpy(parts.hash.view(np.int32).copy())
    pos, vel = torch.from_numpy(parts.pos.copy()), torch.from_numpy(parts.vel.copy())
    whole, cnt = state_checksum(info, hashv, pos, vel)
    assert cnt == n and 0 < whole < (1 << 57)
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(1))
    assert state_checksum(info[perm], hashv[perm], pos[perm], vel[perm])[0] == whole
    cut = n // 3
    a = state_checksum(info[:cut], hashv[:cut], pos[:cut], vel[:cut])
    b = state_checksum(info[cut:], hashv[cut:], pos[cut:], vel[cut:])
    assert a[0] + b[0] == whole and a[1] + b[1] == n
    # the multi-GPU cell-type bits in the hash do not count, everything else does, down to one bit of one float
    tagged = hashv | torch.tensor(-(1 << 31), dtype=torch.int32)
    assert state_checksum(info, tagged, pos, vel)[0] == whole
    v2 = vel.clone()
    v2.view(torch.int32)[n // 2, 1] ^= 1
    assert state_checksum(info, hashv, pos, v2)[0] != whole
    h2 = hashv.clone(); h2[5] += 1
    assert state_checksum(info, h2, pos, vel)[0] != whole
    assert state_checksum(info[:0], hashv[:0], pos[:0], vel[:0]) == (0, 0)
