"""2-GPU slab run (NCCL halo exchange, CUDA engines) vs the single-GPU run: bitwise identical."""
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

pytestmark = pytest.mark.gpu


def make_problem(periodic=False):
    from gpusph_b200 import capi
    from gpusph_b200.problems import dambreak_problem, lattice_problem
    if periodic:
        # periodic along x, the slab axis: the first and the last slab are neighbours across the periodic face
        params, parts = lattice_problem(32, ny=12, nz=12, jitter=0.2, densitydiffusion=capi.RHODIFF_COLAGROSSI,
                                        periodic=capi.PERIODIC_X)
        parts.vel[:, 0] += 6.0
        return params, parts
    params, parts = dambreak_problem(0.02, densitydiffusion=capi.RHODIFF_FERRARI, density_diff_coeff=0.1)
    parts.vel[:, 0] += 3.0 * ((parts.info[:, 0] & 7) == 0)        # push the column across the slab faces
    return params, parts


def _rank_main(rank, world, port, steps, outdir, host_state=False, periodic=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from gpusph_b200.multigpu import SlabWorker
    params, parts = make_problem(periodic)
    w = SlabWorker(params, parts, rank, rank=rank, world=world)
    dts = []
    if host_state:
        # the state of the owned particles lives in pinned host buffers: every step uploads it and downloads the result
        # (SlabWorker.step_host, what bench.py's e2e leg times on N > 1 GPUs); no host synchronisation in between
        A = w.pos[0].shape[0]
        hp, hv = torch.zeros((A, 4)).pin_memory(), torch.zeros((A, 4)).pin_memory()
        w.step()                                   # the first rebuild decides who owns what
        dts.append(w.dt)
        n = w.numOwn
        hp[:n].copy_(w.pos[w.cur][:n]); hv[:n].copy_(w.vel[w.cur][:n])
        torch.cuda.synchronize()
        for _ in range(steps - 1):
            w.step_host(hp, hv)
        w.host_fence()
        torch.cuda.synchronize()
        n = w.numOwn
        assert torch.equal(hp[:n], w.pos[w.cur][:n].cpu()) and torch.equal(hv[:n], w.vel[w.cur][:n].cpu())
        dts = [w.dt] * steps
    else:
        for _ in range(steps):
            w.step()
            dts.append(w.dt)
    out = w.download_own()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), pos=out.pos, vel=out.vel, info=out.info, hash=out.hash,
             dts=np.array(dts))
    dist.destroy_process_group()


def ids_of(info):
    return (info[:, 3].astype(np.int64) << 16) | info[:, 2]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 4])
def test_periodic_slab_axis_matches_single_gpu_bitwise(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _compare_with_single_gpu(world, False, True)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [4, 8])
def test_many_gpu_slab_run_matches_single_gpu_bitwise(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _compare_with_single_gpu(world, False, False)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("host_state", [False, True], ids=["resident", "host_state"])
def test_two_gpu_slab_run_matches_single_gpu_bitwise(host_state):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _compare_with_single_gpu(2, host_state, False)


def _compare_with_single_gpu(world, host_state, periodic):
    from gpusph_b200.simulation import Worker
    steps = 12
    params, parts = make_problem(periodic)
    ref = Worker(params, parts, 0)
    ref_dts = []
    for _ in range(steps):
        ref.step()
        ref_dts.append(ref.dt)
    exp = ref.download()
    with tempfile.TemporaryDirectory() as d:
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_rank_main, args=(world, port, steps, d, host_state, periodic), nprocs=world, join=True)
        r = [np.load(os.path.join(d, f"rank{k}.npz")) for k in range(world)]
    assert all(np.array_equal(r[0]["dts"], r[k]["dts"]) for k in range(1, world))
    if host_state:
        assert np.float32(r[0]["dts"][-1]) == np.float32(ref_dts[-1])
    else:
        assert np.array_equal(r[0]["dts"].astype(np.float32), np.array(ref_dts, dtype=np.float32))
    ids = np.concatenate([ids_of(r[k]["info"]) for k in range(world)])
    assert np.array_equal(np.sort(ids), np.arange(parts.n))
    pos = np.concatenate([r[k]["pos"] for k in range(world)])
    vel = np.concatenate([r[k]["vel"] for k in range(world)])
    hashv = np.concatenate([r[k]["hash"] for k in range(world)]) & 0x3FFFFFFF
    o, oe = np.argsort(ids), np.argsort(ids_of(exp.info))
    assert np.array_equal(hashv[o], exp.hash[oe])
    assert np.array_equal(pos[o].view(np.uint32), exp.pos[oe].view(np.uint32))
    assert np.array_equal(vel[o].view(np.uint32), exp.vel[oe].view(np.uint32))
