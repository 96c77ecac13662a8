#!/usr/bin/env python
"""Time line of the resident multi-GPU step (SlabWorker.step): run under torchrun like bench.py,
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/diag_slab_step.py [workload]
One whole neighbour-rebuild period (rebuild + buildneibsfreq steps) is traced with CUDA events on the compute stream and
the host clock at enqueue time. Per rank: GPU ms per phase (summed over the period), host ms spent enqueueing, and how
far the host ran ahead of the GPU at the end of the period; then the phase-by-phase time line of rank 0."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from gpusph_b200.multigpu import SlabWorker
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    params, parts = bench.make_problem(sys.argv[1] if len(sys.argv) > 1 else "dambreak16m", world, "strong")
    w = SlabWorker(params, parts, local, rank=rank, world=world)
    freq = w.buildneibsfreq
    for _ in range(2 * freq):
        w.step()
    torch.cuda.synchronize(); dist.barrier()
    # plain timing of two periods
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2 * freq):
        w.step()
    b.record()
    torch.cuda.synchronize(); dist.barrier()
    ms = a.elapsed_time(b) / (2 * freq)
    # traced period
    w._trace = []
    h0 = time.perf_counter()
    for _ in range(freq):
        w.step()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    h2 = time.perf_counter()
    tr, w._trace = w._trace, None
    t0 = tr[0][2]
    rows = [(t0.elapsed_time(ev), (th - h0) * 1e3, it, label) for it, label, ev, th in tr]
    phase = {}
    prev = 0.0
    for tg, _, _, label in rows:
        phase[label] = phase.get(label, 0.0) + tg - prev
        prev = tg
    summary = (f"rank {rank}: own {w.numOwn} halo {w.numParticles - w.numOwn} edge {w.numOwn - w.edge_start}; {ms:.3f} ms/step untraced; "
               f"traced period: gpu {rows[-1][0]:.2f} ms, host enqueue {(h1 - h0) * 1e3:.2f} ms, host waited {(h2 - h1) * 1e3:.2f} ms at the end\n    "
               + "  ".join(f"{k}={v:.2f}" for k, v in phase.items() if k != "step: begin" or v > 0.005))
    for r in range(world):
        dist.barrier()
        if rank == r:
            print(summary, flush=True)
    dist.barrier()
    if rank == 0:
        print("--- rank 0 time line: gpu ms | host ms at enqueue | step | phase", flush=True)
        for tg, th, it, label in rows:
            print(f"  {tg:8.3f}  {th:8.3f}  {it}  {label}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
