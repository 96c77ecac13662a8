#!/usr/bin/env python
"""Time line of SlabWorker.step_host (N > 1 GPUs, state in pinned host memory): run under torchrun like bench.py,
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/diag_slab_step_host.py [workload]
Prints, for rank 0 and the last rank, when every piece's upload / predictor / corrector / download finished relative to
the beginning of the first traced step (CUDA events on the streams the work runs on), and the mean ms per step."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from gpusph_b200.hostmem import bind_host_near_gpu
    from gpusph_b200.multigpu import SlabWorker
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    aff = bind_host_near_gpu(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    params, parts = bench.make_problem(sys.argv[1] if len(sys.argv) > 1 else "dambreak16m", world, "strong")
    w = SlabWorker(params, parts, local, rank=rank, world=world)
    for _ in range(11):
        w.step()
    A = w.pos[0].shape[0]
    hp, hv = torch.empty((A, 4)).pin_memory(), torch.empty((A, 4)).pin_memory()
    n = w.numOwn
    hp[:n].copy_(w.pos[w.cur][:n]); hv[:n].copy_(w.vel[w.cur][:n])
    torch.cuda.synchronize(); dist.barrier()
    for _ in range(4):
        w.step_host(hp, hv)
    torch.cuda.synchronize(); dist.barrier()
    # plain timing over the rest of this rebuild period and the next one
    steps = (-w.iterations) % w.buildneibsfreq + w.buildneibsfreq
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        w.step_host(hp, hv)
    w.host_fence(); b.record()
    torch.cuda.synchronize(); dist.barrier()
    ms = a.elapsed_time(b) / steps
    # traced: 3 pipelined steps in the middle of a period
    for _ in range(3):
        w.step_host(hp, hv)
    torch.cuda.synchronize(); dist.barrier()
    w._trace = []
    for _ in range(3):
        w.step_host(hp, hv)
    w.host_fence()
    torch.cuda.synchronize()
    tr, w._trace = w._trace, None
    t0 = tr[0][2]
    for r in (0, world - 1):
        dist.barrier()
        if rank == r:
            print(f"--- rank {rank}: own {w.numOwn} edge_start {w.edge_start} pieces {w._inner_stripes()} affinity {aff}; "
                  f"{ms:.3f} ms/step over {steps} steps (1 rebuild)", flush=True)
            for it, label, ev, _ in sorted(tr, key=lambda x: t0.elapsed_time(x[2])):
                print(f"  {t0.elapsed_time(ev):8.3f} ms  step {it}  {label}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
