"""Host-side placement for the host-resident stepping paths (Worker.step_host / SlabWorker.step_host).

The state crosses PCIe both ways every step there, so where the pinned buffers live matters: on a two-socket node a
buffer allocated on the socket that does NOT host the GPU's PCIe root sends every byte over the inter-socket link, and
with one process per GPU nothing pins the processes anywhere by default. `bind_host_near_gpu()` restricts the calling
thread to the CPUs NVML reports as local to the GPU; memory the thread touches (and pins) afterwards is allocated on
that node by the kernel's default first-touch policy. The reference leaves this to the user (numactl / MPI binding)."""
from __future__ import annotations

import os


def bind_host_near_gpu(device_index: int) -> str | None:
    """Bind the calling thread to the CPUs local to CUDA device `device_index`. Returns a short description of what
    was done, or None when nothing was (B200SPH_BIND_NUMA=0, NVML unavailable, a single-node machine's full mask...).
    Never raises: placement is an optimisation."""
    if os.environ.get("B200SPH_BIND_NUMA", "1") == "0":
        return None
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device_index)
        bus = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus or len(cpus) == len(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} cpus local to {bus}"
    except Exception:
        return None
