"""gpusph_b200 — Blackwell-native WCSPH per-timestep engine behind GPUSPH's engine API.

Layout: csrc/ (hand-written sm_100a CUDA kernels + the C ABI of include/b200sph.h),
capi.py (ctypes binding), engines.py (host-side mirror of the reference's abstract engines),
simulation.py (per-GPU worker driving one predictor-corrector step), problems.py (host setup),
multigpu.py (1-D slab decomposition + halo exchange over torch.distributed).
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
