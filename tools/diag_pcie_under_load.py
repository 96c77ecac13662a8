"""GPU diagnostic: PCIe copy bandwidth while the forces kernel is running on another stream."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_problem
from gpusph_b200.simulation import Worker

params, parts = make_problem("dambreak2m")
w = Worker(params, parts, 0)
for _ in range(3):
    w.step()
MB = 64
h1, h2 = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(), torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d1, d2 = torch.empty(MB << 20, dtype=torch.uint8, device="cuda"), torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")


def measure(copy_fn, load_fn, reps=6):
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        for _ in range(4):
            load_fn()
        a, b = copy_fn()
        evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[len(evs) // 2]
    return MB / 1024 * 1.048576 / (ms / 1e3)


def up():
    with torch.cuda.stream(s1):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); d1.copy_(h1, non_blocking=True); b.record()
    return a, b
def down():
    with torch.cuda.stream(s2):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); h2.copy_(d2, non_blocking=True); b.record()
    return a, b
def both_up():
    r = up(); down(); return r
def both_down():
    up(); return down()
loads = {"idle": lambda: None, "forces kernel": w.forces_once, "euler kernel (HBM streaming)": w.euler_once,
         "memset 1 GB (HBM write)": lambda: big.fill_(1)}
for name, load in loads.items():
    print(f"{name:32s} H2D alone {measure(up, load):5.1f}  D2H alone {measure(down, load):5.1f}  "
          f"H2D with D2H {measure(both_up, load):5.1f}  D2H with H2D {measure(both_down, load):5.1f}  GB/s")
