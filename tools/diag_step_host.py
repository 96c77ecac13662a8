"""GPU diagnostic: (1) PCIe copy bandwidth of this box, one direction at a time and both at once; (2) the time line of
chained b200sph_step_host calls (B200SPH_HOST_TRACE=1)."""
import os, sys, time
os.environ["B200SPH_HOST_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_problem
from gpusph_b200.simulation import Worker

wl = sys.argv[1] if len(sys.argv) > 1 else "dambreak2m"
MB = 64
h1, h2 = torch.empty(MB << 20, dtype=torch.uint8).pin_memory(), torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d1, d2 = torch.empty(MB << 20, dtype=torch.uint8, device="cuda"), torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def T(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def up():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both():
    up(); down()
def chunks(k):
    c = (MB << 20) // k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1): d1[i * c:(i + 1) * c].copy_(h1[i * c:(i + 1) * c], non_blocking=True)
            with torch.cuda.stream(s2): h2[i * c:(i + 1) * c].copy_(d2[i * c:(i + 1) * c], non_blocking=True)
    return f
print("H2D alone GB/s", MB / 1024 / T(up) * 1.048576)
print("D2H alone GB/s", MB / 1024 / T(down) * 1.048576)
print("both at once, GB/s per direction", MB / 1024 / T(both) * 1.048576)
print("both at once in 16 chunks, GB/s per direction", MB / 1024 / T(chunks(16)) * 1.048576)

params, parts = make_problem(wl)
w = Worker(params, parts, 0)
for _ in range(11):
    w.step()
A, n = w.pos[0].shape[0], w.numParticles
hp, hv = torch.empty((A, 4)).pin_memory(), torch.empty((A, 4)).pin_memory()
hp[:n].copy_(w.pos[w.cur][:n]); hv[:n].copy_(w.vel[w.cur][:n])
torch.cuda.synchronize()
for k in range(6):
    w.step_host(hp, hv)
w.host_sync()                       # prints the time line of the 6th chained call
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(3):
    w.step_host(hp, hv)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("3 chained non-rebuild steps: ms/step", (time.perf_counter() - t0) / 3 * 1e3, "host enqueue ms/step", (t1 - t0) / 3 * 1e3, "iteration now", w.iterations)
w.step_host(hp, hv)                 # iteration 20: rebuild step
w.host_sync()
w.step_host(hp, hv)
w.host_sync()
