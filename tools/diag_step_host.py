"""GPU diagnostic: where does Worker.step_host spend its time? (phase timings with host clocks + synchronize)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_problem
from gpusph_b200.simulation import Worker

wl = sys.argv[1] if len(sys.argv) > 1 else "dambreak2m"
params, parts = make_problem(wl)
w = Worker(params, parts, 0)
for _ in range(12):
    w.step()
A, n = w.pos[0].shape[0], w.numParticles
hp, hv = torch.empty((A, 4)).pin_memory(), torch.empty((A, 4)).pin_memory()
hp[:n].copy_(w.pos[w.cur][:n]); hv[:n].copy_(w.vel[w.cur][:n])
torch.cuda.synchronize()
print("pinned", hp.is_pinned(), hp[100:200].is_pinned(), "n", n, "stripes", len(w._stripes()), w._stripes()[:3])


def T(fn, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


pos = w.pos[w.cur]
print("H2D whole   ms", T(lambda: pos[:n].copy_(hp[:n], non_blocking=True)))
print("D2H whole   ms", T(lambda: hp[:n].copy_(pos[:n], non_blocking=True)))
S = w._stripes()
def sliced_h2d():
    for a, b in S:
        pos[a:b].copy_(hp[a:b], non_blocking=True)
def sliced_d2h():
    for a, b in S:
        hp[a:b].copy_(pos[a:b], non_blocking=True)
print("H2D sliced  ms", T(sliced_h2d))
print("D2H sliced  ms", T(sliced_d2h))
side = torch.cuda.Stream()
def side_h2d():
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for a, b in S:
            pos[a:b].copy_(hp[a:b], non_blocking=True)
    torch.cuda.current_stream().wait_stream(side)
print("H2D side    ms", T(side_h2d))
def side_d2h():
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for a, b in S:
            hp[a:b].copy_(pos[a:b], non_blocking=True)
    torch.cuda.current_stream().wait_stream(side)
print("D2H side    ms", T(side_d2h))
rd = w.state(w.cur)
def striped_forces():
    off = 0
    for a, b in S:
        off += w.forces.basicstep(rd, rd, n, a, b, off, step=1, dt_from_device=True)
print("forces one  ms", T(w.forces_once))
print("forces strp ms", T(striped_forces))
# whole steps (avoid rebuild iterations)
def steps_host():
    if w.iterations % 10 == 0:
        w.step()
    w.step_host(hp, hv)
print("step_host   ms", T(steps_host, 8))
def steps_res():
    if w.iterations % 10 == 0:
        w.step()
    w.step()
print("step        ms", T(steps_res, 8))
