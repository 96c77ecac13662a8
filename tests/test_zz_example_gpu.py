"""examples/lattice_steps.cu (plain C++ on the C ABI) against the Python worker on the same lattice: same call
sequence, same library. The two set-ups compute the initial velocities with different sin/cos implementations (libm vs
numpy: a last-bit difference in a few float32 inputs), so the comparison is to 1e-6 on the sums, not bitwise."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_cpp_example_reproduces_the_python_worker():
    import __graft_entry__ as g
    from gpusph_b200.problems import lattice_problem
    from gpusph_b200.simulation import Worker
    exe = g.build_example()
    n, steps = 20, 12
    out = subprocess.run([exe, str(n), str(steps)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-1500:]
    got = json.loads(out.stdout.strip().splitlines()[-1])
    params, parts = lattice_problem(n, jitter=0.0)
    w = Worker(params, parts, 0)
    for _ in range(steps):
        w.step()
    st = w.download()
    assert got["particles"] == st.pos.shape[0] and got["iterations"] == steps
    assert got["dt"] == pytest.approx(w.dt, rel=1e-5) and got["t"] == pytest.approx(w.t, rel=1e-6)
    sv = np.abs(st.vel[:, :3].astype(np.float64)).sum(axis=1)
    assert got["sum_abs_vel"] == pytest.approx(float(np.add.reduce(sv)), rel=1e-6)
    assert got["sum_rho_tilde"] == pytest.approx(float(st.vel[:, 3].astype(np.float64).sum()), rel=1e-4, abs=1e-6 * st.pos.shape[0])
    assert got["neibs_per_particle"] == pytest.approx(w.last_neibs_info.num_interactions / st.pos.shape[0], abs=1e-3)
