"""bench.py --impl reference: the arithmetic that turns the reference's own per-command timers into a step time, run
against a stub that prints what the real binary prints (src/GPUSPH.cc:118-131 showCommandTimes, :1941 printStatus)."""
import json
import os
import stat
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STUB = """#!/bin/sh
echo "Simulation time t=1.0e-02s, iteration=30, dt=1.0e-04s, 2,011,782 parts (4.9e+02, cum. 5e+02 MIPPS), maxneibs 93+0"
echo "Elapsed time of simulation cycle: 0.48s"
printf 'CMDTIMES:COMMAND\\tCMD_NUM\\tCALLS\\tMAX(ms)\\tTOT(ms)\\n'
printf 'CMDTIMES:FORCES_SYNC\\t43\\t60\\t5.0\\t123.0\\n'
printf 'CMDTIMES:EULER\\t50\\t60\\t1.0\\t7.0\\n'
printf 'CMDTIMES:SORT\\t12\\t3\\t700.0\\t704.0\\n'
printf 'CMDTIMES:BUILDNEIBS\\t14\\t3\\t2.5\\t6.5\\n'
printf 'CMDTIMES:INIT_ONCE\\t2\\t1\\t50.0\\t50.0\\n'
"""


def test_reference_arm_drops_the_slowest_call_of_every_command(tmp_path):
    stub = tmp_path / "DamBreak3D"
    stub.write_text(STUB)
    stub.chmod(stub.stat().st_mode | stat.S_IEXEC)
    env = dict(os.environ, B200SPH_REF_BIN=str(stub), B200SPH_NEIBS_PER_PARTICLE="50")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20", "--warmup", "10"],
                         capture_output=True, text=True, env=env, check=True).stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 20
    # per command: (total - max) / (calls - 1) * calls / iterations; one-off commands (1 call) are not per-step work
    exp = ((123.0 - 5.0) / 59 * 60 + (7.0 - 1.0) / 59 * 60 + (704.0 - 700.0) / 2 * 3 + (6.5 - 2.5) / 2 * 3) / 30
    assert abs(d["ms_per_step"] - exp) < 1e-9
    assert d["particles"] == 2011782
    ups = 2011782 / (exp / 1e3)
    assert abs(d["particle_updates_per_s"] - ups) / ups < 1e-12
    assert abs(d["value"] - ups * 50 * 2 / 1e6) / d["value"] < 1e-12
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference"
