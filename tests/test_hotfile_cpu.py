"""HotFile checkpoints (gpusph_b200/hotfile.py) against the reference's on-disk layout (src/writers/HotFile.h:44-58,
src/writers/HotFile.cc:46-58) and against the reference's own states (tests/golden, decoded from real HotFiles)."""
import glob
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from gpusph_b200 import hotfile as hfmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the reference's structs, restated (HotFile.h:44-58, HotFile.cc:46-51)
C_LAYOUT = r"""
#include <stdio.h>
#include <stddef.h>
typedef unsigned int uint; typedef unsigned long ulong;
typedef struct { uint version, buffer_count, particle_count, body_count, numOpenBoundaries; uint reserved[12];
                 ulong iterations; double t; float dt; uint _reserved[3]; } header_t;
typedef struct { uint name_length; char name[64]; uint element_size; uint array_count; } encoded_buffer_t;
int main() { printf("%zu %zu %zu %zu %zu %zu %zu", sizeof(header_t), offsetof(header_t, iterations), offsetof(header_t, t),
                    offsetof(header_t, dt), sizeof(encoded_buffer_t), offsetof(encoded_buffer_t, element_size),
                    offsetof(encoded_buffer_t, array_count)); return 0; }
"""


def test_struct_formats_match_the_reference_structs():
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "l.c"), "w").write(C_LAYOUT)
        subprocess.run(["gcc", "-o", os.path.join(d, "l"), os.path.join(d, "l.c")], check=True)
        v = [int(x) for x in subprocess.run([os.path.join(d, "l")], capture_output=True, text=True, check=True).stdout.split()]
    assert struct.calcsize(hfmod.HEADER) == v[0] == 104
    packed = struct.pack(hfmod.HEADER, 1, 5, 7, 0, 0, 0x1122334455667788, 2.5, 0.125)
    assert struct.unpack_from("@L", packed, v[1])[0] == 0x1122334455667788
    assert struct.unpack_from("@d", packed, v[2])[0] == 2.5
    assert struct.unpack_from("@f", packed, v[3])[0] == 0.125
    assert struct.calcsize(hfmod.BUFHDR) == v[4] == 76
    b = struct.pack(hfmod.BUFHDR, 8, b"Position", 16, 1)
    assert struct.unpack_from("@I", b, v[5])[0] == 16 and struct.unpack_from("@I", b, v[6])[0] == 1
    assert b[4:12] == b"Position" and b[12:68] == bytes(56)            # strcpy into a zeroed char[64]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))))
def test_write_then_read_reproduces_the_reference_state(path, tmp_path):
    g = np.load(path)
    it = int(g["iterations"][-1])
    pos, vel, info, hashv = g[f"pos_{it}"], g[f"vel_{it}"], g[f"info_{it}"], g[f"hash_{it}"]
    f = str(tmp_path / "hot_00021.bin")
    hfmod.write_hotfile(f, pos, vel, info, hashv, iterations=it, t=float(g[f"t_{it}"]), dt=float(g[f"dt_{it}"]))
    n = pos.shape[0]
    assert os.path.getsize(f) == 104 + 4 * 76 + n * (16 + 16 + 8 + 4)
    hf = hfmod.read_hotfile(f)
    assert (hf["buffer_count"], hf["particle_count"], hf["body_count"], hf["iterations"]) == (5, n, 0, it)
    assert hf["t"] == float(g[f"t_{it}"]) and np.float32(hf["dt"]) == g[f"dt_{it}"]
    assert list(hf["buffers"]) == ["Position", "Velocity", "Info", "Hash"]     # buffer-key order, define_buffers.h:49-58
    p2, v2, i2, h2 = hfmod.particle_arrays(hf)
    assert np.array_equal(p2.view(np.uint32), pos.view(np.uint32)) and np.array_equal(v2.view(np.uint32), vel.view(np.uint32))
    assert np.array_equal(i2, info) and np.array_equal(h2, hashv)


def test_reader_rejects_other_versions(tmp_path):
    f = str(tmp_path / "bad.bin")
    open(f, "wb").write(struct.pack(hfmod.HEADER, 2, 5, 0, 0, 0, 0, 0.0, 0.0))
    with pytest.raises(ValueError):
        hfmod.read_hotfile(f)


def test_writer_checks_shapes(tmp_path):
    with pytest.raises(ValueError):
        hfmod.write_hotfile(str(tmp_path / "x.bin"), np.zeros((3, 4), np.float32), np.zeros((2, 4), np.float32),
                            np.zeros((3, 4), np.uint16), np.zeros(3, np.uint32), iterations=1, t=0.1, dt=0.01)
