"""Decoder for GPUSPH HotFile checkpoints (layout: src/writers/HotFile.h:44-58, HotFile.cc:184-250;
the reference's own decoder is scripts/hotinfo.py:12-16). Test infrastructure."""
from __future__ import annotations

import struct

import numpy as np

HEADER = "@IIIII48xLdf12x"     # version, buffer_count, particle_count, body_count, numOpenBoundaries, iterations, t, dt
BUFHDR = "@I64sII"             # name length, name, element size, array count


def read_hotfile(path: str) -> dict:
    out = {}
    with open(path, "rb") as f:
        h = struct.unpack(HEADER, f.read(struct.calcsize(HEADER)))
        version, nbuf, nparts, nbodies, nopen, iterations, t, dt = h
        assert version == 1
        out.update(particle_count=nparts, body_count=nbodies, iterations=iterations, t=t, dt=dt, buffers={})
        for _ in range(nbuf):
            raw = f.read(struct.calcsize(BUFHDR))
            if len(raw) < struct.calcsize(BUFHDR):
                break                       # ephemeral buffers are counted but not stored
            ln, name, elsize, _count = struct.unpack(BUFHDR, raw)
            name = name[:ln].decode()
            data = f.read(elsize * nparts)
            if len(data) < elsize * nparts:
                break
            out["buffers"][name] = (elsize, data)
    return out


def particle_arrays(hf: dict):
    """(pos float32[N,4], vel float32[N,4], info uint16[N,4], hash uint32[N]) from a decoded HotFile."""
    b = hf["buffers"]
    n = hf["particle_count"]

    def get(*names):
        for nm in names:
            for k in b:
                if k.lower() == nm.lower():
                    return b[k]
        raise KeyError(f"{names} not in {list(b)}")
    pos = np.frombuffer(get("Position")[1], dtype=np.float32).reshape(n, 4).copy()
    vel = np.frombuffer(get("Velocity")[1], dtype=np.float32).reshape(n, 4).copy()
    info = np.frombuffer(get("Info")[1], dtype=np.uint16).reshape(n, 4).copy()
    hashv = np.frombuffer(get("Hash")[1], dtype=np.uint32).copy()
    return pos, vel, info, hashv
