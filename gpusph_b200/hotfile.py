"""GPUSPH HotFile checkpoints (src/writers/HotFile.h:44-58, src/writers/HotFile.cc:78-260): reader and writer.

A HotFile is what the reference's HotWriter saves with --checkpoint-every and what `--resume <file>` loads
(src/GPUSPH.cc:250-453). Layout (native little-endian, the reference writes its structs raw):

    header_t          version=1, buffer_count, particle_count, body_count, numOpenBoundaries, 12 reserved uints,
                      ulong iterations, double t, float dt, 3 reserved uints                      (104 bytes)
    per stored buffer encoded_buffer_t { uint name_length; char name[64]; uint element_size; uint array_count }
                      followed by element_size * particle_count bytes (first array only, HotFile.cc:212-225)
    per body          encoded_body_t (not written here: moving bodies do not resume identically in the reference
                      either, src/GPUSPH.cc:426-428)

`buffer_count` is the size of the simulation's host buffer list INCLUDING the ephemeral buffers that are not stored
(HotFile.cc:94-104); on resume the reference compares it with its own list (HotFile.cc:141) and the stored buffers'
names, in buffer-key order, with its own (HotFile.cc:242-245). For the plain WCSPH configurations of this repo the
host list is {Position (double precision) [ephemeral], Position, Velocity, Info, Hash} (src/GPUSPH.cc:868-875,
src/define_buffers.h:49-58), i.e. buffer_count = 5 and four stored buffers in that order.
"""
from __future__ import annotations

import struct

import numpy as np

HEADER = "@IIIII48xLdf12x"     # version, buffer_count, particle_count, body_count, numOpenBoundaries, iterations, t, dt
BUFHDR = "@I64sII"             # name length, name, element size, array count
VERSION = 1
# stored buffers of a plain WCSPH run, in buffer-key order: (name, numpy dtype, components, element size)
PLAIN_BUFFERS = (("Position", np.float32, 4, 16), ("Velocity", np.float32, 4, 16), ("Info", np.uint16, 4, 8),
                 ("Hash", np.uint32, 1, 4))
PLAIN_BUFFER_COUNT = 5          # + the ephemeral "Position (double precision)"


def read_hotfile(path: str) -> dict:
    out = {}
    with open(path, "rb") as f:
        h = struct.unpack(HEADER, f.read(struct.calcsize(HEADER)))
        version, nbuf, nparts, nbodies, nopen, iterations, t, dt = h
        if version != VERSION:
            raise ValueError(f"unsupported HotFile version {version}")        # HotFile.cc:166-172
        out.update(buffer_count=nbuf, particle_count=nparts, body_count=nbodies, num_open_boundaries=nopen,
                   iterations=iterations, t=t, dt=dt, buffers={})
        for _ in range(nbuf):
            raw = f.read(struct.calcsize(BUFHDR))
            if len(raw) < struct.calcsize(BUFHDR):
                break                       # ephemeral buffers are counted but not stored
            ln, name, elsize, _count = struct.unpack(BUFHDR, raw)
            if ln > 64 or elsize == 0 or elsize > 64:
                break                       # not a buffer header: the body records start here
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


def write_hotfile(path: str, pos: np.ndarray, vel: np.ndarray, info: np.ndarray, hashv: np.ndarray, *,
                  iterations: int, t: float, dt: float, buffer_count: int = PLAIN_BUFFER_COUNT,
                  num_open_boundaries: int = 0, order=None) -> None:
    """Write the state of a plain WCSPH run (no bodies, no optional buffers) as the reference's HotWriter would
    (HotFile::save, src/writers/HotFile.cc:87-118): the reference can `--resume` from it. `order`: the stored
    buffers' names in file order (default: buffer-key order, Position, Velocity, Info, Hash)."""
    n = int(pos.shape[0])
    by_name = dict(zip((b[0] for b in PLAIN_BUFFERS), zip(PLAIN_BUFFERS, (pos, vel, info, hashv))))
    order = tuple(order) if order is not None else tuple(b[0] for b in PLAIN_BUFFERS)
    if sorted(order) != sorted(by_name):
        raise ValueError(f"a plain WCSPH HotFile stores exactly {sorted(by_name)}, not {sorted(order)}")
    for (name, dtype, comps, elsize), a in by_name.values():
        if a.shape[0] != n or a.dtype.itemsize * (a.size // max(n, 1)) != elsize and n:
            raise ValueError(f"{name}: expected {n} elements of {elsize} bytes")
    with open(path, "wb") as f:
        f.write(struct.pack(HEADER, VERSION, buffer_count, n, 0, num_open_boundaries, int(iterations), float(t), float(dt)))
        for (name, dtype, comps, elsize), a in (by_name[k] for k in order):
            nm = name.encode()
            f.write(struct.pack(BUFHDR, len(nm), nm, elsize, 1))          # strcpy into a zeroed char[64]
            f.write(np.ascontiguousarray(a).view(np.uint8).tobytes())
