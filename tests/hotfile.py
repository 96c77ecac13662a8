"""Decoder for GPUSPH HotFile checkpoints — now part of the package (gpusph_b200/hotfile.py); kept as an alias for the
golden-fixture tooling (oracle/gen_golden.py)."""
from gpusph_b200.hotfile import BUFHDR, HEADER, particle_arrays, read_hotfile, write_hotfile  # noqa: F401
