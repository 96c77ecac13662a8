"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, validates options like the reference's framework factory (unsupported combinations
fail loudly, src/cuda/cudasimframework.cu:147-155) and refuses to run without a GPU."""
import ctypes as C
import os
import re

import pytest

from gpusph_b200 import capi
from gpusph_b200.problems import lattice_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200sph.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200sph_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libb200sph.so does not export {s}"
    # and the ctypes prototypes cover the header exactly
    assert sorted(capi.PROTOTYPES) == syms


def test_abi_version_and_struct_size():
    lib = capi.load()
    assert lib.b200sph_abi_version() == capi.ABI_VERSION
    # layout check: the C compiler and ctypes must agree on sizeof(b200sph_params)
    src = ('#include "b200sph.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu", sizeof(b200sph_params), '
           'sizeof(b200sph_neibs_info), sizeof(b200sph_forces_args), sizeof(b200sph_fused_euler_args), '
           'sizeof(b200sph_host_step_args), sizeof(b200sph_reorder_extra));return 0;}')
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == C.sizeof(capi.Params)
    assert int(out[1]) == C.sizeof(capi.NeibsInfo)
    assert int(out[2]) == C.sizeof(capi.ForcesArgs)
    assert int(out[3]) == C.sizeof(capi.FusedEulerArgs)
    assert int(out[4]) == C.sizeof(capi.HostStepArgs)
    assert int(out[5]) == C.sizeof(capi.ReorderExtra)
    assert capi.MAX_STRIPES == int(re.search(r"#define B200SPH_MAX_STRIPES (\d+)", open(os.path.join(ROOT, "include", "b200sph.h")).read()).group(1))


def test_validate_accepts_supported_and_rejects_unsupported():
    lib = capi.load()
    params, _ = lattice_problem(4)
    assert lib.b200sph_validate(C.byref(params)) == 0
    for field, bad in [("kerneltype", 1), ("sph_formulation", 3), ("boundarytype", capi.SA_BOUNDARY),
                       ("boundarytype", capi.LJ_BOUNDARY), ("simflags", capi.ENABLE_DTADAPT | capi.ENABLE_DEM),
                       ("turbmodel", 2), ("rheologytype", 3)]:
        p = params.copy()
        setattr(p, field, bad)
        assert lib.b200sph_validate(C.byref(p)) == capi.E_UNSUP, field
        with pytest.raises(capi.B200Unsupported):
            capi.check(lib.b200sph_validate(C.byref(p)))
        assert b"unsupported" in lib.b200sph_last_error()
    # the WHOLE of SimParams::simflags reaches the library: every flag outside the allow-list is refused by name,
    # none is silently dropped (the reference would run the energy equation, open boundaries, ... for these)
    for flag, name in [(capi.ENABLE_INTERNAL_ENERGY, b"ENABLE_INTERNAL_ENERGY"), (capi.ENABLE_INLET_OUTLET, b"ENABLE_INLET_OUTLET"),
                       (capi.ENABLE_WATER_DEPTH, b"ENABLE_WATER_DEPTH"), (capi.ENABLE_DENSITY_SUM, b"ENABLE_DENSITY_SUM"),
                       (capi.ENABLE_GAMMA_QUADRATURE, b"ENABLE_GAMMA_QUADRATURE"), (1 << 12, b"unknown")]:
        p = params.copy()
        p.simflags = capi.ENABLE_DTADAPT | flag
        assert lib.b200sph_validate(C.byref(p)) == capi.E_UNSUP, name
        assert name in lib.b200sph_last_error()
    # DamBreak3D's own flag word: DTADAPT | REPACKING (src/problems/DamBreak3D.cu:53-57), with a body: MOVING_BODIES
    for ok in (capi.ENABLE_DTADAPT | capi.ENABLE_REPACKING, capi.ENABLE_DTADAPT | capi.ENABLE_REPACKING | capi.ENABLE_MOVING_BODIES):
        p = params.copy()
        p.simflags = ok
        assert lib.b200sph_validate(C.byref(p)) == 0
    # neighbour-list block: 0 (default) or a power of two >= 32
    for blk, rc in [(0, 0), (128, 0), (1 << 21, 0), (100, capi.E_INVAL), (16, capi.E_INVAL)]:
        p = params.copy()
        p.neiblist_block = blk
        assert lib.b200sph_validate(C.byref(p)) == rc, blk
    p = params.copy()
    p.num_fluids = 2                    # more than one fluid needs ENABLE_MULTIFLUID (src/ProblemCore.cc:108-112)
    assert lib.b200sph_validate(C.byref(p)) == capi.E_INVAL
    p.simflags |= capi.ENABLE_MULTIFLUID
    assert lib.b200sph_validate(C.byref(p)) == 0
    # everything reachable from the CLI options of DamBreak3D / Poiseuille is accepted
    for field, ok in [("densitydiffusiontype", capi.RHODIFF_BREZZI), ("simflags", capi.ENABLE_DTADAPT | capi.ENABLE_XSPH),
                      ("simflags", capi.ENABLE_DTADAPT | capi.ENABLE_PLANES)]:
        p = params.copy()
        setattr(p, field, ok)
        assert lib.b200sph_validate(C.byref(p)) == 0, field
    for vm in (capi.VISCMODEL_MONAGHAN, capi.VISCMODEL_ESPANOL_REVENGA):
        p = params.copy()
        p.rheologytype, p.viscmodel = capi.RHEOLOGY_NEWTONIAN, vm
        assert lib.b200sph_validate(C.byref(p)) == 0
    p = params.copy()
    p.rheologytype, p.viscmodel = capi.RHEOLOGY_NEWTONIAN, 3
    assert lib.b200sph_validate(C.byref(p)) == capi.E_INVAL
    p = params.copy()
    p.abi_version = 99
    assert lib.b200sph_validate(C.byref(p)) == capi.E_INVAL
    p = params.copy()
    p.grid_size[1] = 0
    with pytest.raises(ValueError):
        capi.check(lib.b200sph_validate(C.byref(p)))
    p = params.copy()
    p.coord[0] = p.coord[1]
    assert lib.b200sph_validate(C.byref(p)) == capi.E_INVAL


def test_neighbour_list_layout_helpers_follow_the_documented_formula():
    """engines.neibs_list_rows / neibs_list_blocked against the formula of include/b200sph.h (B200SPH_NEIBLIST_BLOCK)."""
    import torch
    from gpusph_b200.engines import neibs_list_blocked, neibs_list_rows
    rows = 5
    for A, B in [(128, 128), (300, 128), (1000, 256), (77, 128), (4096, 32), (70, 0)]:
        x = torch.arange(rows * A, dtype=torch.int32).view(rows, A)
        b = neibs_list_blocked(x, B)
        assert torch.equal(neibs_list_rows(b, B), x)
        blk = B or capi.NEIBLIST_BLOCK
        flat = b.reshape(-1)
        for i in {0, A - 1, A // 2, min(A - 1, blk + 2)}:
            for k in (0, rows - 1):
                b0 = i // blk * blk
                w = min(blk, A - b0)
                assert flat[b0 * rows + k * w + (i - b0)] == x[k, i]
    # the Python constant is the header's
    hdr = open(os.path.join(ROOT, "include", "b200sph.h")).read()
    assert int(re.search(r"#define\s+B200SPH_NEIBLIST_BLOCK\s+(\d+)", hdr).group(1)) == capi.NEIBLIST_BLOCK
    # up to one block of particles the layout IS the reference's interleaved one
    x = torch.arange(rows * 500, dtype=torch.int32).view(rows, 500)
    assert torch.equal(neibs_list_blocked(x), x)


def test_size_helpers_match_reference_formulas():
    lib = capi.load()
    # getFmaxElements = round_up(div_up(n,128),4); round_particles (src/cuda/forces.cu:540-554,961-965)
    for n in [0, 1, 127, 128, 129, 511, 512, 513, 84444, 8_000_000]:
        assert lib.b200sph_fmax_elements(n) == ((n + 127) // 128 + 3) // 4 * 4
        assert lib.b200sph_round_particles(n) == (n // 128) * 128
    assert lib.b200sph_fmax_temp_elements(4) == 1
    assert lib.b200sph_fmax_temp_elements(1024) == 1
    assert lib.b200sph_fmax_temp_elements(1028) == 4
    assert lib.b200sph_fmax_temp_elements(62500) == 64


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly (never route through a CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load()
    assert lib.b200sph_device_count() == 0
    params, parts = lattice_problem(4)
    h = C.c_void_p()
    assert lib.b200sph_create(C.byref(params), C.byref(h)) == capi.E_NODEV
    assert b"no CPU fallback" in lib.b200sph_last_error()
    from gpusph_b200.simulation import Worker
    with pytest.raises(capi.B200Error):
        Worker(params, parts)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under gpusph_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gpusph_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
