"""ctypes binding of the C ABI declared in include/b200sph.h.

This module is plumbing only: it loads ``libb200sph.so`` (built in-tree by
``__graft_entry__.build()``), declares the prototypes and turns error codes into
exceptions the way the C++ adapter does (B200SPH_EINVAL -> ValueError, the rest ->
RuntimeError). There is no fallback: if the library is missing, importing the engines
fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200SPH_LIB: alternative build of the same library (kernel tuning experiments, tools/build_variants.sh)
LIB_PATH = os.environ.get("B200SPH_LIB") or os.path.join(_HERE, "libb200sph.so")

ABI_VERSION = 4
NEIBLIST_BLOCK = 2097152          # B200SPH_NEIBLIST_BLOCK: particles per block of the neighbour-list layout (include/b200sph.h)
MAX_FLUIDS = 4
MAX_PLANES = 8

# enums (values are the reference's, src/particledefine.h:79-224, src/visc_spec.h)
KERNEL_WENDLAND = 3
SPH_F1 = 1
RHODIFF_NONE, RHODIFF_FERRARI, RHODIFF_COLAGROSSI, RHODIFF_BREZZI = 0, 1, 2, 3
LJ_BOUNDARY, MK_BOUNDARY, SA_BOUNDARY, DYN_BOUNDARY = 0, 1, 2, 3
PERIODIC_X, PERIODIC_Y, PERIODIC_Z = 1, 2, 4
RHEOLOGY_INVISCID, RHEOLOGY_NEWTONIAN = 0, 1
TURB_LAMINAR, TURB_ARTIFICIAL = 0, 1
COMPVISC_KINEMATIC, COMPVISC_DYNAMIC = 0, 1
VISCMODEL_MORRIS, VISCMODEL_MONAGHAN, VISCMODEL_ESPANOL_REVENGA = 0, 1, 2
# simulation flags (src/simflags.h:71-86)
ENABLE_DTADAPT, ENABLE_XSPH, ENABLE_PLANES, ENABLE_DEM = 1, 2, 4, 8
ENABLE_MOVING_BODIES, ENABLE_INLET_OUTLET, ENABLE_WATER_DEPTH, ENABLE_DENSITY_SUM = 16, 32, 64, 128
ENABLE_GAMMA_QUADRATURE, ENABLE_REPACKING, ENABLE_INTERNAL_ENERGY, ENABLE_MULTIFLUID = 256, 512, 1024, 2048
AVG_ARITHMETIC, AVG_HARMONIC, AVG_GEOMETRIC = 0, 1, 2

PT_FLUID, PT_BOUNDARY, PT_VERTEX, PT_TESTPOINT = 0, 1, 2, 3
FG_COMPUTE_FORCE = 1 << 3
FG_MOVING_BOUNDARY = 1 << 4

CELL_EMPTY = 0xFFFFFFFF
CELL_HASH_MAX = 0xFFFFFFFF
NEIBS_END = 0xFFFF

E_INVAL, E_UNSUP, E_CUDA, E_NODEV, E_NOMEM = -1, -2, -3, -4, -5


class Params(C.Structure):
    """struct b200sph_params (include/b200sph.h)."""
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("world_origin", C.c_float * 3),
        ("cell_size", C.c_float * 3),
        ("grid_size", C.c_uint32 * 3),
        ("coord", C.c_uint32 * 3),
        ("periodic", C.c_uint32),
        ("neiblistsize", C.c_uint32),
        ("neibboundpos", C.c_uint32),
        ("neiblist_stride", C.c_uint32),
        ("nl_sq_influence_radius", C.c_float),
        ("kerneltype", C.c_uint32), ("sph_formulation", C.c_uint32),
        ("densitydiffusiontype", C.c_uint32), ("boundarytype", C.c_uint32),
        ("rheologytype", C.c_uint32), ("turbmodel", C.c_uint32), ("compvisc", C.c_uint32),
        ("viscmodel", C.c_uint32), ("viscavgop", C.c_uint32),
        ("is_const_visc", C.c_uint32),
        ("slength", C.c_float), ("influenceradius", C.c_float), ("deltap", C.c_float),
        ("density_diff_coeff", C.c_float), ("dtadaptfactor", C.c_float),
        ("num_fluids", C.c_uint32),
        ("rho0", C.c_float * MAX_FLUIDS), ("bcoeff", C.c_float * MAX_FLUIDS),
        ("gammacoeff", C.c_float * MAX_FLUIDS), ("sscoeff", C.c_float * MAX_FLUIDS),
        ("sspowercoeff", C.c_float * MAX_FLUIDS), ("visccoeff", C.c_float * MAX_FLUIDS),
        ("gravity", C.c_float * 3),
        ("artvisccoeff", C.c_float), ("epsartvisc", C.c_float),
        ("max_sound_speed_cfl", C.c_float), ("max_kinvisc", C.c_float),
        ("dtadapt", C.c_uint32),
        # ABI version 2
        ("simflags", C.c_uint32),
        ("epsxsph", C.c_float), ("monaghan_visc_coeff", C.c_float),
        ("visc2coeff", C.c_float * MAX_FLUIDS),
        ("r0", C.c_float), ("dcoeff", C.c_float), ("p1coeff", C.c_float), ("p2coeff", C.c_float), ("partsurf", C.c_float),
        # ABI version 4
        ("neiblist_block", C.c_uint32),
    ]

    def copy(self) -> "Params":
        p = Params()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(Params))
        return p

    @property
    def num_cells(self) -> int:
        return int(self.grid_size[0]) * int(self.grid_size[1]) * int(self.grid_size[2])


class NeibsInfo(C.Structure):
    _fields_ = [
        ("num_interactions", C.c_int32),
        ("max_fluid_boundary_neibs", C.c_int32),
        ("max_vertex_neibs", C.c_int32),
        ("has_too_many_neibs", C.c_int32),
        ("has_max_neibs", C.c_int32 * 3),
    ]


class ForcesArgs(C.Structure):
    """struct b200sph_forces_args (include/b200sph.h): the complete argument set of AbstractForcesEngine::basicstep."""
    _fields_ = [
        ("pos", C.c_void_p), ("vel", C.c_void_p), ("info", C.c_void_p),
        ("hash", C.c_void_p), ("cell_start", C.c_void_p), ("neibs_list", C.c_void_p),
        ("forces", C.c_void_p), ("cfl", C.c_void_p),
        ("rb_forces", C.c_void_p), ("rb_torques", C.c_void_p), ("xsph", C.c_void_p),
        ("num_particles", C.c_uint32), ("from_particle", C.c_uint32), ("to_particle", C.c_uint32), ("cfl_offset", C.c_uint32),
        ("dt", C.c_float), ("step", C.c_int), ("dt_from_device", C.c_int),
        ("packed", C.c_void_p),
    ]


class FusedEulerArgs(C.Structure):
    """struct b200sph_fused_euler_args (include/b200sph.h): the integration half of b200sph_forces_euler."""
    _fields_ = [
        ("old_pos", C.c_void_p), ("old_vel", C.c_void_p), ("new_pos", C.c_void_p), ("new_vel", C.c_void_p),
        ("dt", C.c_float), ("step", C.c_int), ("dt_from_device", C.c_int),
        ("new_packed", C.c_void_p),
    ]


class HostStepArgs(C.Structure):
    """struct b200sph_host_step_args (include/b200sph.h): one time step of a state that lives in host memory."""
    _fields_ = [
        ("host_pos", C.c_void_p), ("host_vel", C.c_void_p),
        ("pos", C.c_void_p), ("vel", C.c_void_p), ("pos_star", C.c_void_p), ("vel_star", C.c_void_p),
        ("info", C.c_void_p), ("hash", C.c_void_p), ("cell_start", C.c_void_p), ("neibs_list", C.c_void_p),
        ("forces", C.c_void_p), ("cfl", C.c_void_p), ("xsph", C.c_void_p),
        ("cfl_elements", C.c_uint32), ("num_particles", C.c_uint32),
        ("stripe_bounds", C.POINTER(C.c_uint32)), ("num_stripes", C.c_uint32), ("resident", C.c_int),
    ]


MAX_STRIPES = 32


class ReorderExtra(C.Structure):
    _fields_ = [("unsorted", C.c_void_p), ("sorted", C.c_void_p), ("elem_size", C.c_uint32)]


# every symbol include/b200sph.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_U = C.c_uint32
PROTOTYPES = {
    "b200sph_last_error": (C.c_char_p, []),
    "b200sph_abi_version": (C.c_int, []),
    "b200sph_device_count": (C.c_int, []),
    "b200sph_create": (C.c_int, [C.POINTER(Params), C.POINTER(_P)]),
    "b200sph_destroy": (C.c_int, [_P]),
    "b200sph_validate": (C.c_int, [C.POINTER(Params)]),
    "b200sph_set_stream": (C.c_int, [_P, _P]),
    "b200sph_set_gravity": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "b200sph_get_neibboundpos": (C.c_int, [_P, C.POINTER(_U)]),
    "b200sph_set_planes": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int]),
    "b200sph_forces_ex": (C.c_int, [_P, C.POINTER(ForcesArgs), C.POINTER(_U)]),
    "b200sph_euler_ex": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _U, _U, C.c_float, C.c_int, C.c_int]),
    "b200sph_filter_shepard": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _U, _U]),
    "b200sph_filter_mls": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _U, _U]),
    "b200sph_testpoints": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _U, _U]),
    "b200sph_calc_hash": (C.c_int, [_P, _P, _P, _P, _P, _P, _U]),
    "b200sph_fix_hash": (C.c_int, [_P, _P, _P, _P, _P, _U]),
    "b200sph_sort": (C.c_int, [_P, _P, _P, _P, _U]),
    "b200sph_reorder": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.POINTER(ReorderExtra), _U, _P, _P, _P, _U, _P]),
    "b200sph_neibs_resetinfo": (C.c_int, [_P]),
    "b200sph_neibs_getinfo": (C.c_int, [_P, C.POINTER(NeibsInfo)]),
    "b200sph_build_neibs": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _U, _U]),
    "b200sph_fmax_elements": (_U, [_U]),
    "b200sph_fmax_temp_elements": (_U, [_U]),
    "b200sph_round_particles": (_U, [_U]),
    "b200sph_forces": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _U, _U, _U, _U, C.POINTER(_U)]),
    "b200sph_set_rbcg": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int]),
    "b200sph_set_rbcg_euler": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int]),
    "b200sph_set_rbstart": (C.c_int, [_P, C.POINTER(C.c_int), C.c_int]),
    "b200sph_set_rbtrans": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    "b200sph_set_rbsteprot": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    "b200sph_set_rblinearvel": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    "b200sph_set_rbangularvel": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    "b200sph_forces_bodies": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _U, _U, _U, _U, C.POINTER(_U)]),
    "b200sph_euler_packed": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _U, _U, C.c_float, C.c_int, C.c_int]),
    "b200sph_unpack_state": (C.c_int, [_P, _P, _P, _P, _U, _U]),
    "b200sph_pack_state": (C.c_int, [_P, _P, _P, _P, _U, _U]),
    "b200sph_reduce_rb_forces": (C.c_int, [_P, _P, _P, _P, C.POINTER(_U), C.POINTER(C.c_float), C.POINTER(C.c_float), _U, _U]),
    "b200sph_eos_probe": (C.c_int, [_P, _P, _P, _P, _U]),
    "b200sph_dtreduce": (C.c_int, [_P, _P, _P, _U, C.POINTER(C.c_float)]),
    "b200sph_dtreduce_ex": (C.c_int, [_P, _P, _P, _U, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float)]),
    "b200sph_cflmax": (C.c_int, [_P, _P, _U, _P]),
    "b200sph_dt_from_cfl": (C.c_int, [_P, C.c_float, C.POINTER(C.c_float)]),
    "b200sph_step_set_dt": (C.c_int, [_P, C.c_float]),
    "b200sph_dtreduce_async": (C.c_int, [_P, _P, _U, C.c_int]),
    "b200sph_euler_async": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _U, _U, C.c_int]),
    "b200sph_step_end": (C.c_int, [_P]),
    "b200sph_step_query": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "b200sph_euler": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _U, _U, C.c_float, C.c_int]),
    "b200sph_forces_euler": (C.c_int, [_P, C.POINTER(ForcesArgs), C.POINTER(FusedEulerArgs), C.POINTER(_U)]),
    "b200sph_step_host": (C.c_int, [_P, C.POINTER(HostStepArgs)]),
    "b200sph_host_upload": (C.c_int, [_P, _P, _P, _P, _P, _U]),
    "b200sph_host_fence": (C.c_int, [_P]),
    "b200sph_host_sync": (C.c_int, [_P]),
}

_lib = None


class B200Error(RuntimeError):
    """CUDA / device / unsupported-option failure reported by the library."""


class B200Unsupported(B200Error):
    """Option combination that is not implemented (the library never falls back)."""


def load():
    """Load libb200sph.so (once) and declare all prototypes. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.b200sph_abi_version() != ABI_VERSION:
        raise ImportError("libb200sph.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Turn a C return code into the exception the reference would have thrown."""
    if rc == 0:
        return
    msg = load().b200sph_last_error().decode("utf-8", "replace")
    if rc == E_INVAL:
        raise ValueError(msg)            # reference: std::invalid_argument
    if rc == E_UNSUP:
        raise B200Unsupported(msg)
    if rc == E_NOMEM:
        raise MemoryError(msg)
    raise B200Error(msg)                 # reference: std::runtime_error from CUDA_SAFE_CALL
