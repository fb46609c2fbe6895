"""ctypes binding of libsphb200.so (the C ABI declared in include/sphb200.h).

This is the only place the shared library is loaded.  There is deliberately no
fallback: if the library is missing, or no CUDA device is present when a
compute entry point is called, the caller gets an exception.
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsphb200.so")

# ---- constants mirrored from include/sphb200.h ------------------------------
ABI_VERSION = 3
OK, EINVAL, ENOMEM, ECUDA, EDTYPE, EUNSUP, ENODEV = 0, -1, -2, -3, -4, -5, -6
ERR_NEIGHBOR_OVERFLOW, ERR_CELL_OVERFLOW, ERR_STAGE_OVERFLOW, ERR_NONFINITE = 1, 2, 4, 8
ERR_OUTSIDE_BOX = 16
ERR_HINT = 128
ERR_SLAB_TIMEOUT = 256
HINT_UNIFORM_ETA = 1
SOLVER = {"SPH": 0, "RIE": 1, "DELTA": 2}
KERNEL = {"QSK": 0, "WC2K": 1, "CSK": 2, "WC4K": 3, "WC6K": 4, "GK": 5, "SGK": 6}
EOS_TAIT, EOS_RIEMANN = 0, 1
F_BC_TRICK, F_RHO_EVOL, F_RHO_RENORM, F_FREE_SLIP, F_HEAT = 1, 2, 4, 8, 16
G_NONE, G_CONST, G_BAND, G_ARRAY = 0, 1, 2, 3
BC_SET_U, BC_SET_V, BC_ZERO_DUDT, BC_ZERO_DVDT, BC_SET_P, BC_SET_T, BC_ZERO_DTDT = (
    1, 2, 4, 8, 16, 32, 64)
STEP_INTEGRATE, STEP_BC = 1, 2
VEL_REST, VEL_TGV2D, VEL_TGV3D = 0, 1, 2

VECTOR_FIELDS = ("r", "u", "v", "dudt", "dvdt", "nw")
SCALAR_FIELDS = ("rho", "p", "drhodt", "mass", "eta", "T", "dTdt", "kappa", "Cp")


class BcRule(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("u", C.c_float * 3), ("v", C.c_float * 3),
                ("p", C.c_float), ("T", C.c_float)]


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dim", C.c_int32), ("solver", C.c_int32),
        ("kernel", C.c_int32), ("eos", C.c_int32), ("flags", C.c_uint32),
        ("box", C.c_double * 3), ("dx", C.c_double), ("h", C.c_double), ("dt", C.c_double),
        ("tvf", C.c_double), ("c_ref", C.c_double), ("eta_limiter", C.c_double),
        ("artificial_alpha", C.c_double),
        ("p_ref", C.c_double), ("rho_ref", C.c_double), ("p_bg", C.c_double),
        ("gamma", C.c_double), ("u_ref", C.c_double), ("r_cutoff", C.c_double),
        ("g_mode", C.c_int32), ("g_axis", C.c_int32), ("g", C.c_double * 3),
        ("g_lo", C.c_double), ("g_hi", C.c_double),
        ("bc", BcRule * 4),
        ("bc_inflow_on", C.c_int32), ("bc_inflow_x", C.c_float), ("bc_inflow_T", C.c_float),
        ("bc_outflow_on", C.c_int32), ("bc_outflow_x", C.c_float),
        ("cell_sub", C.c_int32 * 3), ("tile", C.c_int32 * 3), ("threads", C.c_int32),
        ("list_cap", C.c_int32), ("stage_cap", C.c_int32), ("nl_cap", C.c_int32),
        ("diff_delta", C.c_float), ("diff_alpha", C.c_float), ("skin", C.c_float),
        ("hints", C.c_uint32), ("reserved", C.c_int32 * 3),
    ]


class State(C.Structure):
    _fields_ = ([(k, C.c_void_p) for k in VECTOR_FIELDS] + [(k, C.c_void_p) for k in SCALAR_FIELDS]
                + [("tag", C.c_void_p), ("g_ext", C.c_void_p)])


class Lattice(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dim", C.c_int32), ("n", C.c_int32 * 3),
        ("k_lo", C.c_int32), ("k_hi", C.c_int32), ("velocity", C.c_int32),
        ("wall_axis", C.c_int32), ("n_walls", C.c_int32),
        ("hot_lo", C.c_float), ("hot_hi", C.c_float), ("T_hot", C.c_float),
        ("dx", C.c_float), ("rho", C.c_float), ("p", C.c_float), ("mass", C.c_float),
        ("eta", C.c_float), ("T", C.c_float), ("kappa", C.c_float), ("Cp", C.c_float),
    ]


# every symbol include/sphb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "sphb200_abi_version": (C.c_int, []),
    "sphb200_strerror": (C.c_char_p, [C.c_int]),
    "sphb200_config_default": (None, [C.POINTER(Config)]),
    "sphb200_engine_bytes": (C.c_int, [C.POINTER(Config), C.c_int64, C.POINTER(C.c_size_t)]),
    "sphb200_engine_create": (C.c_int, [C.POINTER(Config), C.c_int64, C.POINTER(_P)]),
    "sphb200_engine_create_in": (C.c_int, [C.POINTER(Config), C.c_int64, _P, C.c_size_t,
                                           C.POINTER(_P)]),
    "sphb200_engine_destroy": (C.c_int, [_P]),
    "sphb200_engine_upload": (C.c_int, [_P, C.POINTER(State), C.c_int, _P]),
    "sphb200_engine_refresh": (C.c_int, [_P, C.POINTER(State), C.c_int, _P]),
    "sphb200_engine_step": (C.c_int, [_P, C.c_double, C.c_int, C.c_uint32, _P]),
    "sphb200_engine_advance_host": (C.c_int, [_P, C.c_double, C.POINTER(State), C.POINTER(State),
                                             C.c_uint32, _P]),
    "sphb200_engine_download": (C.c_int, [_P, C.POINTER(State), C.c_int, _P]),
    "sphb200_engine_vjp": (C.c_int, [_P, C.c_double, C.c_uint32, C.POINTER(State), C.POINTER(State), _P]),
    "sphb200_engine_error": (C.c_int, [_P, C.POINTER(C.c_uint32), _P]),
    "sphb200_engine_neighbor_list": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P, _P]),
    "sphb200_engine_stats": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), _P]),
    "sphb200_engine_get_stats": (C.c_int, [_P, C.POINTER(C.c_double * 20), _P]),
    "sphb200_engine_set_wall_layer": (C.c_int, [_P, _P, C.c_int, _P, C.c_double]),
    "sphb200_engine_launches": (C.c_int64, [_P]),
    "sphb200_engine_profile": (C.c_int, [_P, C.c_int]),
    "sphb200_engine_last_times": (C.c_int, [_P, C.POINTER(C.c_float * 8)]),
    "sphb200_engine_plan": (C.c_int, [_P, C.POINTER(C.c_int32 * 16)]),
    "sphb200_engine_counters": (C.c_int, [_P, C.POINTER(C.c_int64 * 8), _P]),
    "sphb200_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), _P]),
    "sphb200_slab_create": (C.c_int, [C.POINTER(Config), C.c_int, C.c_int, C.c_int64, C.c_int64,
                                      C.c_int64, C.POINTER(_P)]),
    "sphb200_slab_info": (C.c_int, [_P, C.POINTER(C.c_int64 * 16), C.POINTER(C.c_double * 4)]),
    "sphb200_slab_upload": (C.c_int, [_P, C.POINTER(State), _P, C.c_int64, C.c_int, _P]),
    "sphb200_slab_download": (C.c_int, [_P, C.POINTER(State), _P, C.c_int64, C.c_int, _P]),
    "sphb200_slab_counts": (C.c_int, [_P, C.POINTER(C.c_int32 * 8), _P]),
    "sphb200_slab_run": (C.c_int, [_P, C.c_int, C.c_double, C.c_uint32, _P, _P, _P, _P, _P,
                                   C.POINTER(C.c_int64)]),
    "sphb200_slab_set_agree": (C.c_int, [_P, C.POINTER(_P), C.c_int]),
    "sphb200_slab_signal": (C.c_int, [_P, _P]),
    "sphb200_lattice_rows": (C.c_int64, [C.POINTER(Lattice)]),
    "sphb200_init_lattice": (C.c_int, [C.POINTER(Lattice), C.POINTER(State), _P, _P]),
    "sphb200_eval_velocity": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "sphb200_add_noise": (C.c_int, [C.c_int32, C.c_int64, _P, _P, _P, C.c_double, C.c_uint64,
                                    C.POINTER(C.c_double * 3), _P]),
    "sphb200_workspace_release": (None, [_P]),
    "sphb200_workspace_bytes": (C.c_int, [C.POINTER(Config), C.c_int64, C.POINTER(C.c_size_t)]),
    "sphb200_neighbor_list": (C.c_int, [C.POINTER(Config), C.c_int64, _P, _P, C.c_int64, C.c_int,
                                        _P, _P, _P, C.c_size_t, _P]),
    "sphb200_forward": (C.c_int, [C.POINTER(Config), C.c_int64, C.POINTER(State),
                                  C.POINTER(State), _P, _P, C.c_size_t, _P]),
    "sphb200_advance": (C.c_int, [C.POINTER(Config), C.c_int64, C.c_double, C.POINTER(State),
                                  C.POINTER(State), _P, _P, C.c_size_t, _P]),
    "sphb200_advance_persistent": (C.c_int, [C.POINTER(Config), C.c_int64, C.c_double,
                                             C.POINTER(State), C.POINTER(State), _P, _P,
                                             C.c_size_t, _P]),
    "sphb200_advance_ordered": (C.c_int, [C.POINTER(Config), C.c_int64, C.c_double,
                                          C.POINTER(State), _P, C.POINTER(State), _P, _P, _P,
                                          C.c_size_t, _P]),
}

_lib = None


class Sphb200Error(RuntimeError):
    pass


def load():
    """Load libsphb200.so (built by ``__graft_entry__.build()``); no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Sphb200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.sphb200_abi_version() != ABI_VERSION:
        raise Sphb200Error("libsphb200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().sphb200_strerror(rc).decode()
        raise Sphb200Error(f"sphb200 error {rc}: {msg}")


def default_config():
    cfg = Config()
    load().sphb200_config_default(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(Config), "Config struct layout drifted from sphb200.h"
    return cfg
