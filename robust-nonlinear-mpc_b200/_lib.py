"""ctypes binding of ``librnmpc_b200.so`` (C ABI: ``include/rnmpc_b200.h``)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RNMPC_B200_LIB") or os.path.join(_HERE, "librnmpc_b200.so")      # the variable: experiment builds (tools/)


class LibraryNotBuilt(RuntimeError):
    pass


class RnmpcError(RuntimeError):
    pass


class Config(C.Structure):
    """``rnmpc_config`` of include/rnmpc_b200.h."""

    _fields_ = [
        ("model_id", C.c_int), ("N", C.c_int), ("rti", C.c_int), ("sls_steps", C.c_int),
        ("Q", C.POINTER(C.c_double)), ("R", C.POINTER(C.c_double)), ("Qf", C.POINTER(C.c_double)),
        ("Q_reg", C.POINTER(C.c_double)), ("R_reg", C.POINTER(C.c_double)), ("Q_reg_f", C.POINTER(C.c_double)),
        ("E", C.POINTER(C.c_double)),
        ("x_ub", C.POINTER(C.c_double)), ("x_lb", C.POINTER(C.c_double)),
        ("u_ub", C.POINTER(C.c_double)), ("u_lb", C.POINTER(C.c_double)),
        ("xf_ub", C.POINTER(C.c_double)), ("xf_lb", C.POINTER(C.c_double)),
        ("h", C.c_double), ("qp_max_iter", C.c_int), ("qp_tol_mu", C.c_double), ("qp_tol_res", C.c_double),
    ]


# every symbol include/rnmpc_b200.h declares: name -> (restype, argtypes)
_VP, _I, _D, _LL = C.c_void_p, C.c_int, C.c_double, C.c_longlong
SYMBOLS = {
    "rnmpc_last_error": (C.c_char_p, []),
    "rnmpc_model_dims": (_I, [_I, C.POINTER(_I), C.POINTER(_I)]),
    "rnmpc_create": (_I, [C.POINTER(_VP), C.POINTER(Config), _I, _I]),
    "rnmpc_destroy": (_I, [_VP]),
    "rnmpc_ddyn": (_I, [_VP, _I, _VP, _VP, _VP, _D, _VP, _VP]),
    "rnmpc_linearize": (_I, [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rnmpc_sls": (_I, [_VP, _I] + [_VP] * 13),
    "rnmpc_qp": (_I, [_VP, _I, _I] + [_VP] * 16),
    "rnmpc_qp_general": (_I, [_VP, _I] + [_VP] * 15),
    "rnmpc_sls_duals": (_I, [_VP, _I] + [_VP] * 15),
    "rnmpc_costates": (_I, [_VP, _I, _VP, _VP]),
    "rnmpc_soft_nmpc": (_I, [_VP, _I, _VP, _VP, _VP, _D, _D, _I, _D, _VP, _VP, _VP, _VP, _VP]),
    "rnmpc_set_nominal": (_I, [_VP, _I, _VP, _VP, _VP]),
    "rnmpc_init_nominal": (_I, [_VP, _I, _VP, _I, _VP]),
    "rnmpc_solve": (_I, [_VP, _I, _VP, _VP]),
    "rnmpc_reset_warm_start": (_I, [_VP, _I, _VP]),
    "rnmpc_plant_step": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "rnmpc_get_buffer": (_I, [_VP, C.c_char_p, C.POINTER(_VP), C.POINTER(_LL)]),
    "rnmpc_enable_phi": (_I, [_VP, _I]),
    "rnmpc_copy_buffer": (_I, [_VP, C.c_char_p, _VP, _LL]),
    "rnmpc_solve_host": (_I, [_VP, _I] + [_VP] * 7),
    "rnmpc_closed_loop_host": (_I, [_VP, _I, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP]),
    "rnmpc_fp64_peak": (_I, [_I, C.POINTER(_D)]),
    "rnmpc_fp64_mma_peak": (_I, [_I, C.POINTER(_D)]),
    "rnmpc_set_option": (_I, [_VP, C.c_char_p, _D]),
    "rnmpc_enable_timing": (_I, [_VP, _I]),
    "rnmpc_get_timings": (_I, [_VP, C.POINTER(_D)]),
    "rnmpc_kernel_time": (_I, [_VP, C.c_char_p, C.POINTER(_D), C.POINTER(C.c_longlong)]),
    "rnmpc_launch_count": (_LL, [_VP]),
}

_lib = None


def load_library():
    """Load the CUDA library; raises LibraryNotBuilt when it has not been compiled (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryNotBuilt(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  rnmpc_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load_library().rnmpc_last_error()
        raise RnmpcError(f"rnmpc error {rc}: {msg.decode() if msg else ''}")
