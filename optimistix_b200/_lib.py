"""ctypes binding of libb200opt.so (the C ABI declared in include/b200opt.h).

There is NO CPU fallback: if the library is missing or a call fails this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200OPT_LIB") or os.path.join(_HERE, "libb200opt.so")  # env: tuning builds (tools/gram_variants.sh)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b200opt.h")

F32, F64 = 0, 1
NORM_IDS = {"max": 0, "two": 1, "rms": 2}
SOLVER_LM, SOLVER_GN, SOLVER_NEWTON, SOLVER_CHORD, SOLVER_BFGS = 0, 1, 2, 3, 4
UPDATE_BFGS, UPDATE_DFP = 0, 1
CSRC_DIR = os.path.join(_HERE, "csrc")


class Options(C.Structure):
    _fields_ = [
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("norm", C.c_int32),
        ("max_steps", C.c_int32),
        ("cauchy_termination", C.c_int32),
        ("root_near_zero", C.c_int32),
        ("kappa", C.c_double),
        ("linear_solver", C.c_int32),
        ("data_dtype", C.c_int32),
        ("scratch", C.c_void_p),
    ]


class ModelInfo(C.Structure):
    _fields_ = [
        ("id", C.c_int32),
        ("kind", C.c_int32),
        ("n", C.c_int32),
        ("n_max", C.c_int32),
        ("m_max", C.c_int32),
        ("shared_args", C.c_int32),
        ("name", C.c_char_p),
        ("params", C.c_char_p),
        ("shared_layout", C.c_char_p),
        ("problem_layout", C.c_char_p),
    ]


class B200OptError(RuntimeError):
    pass


_lib = None

_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_SIGNATURES = {
    "b200opt_version": (_i32, []),
    "b200opt_last_error": (C.c_char_p, []),
    "b200opt_result_name": (C.c_char_p, [_i32]),
    "b200opt_model_count": (_i32, []),
    "b200opt_model_info_get": (_i32, [_i32, C.POINTER(ModelInfo)]),
    "b200opt_model_id": (_i32, [C.c_char_p]),
    "b200opt_default_options": (None, [C.POINTER(Options)]),
    "b200opt_register_curve_family": (_i32, [C.c_char_p, _i32, _i32, C.c_char_p, C.c_char_p]),
    "b200opt_register_root_family": (_i32, [C.c_char_p, _i32, _i32, C.c_char_p, C.c_char_p]),
    "b200opt_least_squares": (_i32, [_i32, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _i64, C.POINTER(Options), _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_root_find": (_i32, [_i32, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp, C.POINTER(Options), _vp, _vp, _vp, _vp]),
    "b200opt_minimise": (_i32, [_i32, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _i64, C.POINTER(Options), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_least_squares_host": (_i32, [_i32, _i32, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _i64, _vp, _i64, C.POINTER(Options), _vp, _vp, _vp, _i64]),
    "b200opt_root_find_host": (_i32, [_i32, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _i64, _vp, _i64, C.POINTER(Options), _vp, _vp, _vp, _i64]),
    "b200opt_root_find_host_parts": (_i32, [_i32, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _i64, C.POINTER(_vp), C.POINTER(C.c_int32), _i32, _i64,
                                            C.POINTER(Options), _vp, _vp, _vp, _i64]),
    "b200opt_minimise_host": (_i32, [_i32, _i32, _i32, _i32, _i64, _i32, _vp, _vp, _i64, _vp, _i64, C.POINTER(Options), _vp, _vp, _vp, _i64]),
    "b200opt_host_cache_release": (None, []),
    "b200opt_lsq_step_state_size": (_i64, [_i32, _i32, _i32]),
    "b200opt_lsq_step_init": (_i32, [_i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_lsq_step": (_i32, [_i32, _i32, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, C.POINTER(Options), _vp, _vp, _vp, _vp]),
    "b200opt_lsq_step_finish": (_i32, [_i32, _i64, _i32, _i32, _vp, _vp, _vp]),
    "b200opt_root_step_state_size": (_i64, [_i32, _i32]),
    "b200opt_root_step_init": (_i32, [_i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_root_step": (_i32, [_i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_root_step_finish": (_i32, [_i32, _i64, _i32, _vp, _vp, _vp]),
    "b200opt_lu_solve": (_i32, [_i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_bfgs_step_state_size": (_i64, [_i32, _i32]),
    "b200opt_bfgs_step_init": (_i32, [_i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_bfgs_step": (_i32, [_i32, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_bfgs_step_finish": (_i32, [_i32, _i64, _i32, _vp, _vp, _vp]),
    "b200opt_gn_gram_workspace_size": (_i64, [_i32]),
    "b200opt_gn_gram": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "b200opt_gn_gram_linear": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _i64, _i32, _vp, _vp, _vp]),
    "b200opt_gn_solve": (_i32, [_vp, _i32, _f64, _vp, _vp, _vp, _vp]),
    "b200opt_gn_solve_parts": (_i32, [_vp, _i32, _i64, _i32, _f64, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_xchg_bytes": (_i64, [_i32, _i32]),
    "b200opt_xchg_alloc": (_i32, [_i64, C.POINTER(_vp), _vp]),
    "b200opt_xchg_open": (_i32, [_vp, C.POINTER(_vp)]),
    "b200opt_xchg_close": (_i32, [_vp]),
    "b200opt_xchg_free": (_i32, [_vp]),
    "b200opt_gn_gram_linear_xchg": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_float, C.c_float, _i64, _i32, _vp, _vp, _vp, _i32, _i32, _i32,
                                          C.c_uint32, _vp]),
    "b200opt_gn_solve_xchg": (_i32, [_vp, _i32, _i32, C.c_uint32, _i32, _f64, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_gn_solve_fast": (_i32, [_vp, _i32, _i64, _i32, _f64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200opt_gn_state_size": (_i64, []),
    "b200opt_gn_init": (_i32, [_vp, _vp, _i32, _f64, _f64, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "b200opt_gn_iterate": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f64, _f64, _f64, _i32, _vp]),
    "b200opt_linear_residual": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, C.c_float, C.c_float, _vp, _vp]),
}


def load():
    """Load the in-tree library; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200OptError(
            f"{LIB_PATH} is missing: build it with `python -m optimistix_b200.build` "
            "(there is no CPU fallback for the CUDA hot path)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str):
    if code != 0:
        msg = load().b200opt_last_error().decode()
        raise B200OptError(f"{what} failed ({code}): {msg}")


LINSOLVE_NORMAL_CHOLESKY, LINSOLVE_QR = 0, 1
DATA_AS_COMPUTE, DATA_F16, DATA_U16, DATA_I16 = 0, 1, 2, 3


def linear_solver_id(linear_solver) -> int:
    """Map the `linear_solver=` field of LevenbergMarquardt / GaussNewton to the kernels' two solvers.  Anything the kernels
    do not implement raises: the argument is never silently ignored."""
    if linear_solver is None:
        return LINSOLVE_NORMAL_CHOLESKY
    name = type(linear_solver).__name__
    if name in ("NormalCholesky", "Cholesky"):
        return LINSOLVE_NORMAL_CHOLESKY
    if name == "QR":
        return LINSOLVE_QR
    if name == "AutoLinearSolver" and getattr(linear_solver, "well_posed", None) is None:
        return LINSOLVE_NORMAL_CHOLESKY  # GaussNewton's default token (gauss_newton.py:387): see _solvers.py
    raise NotImplementedError(f"linear_solver={linear_solver!r}: the CUDA least-squares kernels implement NormalCholesky() (Cholesky of "
                              "J^T J + lambda I, the default) and QR() (Householder QR of the augmented system)")


def make_options(rtol, atol, norm="max", max_steps=256, cauchy_termination=True, root_near_zero=False, kappa=1e-2, linear_solver=None,
                 data_dtype=0, scratch=None):
    if isinstance(norm, str):
        norm_id = NORM_IDS[norm]
    else:
        norm_id = NORM_IDS[getattr(norm, "norm_name")]
    return Options(float(rtol), float(atol), norm_id, int(max_steps), int(bool(cauchy_termination)), int(bool(root_near_zero)), float(kappa),
                   linear_solver_id(linear_solver), int(data_dtype), scratch)
