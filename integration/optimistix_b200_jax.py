"""Reference-side glue for patrick-kidger/optimistix: registers the XLA FFI targets of
integration/b200opt_xla.cc and exposes the registered-family fast path of `optx.least_squares`.

NOT IMPORTED OR TESTED IN THIS REPOSITORY (jax is not installed in the build image); the C ABI underneath is
what the test-suite exercises (through ctypes, optimistix_b200/_lib.py).  tests/test_capi_symbols.py checks that every
b200opt_* function used by the C++ shim is declared in include/b200opt.h.
"""
import ctypes

import jax
import jax.numpy as jnp

_lib = ctypes.cdll.LoadLibrary("libb200opt_xla.so")
for _name, _sym in (("b200opt_least_squares", "B200optLeastSquares"), ("b200opt_lsq_step", "B200optLsqStep"),
                    ("b200opt_root_step", "B200optRootStep"), ("b200opt_bfgs_step", "B200optBfgsStep"),
                    ("b200opt_lu_solve", "B200optLuSolve")):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_lib, _sym)), platform="CUDA")

SOLVER_LM, SOLVER_GN = 0, 1
NORM_MAX = 0


def least_squares_registered(model_id, solver, y0, shared, prob, *, max_steps=256, gauss_newton=False):
    """The body that replaces `iterative_solve(...)` in optimistix/_least_squares.py:115-135 when `fn` is a registered
    model family and `solver` is LevenbergMarquardt / GaussNewton with the default search and descent.  Under `jax.vmap`
    the batch axis is already the leading axis the kernel expects: vmap_method="broadcast_all" hands the whole batch to ONE
    call (one persistent kernel for the batch instead of a masked while_loop)."""
    B, n = y0.shape
    out = (jax.ShapeDtypeStruct((B, n), y0.dtype), jax.ShapeDtypeStruct((B,), jnp.int32), jax.ShapeDtypeStruct((B,), jnp.int32))
    return jax.ffi.ffi_call("b200opt_least_squares", out, vmap_method="broadcast_all")(
        y0, shared, prob, solver=SOLVER_GN if gauss_newton else SOLVER_LM, model_id=model_id, rtol=float(solver.rtol),
        atol=float(solver.atol), norm=NORM_MAX, max_steps=int(max_steps),
        # the solver's `linear_solver=` field: lx.QR() (the reference's default for LevenbergMarquardt) -> B200OPT_LINSOLVE_QR = 1,
        # the normal equations of the north_star -> 0; 16-bit data rows are handed over as they are (B200OPT_DATA_F16/U16/I16)
        linear_solver=1 if type(solver.descent.linear_solver).__name__ == "QR" and getattr(solver, "b200_reference_linear_algebra", False) else 0,
        data_dtype={"float16": 1, "uint16": 2, "int16": 3}.get(str(prob.dtype), 0))
