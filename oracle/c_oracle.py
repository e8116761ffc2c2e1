"""ctypes wrapper of oracle/c/liboracle.so (C + OpenMP restatement of the LM step; TEST INFRASTRUCTURE)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_SO = os.path.join(_DIR, "liboracle.so")
_SO_FMA = os.path.join(_DIR, "liboracle_fma.so")  # same sources, a*b+c contracted to fma (what nvcc does by default)
_lib = None
_lib_fma = None
MODELS = {"exp_decay": 0, "gauss2": 1}


def build():
    subprocess.run(["make", "-C", _DIR], check=True, capture_output=True)
    return _SO


def _nb_sig(lib):
    lib.oracle_newton_implicit_euler.restype = C.c_int
    lib.oracle_newton_implicit_euler.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.oracle_bfgs_rosenbrock.restype = C.c_int
    lib.oracle_bfgs_rosenbrock.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_int]


def _lm_sig(lib):
    _nb_sig(lib)
    lib.oracle_lm_solve.restype = C.c_int
    lib.oracle_lm_solve.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_int]


def fma_available() -> bool:
    """The contracted build exists and this CPU can run it (x86 FMA3)."""
    if not os.path.exists(_SO_FMA):
        try:
            build()
        except Exception:  # noqa: BLE001
            return False
    if not os.path.exists(_SO_FMA):
        return False
    try:
        with open("/proc/cpuinfo") as fh:
            return " fma " in fh.read().replace("\n", " ")
    except OSError:
        return False


def load_fma():
    global _lib_fma
    if _lib_fma is None:
        if not fma_available():
            raise RuntimeError("liboracle_fma.so is not available on this machine")
        _lib_fma = C.CDLL(_SO_FMA)
        _lm_sig(_lib_fma)
    return _lib_fma


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lm_sig(_lib)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return load().oracle_max_threads()


def levenberg_marquardt(model: str, y0, x, d, *, rtol, atol, max_steps=256, linsolve="qr", nthreads=0, fma=False):
    """linsolve: "qr" | "gram" | "kernel" | "kernel_exact_pred" (oracle/c/lm_oracle.c); fma=True: the contracted build."""
    lib = load_fma() if fma else load()
    y0 = np.ascontiguousarray(y0)
    dt = y0.dtype
    x = np.ascontiguousarray(x, dtype=dt)
    d = np.ascontiguousarray(d, dtype=dt)
    B, n = y0.shape
    m = x.shape[0]
    y = np.empty_like(y0)
    steps = np.empty(B, np.int32)
    res = np.empty(B, np.int32)
    rc = lib.oracle_lm_solve(MODELS[model], 1 if dt == np.float64 else 0, B, n, m, y0.ctypes.data, x.ctypes.data, d.ctypes.data,
                             d.shape[1], rtol, atol, max_steps, {"qr": 0, "gram": 1, "kernel": 2, "kernel_exact_pred": 3}[linsolve], y.ctypes.data, steps.ctypes.data,
                             res.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_lm_solve failed: {rc}")
    return dict(value=y, num_steps=steps, result=res)


def newton_implicit_euler(y0, A, B, y_prev, dt, *, rtol, atol, max_steps=256, chord=False, nthreads=0):
    """Newton / Chord on the implicit-Euler family (fp64, max_norm, Cauchy termination), C + OpenMP."""
    lib = load()
    y0, A, B, y_prev, dt = (np.ascontiguousarray(a, dtype=np.float64) for a in (y0, A, B, y_prev, dt))
    Bn, n = y0.shape
    y = np.empty_like(y0)
    steps = np.empty(Bn, np.int32)
    res = np.empty(Bn, np.int32)
    rc = lib.oracle_newton_implicit_euler(Bn, n, A.ctypes.data, B.ctypes.data, y_prev.ctypes.data, dt.ctypes.data, y0.ctypes.data, rtol, atol,
                                          max_steps, int(chord), y.ctypes.data, steps.ctypes.data, res.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_newton_implicit_euler failed: {rc}")
    return dict(value=y, num_steps=steps, result=res)


def bfgs_rosenbrock(y0, *, rtol, atol, max_steps=256, nthreads=0, fma=False):
    """BFGS (inverse update, backtracking Armijo) on the chained Rosenbrock function (fp64, max_norm), C + OpenMP."""
    lib = load_fma() if fma else load()
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    Bn, n = y0.shape
    y = np.empty_like(y0)
    steps = np.empty(Bn, np.int32)
    res = np.empty(Bn, np.int32)
    rc = lib.oracle_bfgs_rosenbrock(Bn, n, y0.ctypes.data, rtol, atol, max_steps, y.ctypes.data, steps.ctypes.data, res.ctypes.data, nthreads)
    if rc != 0:
        raise RuntimeError(f"oracle_bfgs_rosenbrock failed: {rc}")
    return dict(value=y, num_steps=steps, result=res)
