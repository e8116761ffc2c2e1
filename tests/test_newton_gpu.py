"""Parity of the CUDA Newton/Chord kernel (through the C ABI) against the oracle and the golden vectors. GPU."""
import os

import numpy as np
import pytest
import torch

import optimistix_b200 as optx
from optimistix_b200 import synthetic
from oracle import models as M
from oracle import solvers as S

from .parity import compare, to_np

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _cuda(x):
    return torch.as_tensor(x, device=DEV)


def _solve(w, solver, **kw):
    args = (_cuda(w["A"]), _cuda(w["B"]), _cuda(w["y_prev"]), _cuda(w["dt"]))
    return optx.root_find(optx.models.implicit_euler, solver, _cuda(w["y0"]), args, throw=False, **kw)


@pytest.mark.parametrize("chord", [False, True], ids=["newton", "chord"])
def test_newton_vs_oracle(chord):
    """configs[2] shape (n=32, fp64, rtol=atol=1e-8, Cauchy termination) at an oracle-sized batch."""
    w = synthetic.implicit_euler_batch(4096)
    solver = (optx.Chord if chord else optx.Newton)(1e-8, 1e-8)
    sol = _solve(w, solver)
    ref = S.newton_chord(M.implicit_euler, w["y0"], w, rtol=1e-8, atol=1e-8, chord=chord)
    s = compare(sol, ref, 1e-8, 1e-8, "chord" if chord else "newton")
    assert s["result_match"] == 1.0
    assert s["steps_match"] >= 0.999          # well-conditioned: identical iteration counts
    assert s["rel_same_max"] <= 1e-10         # north_star: rtol 1e-10 fp64 on solution values
    # the root is a root: |F(y)| tiny
    F = optx.models.implicit_euler(sol.value, (_cuda(w["A"]), _cuda(w["B"]), _cuda(w["y_prev"]), _cuda(w["dt"])))
    assert float(F.abs().max()) < 1e-10


@pytest.mark.parametrize("name,chord", [("newton_implicit_euler_f64", False), ("chord_implicit_euler_f64", True)])
def test_newton_golden(name, chord):
    g = np.load(os.path.join(G, name + ".npz"))
    w = synthetic.implicit_euler_batch(512)
    sol = _solve(w, (optx.Chord if chord else optx.Newton)(1e-8, 1e-8))
    s = compare(sol, g, 1e-8, 1e-8, name)
    assert s["result_match"] == 1.0 and s["steps_match"] == 1.0 and s["rel_same_max"] <= 1e-10


def test_newton_fp32():
    w = synthetic.implicit_euler_batch(1024, dtype=np.float32)
    sol = _solve(w, optx.Newton(1e-4, 1e-4))
    ref = S.newton_chord(M.implicit_euler, w["y0"], w, rtol=1e-4, atol=1e-4)
    s = compare(sol, ref, 1e-4, 1e-4, "newton f32")
    assert s["result_match"] >= 0.999 and s["within_tol"] >= 0.999 and s["steps_match"] >= 0.95


@pytest.mark.parametrize("chord", [False, True], ids=["newton", "chord"])
def test_hairer_wanner_termination(chord):
    """cauchy_termination=False (newton_chord.py:179-192), the mode diffrax uses."""
    w = synthetic.implicit_euler_batch(512)
    cls = optx.Chord if chord else optx.Newton
    sol = _solve(w, cls(1e-8, 1e-8, cauchy_termination=False))
    ref = S.newton_chord(M.implicit_euler, w["y0"], w, rtol=1e-8, atol=1e-8, chord=chord, cauchy=False)
    s = compare(sol, ref, 1e-8, 1e-8, "hairer-wanner")
    assert s["result_match"] == 1.0 and s["steps_match"] >= 0.99 and s["within_tol"] == 1.0


def test_newton_stiff_large_dt_and_max_steps():
    """Larger steps of the stiff system: more iterations, some hit max_steps; codes and counts must match."""
    w = synthetic.implicit_euler_batch(512)
    w["dt"] = w["dt"] * 300.0
    for max_steps in (2, 256):
        sol = _solve(w, optx.Newton(1e-8, 1e-8), max_steps=max_steps)
        ref = S.newton_chord(M.implicit_euler, w["y0"], w, rtol=1e-8, atol=1e-8, max_steps=max_steps)
        s = compare(sol, ref, 1e-8, 1e-8, f"stiff max_steps={max_steps}")
        assert s["result_match"] >= 0.995 and s["steps_match"] >= 0.99


def test_newton_bounds_clip():
    """options lower/upper (newton_chord.py:112-135): iterates are clipped to the box."""
    w = synthetic.implicit_euler_batch(256)
    lower = np.full(32, 0.05)
    upper = np.full(32, 0.9)
    args = (_cuda(w["A"]), _cuda(w["B"]), _cuda(w["y_prev"]), _cuda(w["dt"]))
    sol = optx.root_find(optx.models.implicit_euler, optx.Newton(1e-8, 1e-8), _cuda(w["y0"]), args,
                         options=dict(lower=lower, upper=upper), throw=False, max_steps=64)
    ref = S.newton_chord(M.implicit_euler, w["y0"], w, rtol=1e-8, atol=1e-8, lower=lower[None], upper=upper[None], max_steps=64)
    s = compare(sol, ref, 1e-8, 1e-8, "bounded")
    v = to_np(sol.value)
    assert v.min() >= 0.05 and v.max() <= 0.9
    assert s["result_match"] >= 0.99 and s["steps_match"] >= 0.99 and s["within_tol"] >= 0.99


def test_newton_nonfinite_and_singular():
    w = synthetic.implicit_euler_batch(8)
    w["y0"][2, 5] = np.nan
    sol = _solve(w, optx.Newton(1e-8, 1e-8))
    ref = S.newton_chord(M.implicit_euler, w["y0"], w, rtol=1e-8, atol=1e-8)
    assert np.array_equal(to_np(sol.result), ref["result"]) and to_np(sol.result)[2] == optx.RESULTS.nonfinite
    assert np.array_equal(to_np(sol.stats["num_steps"]), ref["num_steps"])
    # dt = inf makes J non-finite -> linear solve reports singular (lineax semantics)
    w = synthetic.implicit_euler_batch(8)
    w["dt"][1] = np.inf
    sol = _solve(w, optx.Newton(1e-8, 1e-8))
    ref = S.newton_chord(M.implicit_euler, w["y0"], w, rtol=1e-8, atol=1e-8)
    assert np.array_equal(to_np(sol.result), ref["result"]) and to_np(sol.result)[1] == optx.RESULTS.singular


def test_newton_host_entry_point():
    w = synthetic.implicit_euler_batch(777)
    dev = _solve(w, optx.Newton(1e-8, 1e-8))
    host = optx.root_find(optx.models.implicit_euler, optx.Newton(1e-8, 1e-8), w["y0"], (w["A"], w["B"], w["y_prev"], w["dt"]),
                          throw=False, chunk=200)
    assert np.array_equal(host.value, to_np(dev.value)) and np.array_equal(host.stats["num_steps"], to_np(dev.stats["num_steps"]))
    # the Python entry point hands (y_prev, dt) to b200opt_root_find_host_parts; the packed-row entry point gives the same bits
    import ctypes as C

    from optimistix_b200 import _lib

    lib = _lib.load()
    n, B = 32, 777
    rows = np.zeros((B, n + 2))
    rows[:, :n], rows[:, n] = w["y_prev"], w["dt"]
    shared = np.concatenate([w["A"].reshape(-1), w["B"].reshape(-1)])
    y = np.empty((B, n))
    steps, res = np.empty(B, np.int32), np.empty(B, np.int32)
    opts = _lib.make_options(1e-8, 1e-8, "max", 256)
    _lib.check(lib.b200opt_root_find_host(0, _lib.SOLVER_NEWTON, optx.models.implicit_euler.model_id, _lib.F64, B, n, w["y0"].ctypes.data,
                                          shared.ctypes.data, shared.size, rows.ctypes.data, n + 2, C.byref(opts), y.ctypes.data,
                                          steps.ctypes.data, res.ctypes.data, 300), "root_find_host")
    assert np.array_equal(y, host.value) and np.array_equal(steps, host.stats["num_steps"]) and np.array_equal(res, host.result)
    # single problem and a pageable chunk layout that does not divide the batch
    one = optx.root_find(optx.models.implicit_euler, optx.Newton(1e-8, 1e-8), w["y0"][5], (w["A"], w["B"], w["y_prev"][5], w["dt"][5]), throw=False)
    assert np.array_equal(one.value, host.value[5])
