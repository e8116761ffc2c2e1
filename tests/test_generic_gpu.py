"""The generic path (arbitrary user fn -> torch-computed r, J -> CUDA step kernel) on the REFERENCE'S OWN test
problems (tests/helpers.py of the reference, restated in tests/ref_fixtures.py), and against the oracle. GPU."""
import numpy as np
import pytest
import torch

import optimistix_b200 as optx
from optimistix_b200 import synthetic
from oracle import models as M
from oracle import solvers as S

from . import ref_fixtures as F
from .parity import compare, to_np

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dev(a):
    if a is None:
        return None
    if isinstance(a, (tuple, list)):
        return type(a)(_dev(x) for x in a)
    if isinstance(a, torch.Tensor):
        return a.to(DEV)
    return a


# tests/test_best_so_far.py:27-34 of the reference: LM(1e-6) on y - tanh(y+1) -> 0.96118069 at 1e-5
def test_known_answer_tanh_lm():
    sol = optx.least_squares(lambda y, _: y - torch.tanh(y + 1), optx.LevenbergMarquardt(1e-6, 1e-6),
                             torch.zeros(1, dtype=torch.float64, device=DEV))
    assert np.allclose(to_np(sol.value), 0.96118069, rtol=1e-5, atol=1e-5)
    assert int(sol.result) == optx.RESULTS.successful


# tests/test_best_so_far.py:60-66: GaussNewton(1e-5) on y - 3 -> exactly 3.0
def test_known_answer_gauss_newton_exact_three():
    sol = optx.least_squares(lambda y, _: y - 3.0, optx.GaussNewton(1e-5, 1e-5), torch.zeros(1, dtype=torch.float64, device=DEV))
    assert float(sol.value[0]) == 3.0


# tests/test_least_squares.py:27-53: LM and GN on the least-squares zoo (problems with n <= 16), sum r^2 ~ 0 at 1e-4
@pytest.mark.parametrize("solver_cls", [optx.LevenbergMarquardt, optx.GaussNewton])
def test_least_squares_zoo_rosenbrock2(solver_cls):
    name, fn, minimum, y0, args = F.LEAST_SQUARES[1]
    sol = optx.least_squares(fn, solver_cls(1e-8, 1e-8), torch.as_tensor(y0, device=DEV), _dev(args), max_steps=10_000, throw=False)
    r = fn(sol.value, _dev(args))
    assert np.allclose(float((r * r).sum()), minimum, atol=1e-4, rtol=1e-4)


# tests/test_root_find.py:25-28,32-55: LM as a root finder (through _LstsqToRoot) on the Robertson midpoint problems
@pytest.mark.parametrize("idx", [5, 6, 7])
def test_root_find_via_lm_robertson(idx):
    name, fn, y0, args = F.FIXED_POINT[idx]
    sol = optx.fixed_point(fn, optx.LevenbergMarquardt(1e-8, 1e-8), torch.as_tensor(y0, device=DEV), args, max_steps=10_000, throw=False)
    out = fn(sol.value, args) - sol.value
    assert np.allclose(to_np(out), 0.0, atol=1e-5, rtol=1e-5)
    ref = S.levenberg_marquardt(F.batched_lsq(F.as_root(fn)), y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000,
                                root_near_zero=True, linsolve="gram")
    assert optx.RESULTS.name(int(sol.result)) == S.RESULT_NAMES[ref["result"][0]]
    assert np.allclose(to_np(sol.value), ref["value"][0], rtol=1e-6, atol=1e-12)


# tests/test_solve.py:78-84: NaN start -> RESULTS.nonfinite
def test_generic_nonfinite():
    sol = optx.least_squares(lambda y, _: y**2, optx.LevenbergMarquardt(0.1, 0.1),
                             torch.tensor([float("nan")], dtype=torch.float64, device=DEV), throw=False)
    assert int(sol.result) == optx.RESULTS.nonfinite


# tests/test_least_squares.py:132-151: options jac="bwd" works, anything else raises
def test_generic_jac_option():
    f = lambda y, _: torch.stack([y[0] ** 2 - 2.0, y[0] * y[1] - 1.0, y[1] - 0.5])
    y0 = torch.tensor([1.0, 1.0], dtype=torch.float64, device=DEV)
    a = optx.least_squares(f, optx.LevenbergMarquardt(1e-8, 1e-8), y0, options=dict(jac="fwd"), throw=False)
    b = optx.least_squares(f, optx.LevenbergMarquardt(1e-8, 1e-8), y0, options=dict(jac="bwd"), throw=False)
    assert torch.allclose(a.value, b.value, rtol=1e-9)
    with pytest.raises(ValueError):
        optx.least_squares(f, optx.LevenbergMarquardt(1e-8, 1e-8), y0, options=dict(jac="sideways"))


def test_generic_matches_registered_family_and_oracle():
    """The same exp-decay fits through the generic step kernel, the persistent kernel and the oracle."""
    w = synthetic.exp_decay_batch(512)
    t, d, y0 = (torch.as_tensor(w[k], device=DEV) for k in ("t", "d", "y0"))

    def fn(y, a):
        tt, dd = a
        return y[0] * torch.exp(-y[1] * tt) + y[2] - dd

    gen = optx.least_squares(fn, optx.LevenbergMarquardt(1e-8, 1e-8), y0, (t, d), args_in_dims=(None, 0), throw=False)
    per = optx.least_squares(optx.models.exp_decay, optx.LevenbergMarquardt(1e-8, 1e-8), y0, (t, d), throw=False)
    ref = S.levenberg_marquardt(M.exp_decay, w["y0"], dict(t=w["t"], d=w["d"]), rtol=1e-8, atol=1e-8, linsolve="gram")
    s = compare(gen, ref, 1e-8, 1e-8, "generic step kernel vs oracle")
    assert s["result_match"] == 1.0 and s["within_tol"] == 1.0
    assert abs(s["mean_steps"][0] - s["mean_steps"][1]) / s["mean_steps"][1] < 0.05
    scaled = (gen.value - per.value).abs() / (1e-8 + 1e-8 * per.value.abs())
    assert float(scaled.max()) <= 1.0
    # step-by-step: with max_steps = k both CUDA paths walk the same state machine
    for k in (1, 2, 3, 4, 5):
        g = optx.least_squares(fn, optx.LevenbergMarquardt(1e-8, 1e-8), y0, (t, d), args_in_dims=(None, 0), throw=False, max_steps=k)
        r = S.levenberg_marquardt(M.exp_decay, w["y0"], dict(t=w["t"], d=w["d"]), rtol=1e-8, atol=1e-8, linsolve="gram", max_steps=k)
        safe = r["margin"] > 1e3
        assert np.array_equal(to_np(g.stats["num_steps"])[safe], r["num_steps"][safe])
        assert np.allclose(to_np(g.value)[safe], r["value"][safe], rtol=1e-9, atol=0)
