"""Pin the oracle against every known answer the reference's own tests hold for the hot path.

Each test names the reference test it restates (paths relative to /root/reference/tests).
All of these are tolerance checks - the reference never pins iteration counts.
Runs on CPU (not gpu-marked).
"""
import numpy as np
import pytest
import torch

from oracle import solvers as S

from . import ref_fixtures as F

LINSOLVES = ("qr", "gram")


# ---- test_least_squares.py:27-53 : LM and GN x zoo, max_steps=10_000, sum r^2 ~ minimum at 1e-4
@pytest.mark.parametrize("linsolve", LINSOLVES)
@pytest.mark.parametrize("search", ("trust_region", "learning_rate"))
@pytest.mark.parametrize("name,fn,minimum,y0,args", F.LEAST_SQUARES, ids=[p[0] for p in F.LEAST_SQUARES])
def test_least_squares_zoo(name, fn, minimum, y0, args, search, linsolve):
    if search == "learning_rate" and linsolve == "gram" and name == "diagonal_quadratic_bowl":
        pytest.skip("undamped normal equations on a rank-deficient-at-0 Jacobian (reference uses QR here)")
    model = F.batched_lsq(fn)
    out = S.gauss_newton_family(model, y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000, search=search, linsolve=linsolve)
    r, _ = model(out["value"], args)
    assert np.allclose(np.sum(r**2), minimum, atol=1e-4, rtol=1e-4)


# ---- test_best_so_far.py:27-34 : y - tanh(y+1), LM(1e-6) -> 0.96118069 at 1e-5
@pytest.mark.parametrize("linsolve", LINSOLVES)
def test_lm_tanh_known_answer(linsolve):
    model = F.batched_lsq(lambda y, _: y - torch.tanh(y + 1))
    out = S.levenberg_marquardt(model, np.zeros((1, 1)), None, rtol=1e-6, atol=1e-6, linsolve=linsolve)
    assert S.RESULT_NAMES[out["result"][0]] == "successful"
    assert np.allclose(out["value"], 0.96118069, rtol=1e-5, atol=1e-5)


# ---- test_best_so_far.py:37-44 : 0.5 (y - tanh(y+1))^2, BFGS(1e-6) -> 0.96118069
def test_bfgs_tanh_known_answer():
    model = F.batched_min(lambda y, _: 0.5 * torch.sum((y - torch.tanh(y + 1)) ** 2))
    out = S.bfgs(model, np.zeros((1, 1)), None, rtol=1e-6, atol=1e-6)
    assert S.RESULT_NAMES[out["result"][0]] == "successful"
    assert np.allclose(out["value"], 0.96118069, rtol=1e-5, atol=1e-5)


# ---- test_best_so_far.py:48-57 : (y-3)^2 with BFGS(1e-5) -> exactly 3.0
def test_bfgs_exact_three():
    model = F.batched_min(lambda y, _: torch.sum((y - 3.0) ** 2))
    out = S.bfgs(model, np.zeros((1, 1)), None, rtol=1e-5, atol=1e-5)
    assert out["value"][0, 0] == 3.0


# ---- test_best_so_far.py:60-66 : y-3 with GaussNewton(1e-5) -> exactly 3.0
@pytest.mark.parametrize("linsolve", LINSOLVES)
def test_gauss_newton_exact_three(linsolve):
    model = F.batched_lsq(lambda y, _: y - 3.0)
    out = S.gauss_newton(model, np.zeros((1, 1)), None, rtol=1e-5, atol=1e-5, linsolve=linsolve)
    assert out["value"][0, 0] == 3.0
    assert S.RESULT_NAMES[out["result"][0]] == "successful"


# ---- test_minimise.py:30-62 : BFGS x zoo, f(argmin) ~ min at 1e-4, max_steps=10_000
@pytest.mark.parametrize("name,fn,minimum,y0,args", F.MINIMISATION, ids=[p[0] for p in F.MINIMISATION])
def test_bfgs_zoo(name, fn, minimum, y0, args):
    model = F.batched_min(fn)
    out = S.bfgs(model, y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000)
    fmin, _ = model(out["value"], args)
    assert np.allclose(fmin, minimum, atol=1e-4, rtol=1e-4)


# ---- test_compat.py:19-24 : Beale optimum [3.0, 0.5], f = 0 (BFGS)
def test_bfgs_beale_known_answer():
    model = F.batched_min(F.beale)
    out = S.bfgs(model, np.array([[2.0, 0.0]]), (1.5, 2.25, 2.625), rtol=1e-8, atol=1e-8, max_steps=10_000)
    assert np.allclose(out["value"], [[3.0, 0.5]], rtol=1e-5, atol=1e-5)


# ---- test_compat.py:41-45 : maxiter=2 -> not successful (status 1 == max steps)
def test_bfgs_maxiter():
    model = F.batched_min(F.beale)
    out = S.bfgs(model, np.array([[2.0, 0.0]]), (1.5, 2.25, 2.625), rtol=1e-8, atol=1e-8, max_steps=2)
    assert S.RESULT_NAMES[out["result"][0]] == "nonlinear_max_steps_reached"
    assert out["num_steps"][0] == 2


# ---- test_root_find.py:32-55 : Newton and Chord on the fixed-point zoo, |f(root)| < 1e-5
@pytest.mark.parametrize("chord", (False, True), ids=("newton", "chord"))
@pytest.mark.parametrize("name,fn,y0,args", F.FIXED_POINT, ids=[p[0] for p in F.FIXED_POINT])
def test_root_find_zoo(name, fn, y0, args, chord):
    model = F.batched_lsq(F.as_root(fn))  # (F, J) of the root function
    out = S.newton_chord(model, y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000, chord=chord)
    fval, _ = model(out["value"], args)
    assert np.allclose(fval, 0.0, atol=1e-5, rtol=1e-5)


# ---- test_root_find.py (LM in _solvers, :25-28) : LM as a root finder through _LstsqToRoot
@pytest.mark.parametrize("name,fn,y0,args", F.FIXED_POINT[:5], ids=[p[0] for p in F.FIXED_POINT[:5]])
def test_root_find_via_lm(name, fn, y0, args):
    model = F.batched_lsq(F.as_root(fn))
    out = S.levenberg_marquardt(model, y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000, root_near_zero=True)
    fval, _ = model(out["value"], args)
    assert np.allclose(fval, 0.0, atol=1e-5, rtol=1e-5)


# ---- test_root_find.py:160-193 : clipping to bounds selects the root
@pytest.mark.parametrize("chord,tol", ((False, 1e-4), (True, 1e-2)), ids=("newton", "chord"))
def test_newton_bounded(chord, tol):
    model = F.batched_lsq(F.newton_bounded)
    inf = np.inf
    lower = np.array([[-inf, 0.0, -inf, -inf, -inf]])
    out = S.newton_chord(model, np.array([[0.1, 1.1, 0.1, 0.1, -1.1]]), None, rtol=1e-5, atol=1e-5, chord=chord, lower=lower)
    assert np.allclose(out["value"], [[0.0, 1.0, 0.0, 0.0, -1.0]], rtol=tol, atol=tol)
    upper = np.array([[inf, 0.0, inf, inf, inf]])
    out = S.newton_chord(model, np.array([[0.1, -1.1, 0.1, 0.1, -1.1]]), None, rtol=1e-5, atol=1e-5, chord=chord, upper=upper)
    assert np.allclose(out["value"], [[0.0, -1.0, 0.0, 0.0, -1.0]], rtol=tol, atol=tol)


# ---- test_root_find.py:196-203 : root via BFGS on norm(root) -> (-0.1, [0.5, 0.0])
def test_root_via_min():
    fmin = lambda y, _: torch.max(torch.abs(F.root_via_min(y, None)))  # _root_find.py:34-36 with max_norm
    model = F.batched_min(fmin)
    resid = lambda Y: F.batched_lsq(F.root_via_min)(Y, None)[0]
    out = S.bfgs(model, np.array([[0.5, -0.3, 0.7]]), None, rtol=1e-8, atol=1e-8, near_zero_residual=resid)
    assert np.allclose(out["value"], [[-0.1, 0.5, 0.0]], rtol=1e-5, atol=1e-5)  # tree_allclose defaults


# ---- test_root_find.py:206-213 : no root -> nonlinear_max_steps_reached
def test_bad_root_via_min():
    fmin = lambda y, _: torch.max(torch.abs(F.bad_root_via_min(y, None)))
    model = F.batched_min(fmin)
    resid = lambda Y: F.batched_lsq(F.bad_root_via_min)(Y, None)[0]
    out = S.bfgs(model, np.array([[0.5, -0.3, 0.7]]), None, rtol=1e-8, atol=1e-8, near_zero_residual=resid)
    assert S.RESULT_NAMES[out["result"][0]] == "nonlinear_max_steps_reached"


# ---- test_root_find.py:238-245 : non-Cauchy (Hairer-Wanner) mode on f(y)=y returns successful
@pytest.mark.parametrize("chord", (False, True), ids=("newton", "chord"))
def test_newton_chord_small_diff(chord):
    model = F.batched_lsq(lambda y, _: y)
    out = S.newton_chord(model, np.zeros((1, 1)), None, rtol=1e-5, atol=1e-5, chord=chord, cauchy=False)
    assert S.RESULT_NAMES[out["result"][0]] == "successful"


# ---- test_solve.py:78-84 : NaN start -> RESULTS.nonfinite (driver-level, every solver)
def test_nonfinite_input():
    nan = np.array([[np.nan]])
    out = S.bfgs(F.batched_min(lambda y, _: torch.sum(y**2)), nan, None, rtol=0.1, atol=0.1)
    assert S.RESULT_NAMES[out["result"][0]] == "nonfinite" and out["num_steps"][0] == 0
    out = S.levenberg_marquardt(F.batched_lsq(lambda y, _: y**2), nan, None, rtol=0.1, atol=0.1)
    assert S.RESULT_NAMES[out["result"][0]] == "nonfinite"
    out = S.newton_chord(F.batched_lsq(lambda y, _: y**2), nan, None, rtol=0.1, atol=0.1)
    assert S.RESULT_NAMES[out["result"][0]] == "nonfinite"


# ---- vmap semantics: a batched solve equals the per-problem solves
def test_batched_equals_unbatched():
    from optimistix_b200 import synthetic
    from oracle import models

    w = synthetic.exp_decay_batch(16)
    args = dict(t=w["t"], d=w["d"])
    full = S.levenberg_marquardt(models.exp_decay, w["y0"], args, rtol=1e-8, atol=1e-8)
    for i in (0, 7, 15):
        one = S.levenberg_marquardt(models.exp_decay, w["y0"][i : i + 1], dict(t=w["t"], d=w["d"][i : i + 1]), rtol=1e-8, atol=1e-8)
        assert one["num_steps"][0] == full["num_steps"][i]
        assert np.array_equal(one["value"][0], full["value"][i])


# ---- helpers.py:250-254 of the reference: BFGS(use_inverse=False), DFP(use_inverse=True/False) in the minimiser list of
# test_minimise.py:30-62 (f(argmin) ~ min at 1e-4)
@pytest.mark.parametrize("update,use_inverse", (("bfgs", False), ("dfp", True), ("dfp", False)))
@pytest.mark.parametrize("name,fn,minimum,y0,args", F.MINIMISATION[:3], ids=[p[0] for p in F.MINIMISATION[:3]])
def test_quasi_newton_variants_zoo(name, fn, minimum, y0, args, update, use_inverse):
    model = F.batched_min(fn)
    out = S.bfgs(model, y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000, update=update, use_inverse=use_inverse)
    fmin, _ = model(out["value"], args)
    assert S.RESULT_NAMES[out["result"][0]] == "successful"
    assert np.allclose(fmin, minimum, atol=1e-4, rtol=1e-4)
