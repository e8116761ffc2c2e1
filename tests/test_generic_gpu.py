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


# ======================================================================================= generic Newton / Chord / BFGS
def _res_name(sol):
    return optx.RESULTS.name(int(sol.result))


# tests/test_root_find.py:32-55 of the reference: Newton and Chord on the fixed-point zoo, |f(root)| < 1e-5;
# iteration counts and result codes against the oracle (LU with partial pivoting on both sides)
@pytest.mark.parametrize("chord", (False, True), ids=("newton", "chord"))
@pytest.mark.parametrize("name,fn,y0,args", F.FIXED_POINT, ids=[p[0] for p in F.FIXED_POINT])
def test_root_find_zoo_generic(name, fn, y0, args, chord):
    solver = (optx.Chord if chord else optx.Newton)(1e-8, 1e-8)
    sol = optx.fixed_point(fn, solver, torch.as_tensor(y0, device=DEV), _dev(args), max_steps=10_000, throw=False)
    out = fn(sol.value, _dev(args)) - sol.value
    assert np.allclose(to_np(out), 0.0, atol=1e-5, rtol=1e-5)
    ref = S.newton_chord(F.batched_lsq(F.as_root(fn)), y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000, chord=chord)
    assert _res_name(sol) == S.RESULT_NAMES[ref["result"][0]]
    assert int(sol.stats["num_steps"]) == int(ref["num_steps"][0])
    assert np.allclose(to_np(sol.value), ref["value"][0], rtol=1e-9, atol=1e-12)


# tests/test_root_find.py:160-193: clipping to bounds selects the root
@pytest.mark.parametrize("chord,tol", ((False, 1e-4), (True, 1e-2)), ids=("newton", "chord"))
def test_newton_bounded_generic(chord, tol):
    solver = (optx.Chord if chord else optx.Newton)(1e-5, 1e-5)
    inf = float("inf")
    y0 = torch.tensor([0.1, 1.1, 0.1, 0.1, -1.1], dtype=torch.float64, device=DEV)
    sol = optx.root_find(F.newton_bounded, solver, y0, None, options=dict(lower=torch.tensor([-inf, 0.0, -inf, -inf, -inf])), throw=False)
    assert np.allclose(to_np(sol.value), [0.0, 1.0, 0.0, 0.0, -1.0], rtol=tol, atol=tol)
    y0 = torch.tensor([0.1, -1.1, 0.1, 0.1, -1.1], dtype=torch.float64, device=DEV)
    sol = optx.root_find(F.newton_bounded, solver, y0, None, options=dict(upper=torch.tensor([inf, 0.0, inf, inf, inf])), throw=False)
    assert np.allclose(to_np(sol.value), [0.0, -1.0, 0.0, 0.0, -1.0], rtol=tol, atol=tol)


# tests/test_root_find.py:238-245: non-Cauchy (Hairer-Wanner) termination on f(y) = y returns successful
@pytest.mark.parametrize("solver_cls", (optx.Newton, optx.Chord))
def test_newton_chord_small_diff_generic(solver_cls):
    sol = optx.root_find(lambda y, _: y, solver_cls(1e-5, 1e-5, cauchy_termination=False), torch.zeros(1, dtype=torch.float64, device=DEV))
    assert _res_name(sol) == "successful"


def test_newton_generic_batched_matches_registered_family():
    """The implicit-Euler root finds of BASELINE configs[2] through the generic step kernel == the persistent kernel."""
    w = synthetic.implicit_euler_batch(256)
    A, Bm, yp, dt, y0 = (torch.as_tensor(w[k], device=DEV) for k in ("A", "B", "y_prev", "dt", "y0"))

    def fn(y, a):
        A_, B_, yp_, dt_ = a
        return y - yp_ - dt_ * (A_ @ y + y * (B_ @ y))

    gen = optx.root_find(fn, optx.Newton(1e-8, 1e-8), y0, (A, Bm, yp, dt), args_in_dims=(None, None, 0, 0), throw=False)
    per = optx.root_find(optx.models.implicit_euler, optx.Newton(1e-8, 1e-8), y0, (A, Bm, yp, dt), throw=False)
    assert torch.equal(gen.stats["num_steps"], per.stats["num_steps"])
    assert torch.equal(gen.result, per.result)
    assert torch.allclose(gen.value, per.value, rtol=1e-11, atol=1e-13)


# tests/test_minimise.py:30-62: BFGS on the minimisation zoo, f(argmin) ~ min at 1e-4, max_steps = 10_000; vs the oracle
@pytest.mark.parametrize("name,fn,minimum,y0,args", F.MINIMISATION, ids=[p[0] for p in F.MINIMISATION])
def test_bfgs_zoo_generic(name, fn, minimum, y0, args):
    sol = optx.minimise(fn, optx.BFGS(1e-8, 1e-8), torch.as_tensor(y0, device=DEV), _dev(args), max_steps=10_000, throw=False)
    fmin = float(fn(sol.value, _dev(args)))
    assert np.allclose(fmin, minimum, atol=1e-4, rtol=1e-4)
    ref = S.bfgs(F.batched_min(fn), y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000)
    assert _res_name(sol) == S.RESULT_NAMES[ref["result"][0]]
    assert abs(int(sol.stats["num_steps"]) - int(ref["num_steps"][0])) <= max(2, int(0.1 * ref["num_steps"][0]))
    assert np.allclose(to_np(sol.value), ref["value"][0], rtol=1e-5, atol=1e-6)


# tests/test_compat.py:19-24,41-45: Beale optimum [3, 0.5]; two steps are not enough
def test_bfgs_beale_generic():
    y0 = torch.tensor([2.0, 0.0], dtype=torch.float64, device=DEV)
    sol = optx.minimise(F.beale, optx.BFGS(1e-8, 1e-8), y0, (1.5, 2.25, 2.625), max_steps=10_000)
    assert np.allclose(to_np(sol.value), [3.0, 0.5], rtol=1e-5, atol=1e-5)
    sol = optx.minimise(F.beale, optx.BFGS(1e-8, 1e-8), y0, (1.5, 2.25, 2.625), max_steps=2, throw=False)
    assert _res_name(sol) == "nonlinear_max_steps_reached" and int(sol.stats["num_steps"]) == 2
    with pytest.raises(RuntimeError):
        optx.minimise(F.beale, optx.BFGS(1e-8, 1e-8), y0, (1.5, 2.25, 2.625), max_steps=2, throw=True)


def test_bfgs_generic_prefix_parity_with_oracle():
    """Step-by-step: with max_steps = k the step kernel and the oracle hold the same iterate and inverse Hessian."""
    name, fn, minimum, y0, args = F.MINIMISATION[1]  # himmelblau
    for k in (1, 2, 3, 5, 8):
        sol = optx.minimise(fn, optx.BFGS(1e-8, 1e-8), torch.as_tensor(y0, device=DEV), args, max_steps=k, throw=False, return_hessian_inv=True)
        ref = S.bfgs(F.batched_min(fn), y0[None], args, rtol=1e-8, atol=1e-8, max_steps=k)
        assert int(sol.stats["num_steps"]) == int(ref["num_steps"][0])
        assert int(sol.stats["num_accepted_steps"][0]) == int(ref["num_accepted"][0])
        assert np.allclose(to_np(sol.value), ref["value"][0], rtol=1e-12, atol=0)
        assert np.allclose(to_np(sol.state["hessian_inv"][0]), ref["hessian_inv"][0], rtol=1e-10, atol=1e-14)


def test_bfgs_generic_matches_persistent_kernel():
    """Chained Rosenbrock n = 64 (BASELINE configs[3]) through the generic step kernel == the persistent kernel."""
    w = synthetic.rosenbrock_batch(8)
    y0 = torch.as_tensor(w["y0"], device=DEV)

    def fn(y, _):
        return (100.0 * (y[1:] - y[:-1] ** 2) ** 2 + (1.0 - y[:-1]) ** 2).sum()

    gen = optx.minimise(fn, optx.BFGS(1e-8, 1e-8), y0, None, max_steps=40, throw=False, return_hessian_inv=True)
    per = optx.minimise(optx.models.rosenbrock, optx.BFGS(1e-8, 1e-8), y0, None, max_steps=40, throw=False, return_hessian_inv=True)
    assert torch.equal(gen.stats["num_accepted_steps"], per.stats["num_accepted_steps"])
    assert torch.allclose(gen.value, per.value, rtol=1e-9, atol=1e-12)
    assert torch.allclose(gen.state["hessian_inv"], per.state["hessian_inv"], rtol=1e-7, atol=1e-10)


# tests/test_root_find.py:196-213: root_find through a minimiser (BFGS on norm(fn)) and the no-root case
def test_root_via_min_generic():
    y0 = torch.tensor([0.5, -0.3, 0.7], dtype=torch.float64, device=DEV)
    sol = optx.root_find(F.root_via_min, optx.BFGS(1e-8, 1e-8), y0, None, throw=False)
    assert np.allclose(to_np(sol.value), [-0.1, 0.5, 0.0], rtol=1e-5, atol=1e-5)
    bad = optx.root_find(F.bad_root_via_min, optx.BFGS(1e-8, 1e-8), y0, None, throw=False)
    assert _res_name(bad) == "nonlinear_max_steps_reached"
    with pytest.raises(RuntimeError):
        optx.root_find(F.bad_root_via_min, optx.BFGS(1e-8, 1e-8), y0, None, throw=True)


# tests/test_least_squares.py:27-53 with a minimiser as the solver (_least_squares.py:102-114)
def test_least_squares_via_bfgs_generic():
    name, fn, minimum, y0, args = F.LEAST_SQUARES[1]
    sol = optx.least_squares(fn, optx.BFGS(1e-8, 1e-8), torch.as_tensor(y0, device=DEV), _dev(args), max_steps=10_000, throw=False)
    r = fn(sol.value, _dev(args))
    assert np.allclose(float((r * r).sum()), minimum, atol=1e-4, rtol=1e-4)


def test_generic_nonfinite_newton_bfgs():
    nan = torch.tensor([float("nan")], dtype=torch.float64, device=DEV)
    assert _res_name(optx.root_find(lambda y, _: y**2, optx.Newton(0.1, 0.1), nan, throw=False)) == "nonfinite"
    assert _res_name(optx.minimise(lambda y, _: (y**2).sum(), optx.BFGS(0.1, 0.1), nan, throw=False)) == "nonfinite"


# ---- quasi-Newton variants (quasi_newton.py:387-399 BFGS on the Hessian, :487-573 DFP) vs the oracle
@pytest.mark.parametrize("cls,update", ((optx.BFGS, "bfgs"), (optx.DFP, "dfp")))
@pytest.mark.parametrize("use_inverse", (True, False))
@pytest.mark.parametrize("idx", (0, 1, 2))  # bowl (n=20), himmelblau, matyas
def test_quasi_newton_variants_generic(cls, update, use_inverse, idx):
    name, fn, minimum, y0, args = F.MINIMISATION[idx]
    sol = optx.minimise(fn, cls(1e-8, 1e-8, use_inverse=use_inverse), torch.as_tensor(y0, device=DEV), _dev(args), max_steps=10_000,
                        throw=False)
    ref = S.bfgs(F.batched_min(fn), y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000, update=update, use_inverse=use_inverse)
    assert _res_name(sol) == S.RESULT_NAMES[ref["result"][0]]
    if S.RESULT_NAMES[ref["result"][0]] == "successful":
        fmin = float(fn(sol.value, _dev(args)))
        assert np.allclose(fmin, minimum, atol=1e-4, rtol=1e-4)
        assert np.allclose(to_np(sol.value), ref["value"][0], rtol=1e-5, atol=1e-6)
    # the first steps are the same state machine
    for k in (1, 2, 3, 4):
        g = optx.minimise(fn, cls(1e-8, 1e-8, use_inverse=use_inverse), torch.as_tensor(y0, device=DEV), _dev(args), max_steps=k, throw=False)
        r = S.bfgs(F.batched_min(fn), y0[None], args, rtol=1e-8, atol=1e-8, max_steps=k, update=update, use_inverse=use_inverse)
        assert int(g.stats["num_steps"]) == int(r["num_steps"][0])
        assert np.allclose(to_np(g.value), r["value"][0], rtol=1e-9, atol=1e-11)  # matyas: iterates cancel to ~1e-13


# tests/test_compat.py:19-24,27-45 of the reference: the jax.scipy.optimize.minimize drop-in
def test_compat_minimize():
    def beale(x, c1, c2, c3):
        a, b = x[0], x[1]
        return (c1 - a + a * b) ** 2 + (c2 - a + a * b**2) ** 2 + (c3 - a + a * b**3) ** 2

    x0 = torch.tensor([2.0, 0.0], dtype=torch.float64, device=DEV)
    res = optx.compat.minimize(beale, x0, (1.5, 2.25, 2.625), method="bfgs")
    assert bool(res.success) and int(res.status) == 0
    assert np.allclose(to_np(res.x), [3.0, 0.5], rtol=1e-4, atol=1e-4)
    assert np.allclose(float(res.fun), 0.0, atol=1e-8)
    assert res.hess_inv.shape == (2, 2) and res.jac.shape == (2,)
    assert int(res.nit) <= int(res.nfev)
    res2 = optx.compat.minimize(beale, x0, (1.5, 2.25, 2.625), method="BFGS", options=dict(maxiter=2))
    assert not bool(res2.success) and int(res2.status) == 1 and int(res2.nfev) == 2
    with pytest.raises(ValueError):
        optx.compat.minimize(beale, x0, (1.5, 2.25, 2.625), method="l-bfgs")
    with pytest.raises(TypeError):
        optx.compat.minimize(beale, x0, [1.5, 2.25, 2.625], method="bfgs")
    with pytest.raises(ValueError):
        optx.compat.minimize(beale, x0, (1.5, 2.25, 2.625), method="bfgs", options=dict(gtol=1.0))
    with pytest.raises(ValueError):
        optx.compat.minimize(beale, x0[None], (1.5, 2.25, 2.625), method="bfgs")


# ---- shapes and dtypes of the generic paths: batches with per-problem args, fp32, one unknown, empty batch
def test_generic_batched_per_problem_args_and_fp32():
    g = torch.Generator(device=DEV).manual_seed(0)
    B = 37
    a = 1.0 + torch.rand(B, generator=g, device=DEV, dtype=torch.float64)
    # root: y^3 = a (n = 1), minimise: (y0 - a)^2 + 2 (y1 + a)^2 (n = 2), batched over a
    sol = optx.root_find(lambda y, a_: y**3 - a_, optx.Newton(1e-10, 1e-10), torch.ones(B, 1, dtype=torch.float64, device=DEV), a,
                         args_in_dims=0)
    assert torch.allclose(sol.value[:, 0], a ** (1 / 3), rtol=1e-9)
    assert sol.stats["num_steps"].shape == (B,) and bool((sol.result == 0).all())
    sol = optx.minimise(lambda y, a_: (y[0] - a_) ** 2 + 2 * (y[1] + a_) ** 2, optx.BFGS(1e-9, 1e-9),
                        torch.zeros(B, 2, dtype=torch.float64, device=DEV), a, args_in_dims=0)
    assert torch.allclose(sol.value, torch.stack([a, -a], 1), rtol=1e-6, atol=1e-8)
    # fp32 end to end
    a32 = a.float()
    sol = optx.root_find(lambda y, a_: y**3 - a_, optx.Newton(1e-5, 1e-5), torch.ones(B, 1, dtype=torch.float32, device=DEV), a32,
                         args_in_dims=0)
    assert sol.value.dtype == torch.float32 and torch.allclose(sol.value[:, 0], a32 ** (1 / 3), rtol=1e-4)
    sol = optx.minimise(lambda y, a_: (y[0] - a_) ** 2 + 2 * (y[1] + a_) ** 2, optx.DFP(1e-4, 1e-4, use_inverse=False),
                        torch.zeros(B, 2, dtype=torch.float32, device=DEV), a32, args_in_dims=0)
    assert sol.value.dtype == torch.float32 and torch.allclose(sol.value, torch.stack([a32, -a32], 1), rtol=1e-3, atol=1e-4)


def test_generic_empty_batch():
    e = torch.zeros(0, 3, dtype=torch.float64, device=DEV)
    for sol in (optx.root_find(lambda y, _: y - 1.0, optx.Newton(1e-8, 1e-8), e),
                optx.minimise(lambda y, _: (y * y).sum(), optx.BFGS(1e-8, 1e-8), e),
                optx.least_squares(lambda y, _: y - 1.0, optx.LevenbergMarquardt(1e-8, 1e-8), e)):
        assert sol.value.shape == (0, 3) and sol.result.shape == (0,) and sol.stats["num_steps"].shape == (0,)


def test_generic_limits_are_reported():
    with pytest.raises(NotImplementedError):
        optx.root_find(lambda y, _: y, optx.Newton(1e-8, 1e-8), torch.zeros(161, dtype=torch.float64, device=DEV))
    with pytest.raises(NotImplementedError):
        optx.minimise(lambda y, _: (y * y).sum(), optx.BFGS(1e-8, 1e-8, use_inverse=False), torch.zeros(65, dtype=torch.float64, device=DEV))
    with pytest.raises(ValueError):
        optx.minimise(lambda y, _: y, optx.BFGS(1e-8, 1e-8), torch.zeros(2, dtype=torch.float64, device=DEV))  # not a scalar
    sol = optx.minimise(lambda y, _: ((y - 2.0) ** 2).sum(), optx.BFGS(1e-8, 1e-8), torch.zeros(200, dtype=torch.float64, device=DEV))
    assert torch.allclose(sol.value, torch.full((200,), 2.0, dtype=torch.float64, device=DEV), atol=1e-6)  # n > 64: H^-1 in HBM


# ---- has_aux, options["autodiff_mode"], max_steps=None (reference: _minimise.py:37-119, _misc.py:245-258, _iterate.py:194-268)
def test_has_aux_autodiff_mode_and_unbounded_steps():
    y0 = torch.tensor([2.0, 0.0], dtype=torch.float64, device=DEV)
    args = (1.5, 2.25, 2.625)
    a = optx.minimise(F.beale, optx.BFGS(1e-8, 1e-8), y0, args, max_steps=None)  # unbounded, like compat.minimize without maxiter
    b = optx.minimise(F.beale, optx.BFGS(1e-8, 1e-8), y0, args, options=dict(autodiff_mode="fwd"), max_steps=10_000)
    assert np.allclose(to_np(a.value), [3.0, 0.5], rtol=1e-5, atol=1e-5) and a.stats["max_steps"] is None
    assert torch.allclose(a.value, b.value, rtol=1e-8, atol=1e-10)
    # has_aux: fn returns (out, aux); Solution.aux is the auxiliary output at the solution
    sol = optx.minimise(lambda y, c: (F.beale(y, c), y * 2.0), optx.BFGS(1e-8, 1e-8), y0, args, has_aux=True, max_steps=10_000)
    assert torch.allclose(sol.aux, 2.0 * sol.value) and torch.allclose(sol.value, a.value)
    ys = torch.ones(5, 1, dtype=torch.float64, device=DEV)
    cs = torch.arange(2.0, 7.0, dtype=torch.float64, device=DEV)
    sol = optx.root_find(lambda y, c: (y * y - c, {"residual": y * y - c, "c": c}), optx.Newton(1e-10, 1e-10), ys, cs, has_aux=True,
                         args_in_dims=0)
    assert torch.allclose(sol.value[:, 0], cs.sqrt()) and torch.allclose(sol.aux["c"], cs) and float(sol.aux["residual"].abs().max()) < 1e-9
    sol = optx.least_squares(lambda y, _: (y - 3.0, y.sum()), optx.GaussNewton(1e-5, 1e-5), torch.zeros(2, dtype=torch.float64, device=DEV),
                             has_aux=True, max_steps=None)
    assert torch.equal(sol.value, torch.full((2,), 3.0, dtype=torch.float64, device=DEV)) and float(sol.aux) == 6.0
    sol = optx.fixed_point(lambda y, c: (torch.cos(c * y), c), optx.Newton(1e-10, 1e-10), torch.full((3,), 0.5, dtype=torch.float64, device=DEV),
                           torch.tensor(1.0, dtype=torch.float64, device=DEV), has_aux=True)
    assert torch.allclose(torch.cos(sol.value), sol.value, atol=1e-9) and float(sol.aux) == 1.0
    with pytest.raises(NotImplementedError):
        optx.minimise(optx.models.rosenbrock, optx.BFGS(1e-8, 1e-8), torch.zeros(4, dtype=torch.float64, device=DEV), has_aux=True)


# tests/test_least_squares.py:27-53 of the reference: LM and GN on the WHOLE least-squares zoo (n = 20 problems included:
# 16 < n <= 64 runs in the CTA-per-problem step kernel), sum r^2 ~ minimum at 1e-4; iterates against the oracle
@pytest.mark.parametrize("solver_cls,oracle", [(optx.LevenbergMarquardt, S.levenberg_marquardt), (optx.GaussNewton, S.gauss_newton)])
@pytest.mark.parametrize("name,fn,minimum,y0,args", F.LEAST_SQUARES, ids=[p[0] for p in F.LEAST_SQUARES])
def test_least_squares_zoo_all_sizes(name, fn, minimum, y0, args, solver_cls, oracle):
    sol = optx.least_squares(fn, solver_cls(1e-8, 1e-8), torch.as_tensor(y0, device=DEV), _dev(args), max_steps=10_000, throw=False)
    r = fn(sol.value, _dev(args))
    assert np.allclose(float((r * r).sum()), minimum, atol=1e-4, rtol=1e-4)
    ref = oracle(F.batched_lsq(fn), y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000, linsolve="gram")
    assert _res_name(sol) == S.RESULT_NAMES[ref["result"][0]]
    assert abs(int(sol.stats["num_steps"]) - int(ref["num_steps"][0])) <= max(1, int(0.05 * ref["num_steps"][0]))
    assert np.allclose(to_np(sol.value), ref["value"][0], rtol=1e-6, atol=1e-9)


# tests/test_root_find.py:25-28,32-55: LM as a root finder on the 28-dimensional fixed-point problems too
@pytest.mark.parametrize("idx", [0, 1, 2, 3, 4])
def test_root_find_via_lm_n28(idx):
    name, fn, y0, args = F.FIXED_POINT[idx]
    sol = optx.fixed_point(fn, optx.LevenbergMarquardt(1e-8, 1e-8), torch.as_tensor(y0, device=DEV), _dev(args), max_steps=10_000, throw=False)
    out = fn(sol.value, _dev(args)) - sol.value
    assert np.allclose(to_np(out), 0.0, atol=1e-5, rtol=1e-5)
    ref = S.levenberg_marquardt(F.batched_lsq(F.as_root(fn)), y0[None], args, rtol=1e-8, atol=1e-8, max_steps=10_000,
                                root_near_zero=True, linsolve="gram")
    assert _res_name(sol) == S.RESULT_NAMES[ref["result"][0]]
    assert np.allclose(to_np(sol.value), ref["value"][0], rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize("dtype,n,tol", [(torch.float32, 16, 1e-5), (torch.float64, 8, 1e-9), (torch.float64, 16, 1e-9)])
def test_generic_lm_step_kernels_agree_across_the_size_boundary(dtype, n, tol):
    """The same n-parameter problem through one step kernel and, padded with a decoupled (n + 1)-th unknown, through the next:
    fp32 n = 16 (register kernel) vs 17 (CTA kernel); fp64 n = 8 (register kernel) vs 9 (CTA kernel: the packed fp64 Gram of 16
    unknowns does not fit a thread's registers); fp64 16 vs 17 (both in the CTA kernel).  Same result code, same solution."""
    g = torch.Generator(device=DEV).manual_seed(2)
    A = torch.randn(40, n, generator=g, dtype=torch.float64, device=DEV).to(dtype)
    b = torch.randn(40, generator=g, dtype=torch.float64, device=DEV).to(dtype)
    fa = lambda y, _: torch.tanh(A @ y) - 0.3 * b
    fb = lambda y, _: torch.cat([torch.tanh(A @ y[:n]) - 0.3 * b, y[n:] - 2.0])
    sa = optx.least_squares(fa, optx.LevenbergMarquardt(tol, tol), torch.zeros(n, dtype=dtype, device=DEV), None, throw=False, max_steps=1000)
    sb = optx.least_squares(fb, optx.LevenbergMarquardt(tol, tol), torch.zeros(n + 1, dtype=dtype, device=DEV), None, throw=False,
                            max_steps=1000)
    assert _res_name(sa) == _res_name(sb) == "successful"
    loose = 1e-6 if dtype == torch.float64 else 2e-3
    assert torch.allclose(sb.value[:n], sa.value, rtol=loose, atol=loose) and abs(float(sb.value[n]) - 2.0) < loose


# ---- Newton / Chord beyond n = 32: one CTA per problem, the Jacobian in shared memory (the reference has no size limit,
# _solver/newton_chord.py:103-152) -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,dtype", [(33, torch.float64), (64, torch.float64), (100, torch.float64), (160, torch.float64), (48, torch.float32)])
def test_newton_chord_large_n_vs_oracle(n, dtype):
    """Implicit-Euler steps of a stiff quadratic ODE with n unknowns through the generic step kernel against the oracle: the LU
    with partial pivoting is the oracle's (first maximal pivot), so iteration counts and result codes are identical and the
    roots agree to rounding."""
    B = 24
    w = synthetic.implicit_euler_batch(B, n=n, seed=7)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    tol = 1e-8 if dtype == torch.float64 else 1e-4
    A, Bm = torch.as_tensor(w["A"].astype(npdt), device=DEV), torch.as_tensor(w["B"].astype(npdt), device=DEV)
    yp, dt = torch.as_tensor(w["y_prev"].astype(npdt), device=DEV), torch.as_tensor(w["dt"].astype(npdt), device=DEV)

    def F_(y, a):
        return y - a["y_prev"] - a["dt"] * (A @ y + y * (Bm @ y))

    args, in_dims = {"y_prev": yp, "dt": dt}, {"y_prev": 0, "dt": 0}
    oargs = dict(A=w["A"].astype(npdt), B=w["B"].astype(npdt), y_prev=w["y_prev"].astype(npdt), dt=w["dt"].astype(npdt))
    for chord in (False, True):
        solver = (optx.Chord if chord else optx.Newton)(tol, tol)
        sol = optx.root_find(F_, solver, yp.clone(), args, args_in_dims=in_dims, throw=False)
        ref = S.newton_chord(M.implicit_euler, w["y_prev"].astype(npdt), oargs, rtol=tol, atol=tol, chord=chord)
        assert (to_np(sol.result) == ref["result"]).all()
        if dtype == torch.float64:
            assert (to_np(sol.stats["num_steps"]) == ref["num_steps"]).all()
            assert np.allclose(to_np(sol.value), ref["value"], rtol=1e-9, atol=1e-12)
        else:
            assert np.abs(to_np(sol.stats["num_steps"]) - ref["num_steps"]).max() <= 1
            assert np.allclose(to_np(sol.value), ref["value"], rtol=2e-4, atol=2e-5)
        assert (to_np(sol.result) == optx.RESULTS.successful).all()


def test_lu_solve_large_n_and_adjoint():
    """b200opt_lu_solve for 32 < n (the linear solve of ImplicitAdjoint): against torch.linalg.solve, singular systems flagged."""
    from optimistix_b200._adjoint import lu_solve

    gen = torch.Generator(device=DEV).manual_seed(1)
    for n in (40, 96, 160):
        Amat = torch.randn(7, n, n, dtype=torch.float64, device=DEV, generator=gen) + n ** 0.5 * torch.eye(n, dtype=torch.float64, device=DEV)
        bvec = torch.randn(7, n, dtype=torch.float64, device=DEV, generator=gen)
        x, status = lu_solve(Amat, bvec)
        assert (status == 0).all()
        assert torch.allclose(x, torch.linalg.solve(Amat, bvec), rtol=1e-9, atol=1e-11)
    Amat[3] = 0.0
    x, status = lu_solve(Amat, bvec)
    assert int(status[3]) == optx.RESULTS.singular and (status[[0, 1, 2, 4, 5, 6]] == 0).all()
