"""ImplicitAdjoint over the CUDA solves (optimistix/_ad.py:26-118 of the reference): gradients of a solution with
respect to `args` by the implicit function theorem, against closed forms and finite differences of the solve itself
(the reference's own check: tests/test_root_find.py:58-130, test_least_squares.py:56-129, test_minimise.py:65-129
compare the JVP with finite differences / the unrolled solve).  GPU."""
import pytest
import torch

import optimistix_b200 as optx

from . import ref_fixtures as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
f64 = dict(dtype=torch.float64, device=DEV)


def test_lu_solve_kernel_matches_torch():
    from optimistix_b200._adjoint import lu_solve

    g = torch.Generator(device=DEV).manual_seed(0)
    for n in (1, 3, 8, 13, 16, 28, 32):
        for dt, tol in ((torch.float64, 1e-10), (torch.float32, 2e-3)):
            A = torch.randn(257, n, n, generator=g, device=DEV, dtype=dt) + 0.5 * n**0.5 * torch.eye(n, device=DEV, dtype=dt)
            b = torch.randn(257, n, generator=g, device=DEV, dtype=dt)
            x, status = lu_solve(A, b)
            ref = torch.linalg.solve(A.double(), b.double())
            assert int(status.abs().sum()) == 0
            assert torch.allclose(x.double(), ref, rtol=tol, atol=tol)
    x, status = lu_solve(torch.zeros(2, 4, 4, **f64), torch.ones(2, 4, **f64))
    assert (status == optx.RESULTS.singular).all()


def test_root_find_gradient_closed_form():
    """y^2 = a  =>  y* = sqrt(a), dy*/da = 1 / (2 sqrt(a)); Newton, batched over a."""
    a = torch.tensor([2.0, 3.0, 5.0, 7.0], requires_grad=True, **f64)
    sol = optx.root_find(lambda y, a_: y * y - a_, optx.Newton(1e-10, 1e-10), torch.ones(4, 1, **f64), a, args_in_dims=0)
    assert torch.allclose(sol.value[:, 0], a.detach().sqrt(), rtol=1e-9)
    sol.value.sum().backward()
    assert torch.allclose(a.grad, 0.5 / a.detach().sqrt(), rtol=1e-8)


@pytest.mark.parametrize("solver", ("newton", "chord", "lm", "bfgs"))
def test_fixed_point_gradient_vs_finite_differences(solver):
    """d/dc of sum(y*(c)) for the reference's `_exponential` fixed point y = exp(-c y) (helpers.py:674-676)."""
    mk = {"newton": lambda: optx.Newton(1e-10, 1e-10), "chord": lambda: optx.Chord(1e-10, 1e-10),
          "lm": lambda: optx.LevenbergMarquardt(1e-10, 1e-10), "bfgs": lambda: optx.BFGS(1e-10, 1e-10)}[solver]
    y0 = torch.full((8,), 0.5, **f64)

    def solve(c):
        return optx.fixed_point(F._exponential, mk(), y0, c, max_steps=10_000, throw=False).value

    c = torch.tensor(1.0, requires_grad=True, **f64)
    y = solve(c)
    (y * torch.arange(1, 9, **f64)).sum().backward()
    h = 1e-6
    with torch.no_grad():
        fd = ((solve(torch.tensor(1.0 + h, **f64)) - solve(torch.tensor(1.0 - h, **f64))) / (2 * h) * torch.arange(1, 9, **f64)).sum()
    # closed form: y = exp(-c y)  =>  dy/dc = -y^2 / (1 + c y)
    yy = y.detach()
    exact = ((-yy * yy / (1 + yy)) * torch.arange(1, 9, **f64)).sum()
    assert torch.allclose(c.grad, exact, rtol=1e-7)
    assert torch.allclose(c.grad, fd, rtol=1e-4)


def test_least_squares_gradient_exp_decay_batched():
    """Curve fits y*(d): the gradient w.r.t. the DATA of every fit (per-problem args), against finite differences."""
    from optimistix_b200 import synthetic

    w = synthetic.exp_decay_batch(16)
    t = torch.as_tensor(w["t"], device=DEV)
    d = torch.as_tensor(w["d"], device=DEV).clone().requires_grad_(True)
    y0 = torch.as_tensor(w["y0"], device=DEV)

    def fn(y, a):
        tt, dd = a
        return y[0] * torch.exp(-y[1] * tt) + y[2] - dd

    def solve(dd):
        return optx.least_squares(fn, optx.LevenbergMarquardt(1e-12, 1e-12), y0, (t, dd), args_in_dims=(None, 0), throw=False).value

    y = solve(d)
    y[:, 1].sum().backward()  # the decay rates
    g = d.grad.clone()
    assert g.shape == d.shape and torch.isfinite(g).all()
    h = 1e-5
    with torch.no_grad():
        for (b, i) in ((0, 3), (5, 40), (15, 63)):
            dp, dm = d.detach().clone(), d.detach().clone()
            dp[b, i] += h
            dm[b, i] -= h
            fd = (solve(dp)[b, 1] - solve(dm)[b, 1]) / (2 * h)
            assert abs(float(g[b, i]) - float(fd)) <= 1e-4 * max(1.0, abs(float(fd))), (b, i, float(g[b, i]), float(fd))
            # other problems' data do not influence this fit
        assert float(g.abs().sum()) > 0


def test_minimise_gradient_quadratic_bowl():
    """min_y 0.5 y^T A y - b^T y  =>  y* = A^-1 b;  d sum(y*) / d b = A^-T 1."""
    g = torch.Generator(device=DEV).manual_seed(1)
    M = torch.randn(6, 6, generator=g, **f64)
    A = M @ M.T + 6 * torch.eye(6, **f64)
    b = torch.randn(6, generator=g, **f64).requires_grad_(True)
    sol = optx.minimise(lambda y, a: 0.5 * y @ a[0] @ y - a[1] @ y, optx.BFGS(1e-10, 1e-10), torch.zeros(6, **f64), (A, b), max_steps=1000)
    assert torch.allclose(sol.value, torch.linalg.solve(A, b.detach()), rtol=1e-7, atol=1e-9)
    sol.value.sum().backward()
    assert torch.allclose(b.grad, torch.linalg.solve(A.T, torch.ones(6, **f64)), rtol=1e-8)


def test_recursive_checkpoint_adjoint_is_refused():
    a = torch.tensor([2.0], requires_grad=True, **f64)
    with pytest.raises(NotImplementedError):
        optx.root_find(lambda y, a_: y * y - a_, optx.Newton(1e-10, 1e-10), torch.ones(1, **f64), a, adjoint=optx.RecursiveCheckpointAdjoint())


def test_registered_family_gradient_matches_generic_path():
    """exp_decay through the persistent kernel: d(decay rate)/d(data) by ImplicitAdjoint == the generic path's gradient."""
    from optimistix_b200 import synthetic

    w = synthetic.exp_decay_batch(32)
    t = torch.as_tensor(w["t"], device=DEV)
    y0 = torch.as_tensor(w["y0"], device=DEV)
    d1 = torch.as_tensor(w["d"], device=DEV).clone().requires_grad_(True)
    d2 = d1.detach().clone().requires_grad_(True)
    reg = optx.least_squares(optx.models.exp_decay, optx.LevenbergMarquardt(1e-12, 1e-12), y0, (t, d1), throw=False)
    (reg.value[:, 1] * torch.arange(32, **f64)).sum().backward()

    def fn(y, a):
        tt, dd = a
        return y[0] * torch.exp(-y[1] * tt) + y[2] - dd

    gen = optx.least_squares(fn, optx.LevenbergMarquardt(1e-12, 1e-12), y0, (t, d2), args_in_dims=(None, 0), throw=False)
    (gen.value[:, 1] * torch.arange(32, **f64)).sum().backward()
    assert torch.allclose(reg.value, gen.value, rtol=1e-8, atol=1e-10)
    assert torch.allclose(d1.grad, d2.grad, rtol=1e-6, atol=1e-9)
    assert float(d1.grad.abs().sum()) > 0


def test_registered_root_family_gradient_vs_finite_differences():
    """implicit_euler (BASELINE configs[2] family): d y_new / d dt through the persistent Newton kernel."""
    from optimistix_b200 import synthetic

    w = synthetic.implicit_euler_batch(8)
    A, Bm, yp, y0 = (torch.as_tensor(w[k], device=DEV) for k in ("A", "B", "y_prev", "y0"))
    dt = torch.as_tensor(w["dt"], device=DEV).clone().requires_grad_(True)
    sol = optx.root_find(optx.models.implicit_euler, optx.Newton(1e-12, 1e-12), y0, (A, Bm, yp, dt), throw=False)
    sol.value.sum().backward()
    h = 1e-7
    with torch.no_grad():
        up = optx.root_find(optx.models.implicit_euler, optx.Newton(1e-12, 1e-12), y0, (A, Bm, yp, dt + h), throw=False).value
        dn = optx.root_find(optx.models.implicit_euler, optx.Newton(1e-12, 1e-12), y0, (A, Bm, yp, dt - h), throw=False).value
    fd = ((up - dn) / (2 * h)).sum(1)
    assert torch.allclose(dt.grad, fd, rtol=1e-4, atol=1e-6)
