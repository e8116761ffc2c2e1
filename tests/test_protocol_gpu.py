"""The reference's solver protocol - init / step / terminate / postprocess (optimistix/_iterate.py:33-184) - driven by
hand, as in docs/examples/interactive.ipynb, against the entry points that run the same CUDA step kernels in a loop."""
import pytest
import torch

import optimistix_b200 as optx
from optimistix_b200 import FunctionInfo, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _interactive(solver, fn, y, args, options=None, max_steps=256):
    """The loop of the reference's interactive example, batch-aware (`done` has one entry per problem)."""
    tags = frozenset()
    state = solver.init(fn, y, args, options or {}, None, None, tags)
    done, result = solver.terminate(fn, y, args, options or {}, state, tags)
    aux = None
    steps = 0
    while not bool(torch.as_tensor(done).all()) and steps < max_steps:
        y, state, aux = solver.step(fn, y, args, options or {}, state, tags)
        done, result = solver.terminate(fn, y, args, options or {}, state, tags)
        steps += 1
    y, aux, stats = solver.postprocess(fn, y, aux, args, options or {}, state, tags, result)
    return y, aux, state, result, steps


def test_interactive_example_root_find():
    """docs/examples/interactive.ipynb with Newton in place of Bisection: y - tanh(y + 1) = 0, fn returns (out, aux)."""
    def fn(y, args):
        return y - torch.tanh(y + 1), {"evaluated_at": y}

    solver = optx.Newton(rtol=1e-8, atol=1e-8)
    y0 = torch.zeros(1, dtype=torch.float64, device=DEV)
    y, aux, state, result, steps = _interactive(solver, fn, y0, None)
    assert int(result) == optx.RESULTS.successful
    assert abs(float(y - torch.tanh(y + 1))) < 1e-8
    assert torch.equal(aux["evaluated_at"], y)
    sol = optx.root_find(lambda y, a: fn(y, a)[0], solver, y0, None, throw=False)
    assert torch.equal(sol.value, y) and int(sol.stats["num_steps"]) == steps == int(state.num_steps)


@pytest.mark.parametrize("cls", [optx.LevenbergMarquardt, optx.GaussNewton])
def test_least_squares_stepping_equals_entry_point(cls):
    w = synthetic.exp_decay_batch(64)
    t, d = torch.as_tensor(w["t"], device=DEV), torch.as_tensor(w["d"], device=DEV)
    y0 = torch.as_tensor(w["y0"], device=DEV)

    def residual(y, a):
        return y[0] * torch.exp(-y[1] * a["t"]) + y[2] - a["d"]

    solver = cls(1e-8, 1e-8)
    args = {"t": t, "d": d}
    in_dims = {"t": None, "d": 0}
    y, _, state, result, steps = _interactive(solver, lambda y, a: (residual(y, a), None), y0, args, {"args_in_dims": in_dims})
    sol = optx.least_squares(residual, solver, y0, args, args_in_dims=in_dims, throw=False)
    assert torch.equal(sol.value, y), "stepping by hand and the entry point run the same kernels: bit-identical iterates"
    assert torch.equal(sol.stats["num_steps"], state.num_steps) and torch.equal(sol.result, result)
    assert steps == int(state.num_steps.max())
    # the state exposes the reference's fields
    info = state.f_info
    assert isinstance(info, FunctionInfo.ResidualGram)
    r = torch.func.vmap(residual, in_dims=(0, in_dims))(y, args)
    J = torch.func.vmap(torch.func.jacfwd(residual), in_dims=(0, in_dims))(y, args)
    assert torch.allclose(info.as_min(), 0.5 * (r * r).sum(1), rtol=1e-12)
    assert torch.allclose(info.compute_grad(), (J * r[:, :, None]).sum(1), rtol=1e-9, atol=1e-12)
    assert torch.allclose(info.gram, J.transpose(1, 2) @ J, rtol=1e-9, atol=1e-12)
    assert not bool(state.first_step.any()) and bool(state.terminate.all())
    assert (state.num_accepted_steps <= state.num_steps).all()
    if cls is optx.GaussNewton:
        assert (state.search_state == 1).all()  # LearningRate(1.0)
    # the registered family through the persistent kernel agrees to tolerance (different arithmetic order, same algorithm)
    reg = optx.least_squares(optx.models.exp_decay, solver, y0, (t, d), throw=False)
    assert ((reg.value - y).abs() / (1e-8 + 1e-8 * y.abs())).max() <= 1.0


def test_single_problem_stepping_and_partial_steps():
    """A single problem ([n] in, 0-d flags out) and 'a few steps at a time' (the use the reference's docs describe)."""
    def fn(y, a):
        return torch.stack([10.0 * (y[1] - y[0] ** 2), 1.0 - y[0]]), None  # Rosenbrock as a residual

    solver = optx.LevenbergMarquardt(1e-10, 1e-10)
    y = torch.tensor([-1.2, 1.0], dtype=torch.float64, device=DEV)
    state = solver.init(fn, y, None, {}, None, None, frozenset())
    assert bool(state.first_step) and int(state.num_steps) == 0
    for _ in range(3):
        y, state, _ = solver.step(fn, y, None, {}, state, frozenset())
    assert int(state.num_steps) == 3 and not bool(state.first_step)
    ref = optx.least_squares(lambda y, a: fn(y, a)[0], solver, torch.tensor([-1.2, 1.0], dtype=torch.float64, device=DEV), None,
                             max_steps=3, throw=False)
    assert torch.equal(ref.value, y) and ref.result == optx.RESULTS.nonlinear_max_steps_reached
    done, result = solver.terminate(fn, y, None, {}, state, frozenset())
    while not bool(done):
        y, state, _ = solver.step(fn, y, None, {}, state, frozenset())
        done, result = solver.terminate(fn, y, None, {}, state, frozenset())
    assert int(result) == optx.RESULTS.successful and torch.allclose(y, torch.ones(2, dtype=torch.float64, device=DEV), atol=1e-8)


@pytest.mark.parametrize("cls,kw", [(optx.BFGS, {}), (optx.BFGS, {"use_inverse": False}), (optx.DFP, {})])
def test_minimise_stepping_equals_entry_point(cls, kw):
    def f(y, a):
        return ((y - a) ** 2).sum() + 0.1 * (y ** 4).sum()

    gen = torch.Generator(device=DEV).manual_seed(0)
    y0 = torch.randn(16, 6, dtype=torch.float64, device=DEV, generator=gen)
    a = torch.randn(16, 6, dtype=torch.float64, device=DEV, generator=gen)
    solver = cls(1e-8, 1e-8, **kw)
    y, _, state, result, _ = _interactive(solver, lambda y, a: (f(y, a), None), y0, a, {"args_in_dims": 0}, max_steps=1000)
    sol = optx.minimise(f, solver, y0, a, args_in_dims=0, max_steps=1000, throw=False)
    assert torch.equal(sol.value, y) and torch.equal(sol.result, result) and torch.equal(sol.stats["num_steps"], state.num_steps)
    info = state.f_info
    assert isinstance(info, FunctionInfo.EvalGradHessianInv if kw.get("use_inverse", True) else FunctionInfo.EvalGradHessian)
    assert torch.allclose(info.as_min(), torch.func.vmap(f)(y, a), rtol=1e-12)
    assert (result == optx.RESULTS.successful).all()


def test_chord_stepping_equals_entry_point():
    w = synthetic.implicit_euler_batch(32, n=8)
    A, Bm = torch.as_tensor(w["A"], device=DEV), torch.as_tensor(w["B"], device=DEV)
    yp, dt = torch.as_tensor(w["y_prev"], device=DEV), torch.as_tensor(w["dt"], device=DEV)

    def F(y, a):
        return y - a["y_prev"] - a["dt"] * (A @ y + y * (Bm @ y))

    args, in_dims = {"y_prev": yp, "dt": dt}, {"y_prev": 0, "dt": 0}
    for solver in (optx.Chord(1e-8, 1e-8), optx.Newton(1e-8, 1e-8, cauchy_termination=False)):
        y, _, state, result, _ = _interactive(solver, lambda y, a: (F(y, a), None), yp.clone(), args, {"args_in_dims": in_dims})
        sol = optx.root_find(F, solver, yp.clone(), args, args_in_dims=in_dims, throw=False)
        assert torch.equal(sol.value, y) and torch.equal(sol.stats["num_steps"], state.num_steps) and torch.equal(sol.result, result)


def test_verbose(capsys):
    """verbose=True prints one line per step with the reference's labels; a callable receives name=(label, value) pairs;
    a registered family with verbose runs through the step kernel (the persistent kernel cannot print)."""
    w = synthetic.exp_decay_batch(1)
    t, d = torch.as_tensor(w["t"], device=DEV), torch.as_tensor(w["d"][0], device=DEV)
    y0 = torch.as_tensor(w["y0"][0], device=DEV)
    quiet = optx.least_squares(optx.models.exp_decay, optx.LevenbergMarquardt(1e-8, 1e-8), y0, (t, d), throw=False)
    capsys.readouterr()
    loud = optx.least_squares(optx.models.exp_decay, optx.LevenbergMarquardt(1e-8, 1e-8, verbose=True), y0, (t, d), throw=False)
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == int(loud.stats["num_steps"])
    for label in ("Step: 0", "Accepted steps:", "Loss on this step:", "Loss on the last accepted step:", "Step size:", "y:",
                  "y on the last accepted step:"):
        assert label in out[0]
    assert ((loud.value - quiet.value).abs() / (1e-8 + 1e-8 * quiet.value.abs())).max() <= 1.0
    seen = []
    optx.minimise(lambda y, a: ((y - 1.0) ** 2).sum(), optx.BFGS(1e-6, 1e-6, verbose=lambda **kw: seen.append(kw)),
                  torch.zeros(3, dtype=torch.float64, device=DEV), None, throw=False)
    assert seen and set(seen[0]) == {"loss_this_step", "loss_last_accepted_step", "step_size", "y", "y_last_accepted_step"}
    assert seen[0]["step_size"][0] == "Step size"


def test_linear_solver_is_never_silently_ignored():
    with pytest.raises(NotImplementedError, match="linear_solver"):
        optx.least_squares(optx.models.exp_decay, optx.LevenbergMarquardt(1e-8, 1e-8, linear_solver=optx.LU()),
                           torch.ones(3, dtype=torch.float64, device=DEV),
                           (torch.linspace(0, 4, 64, dtype=torch.float64, device=DEV), torch.ones(64, dtype=torch.float64, device=DEV)))


def test_least_squares_stepping_n20_cta_kernel_state():
    """16 < n <= 64 runs the one-CTA-per-problem step kernel (Gram in shared memory): same protocol, same state fields."""
    gen = torch.Generator(device=DEV).manual_seed(5)
    A = torch.randn(4, 30, 20, dtype=torch.float64, device=DEV, generator=gen)
    b = torch.randn(4, 30, dtype=torch.float64, device=DEV, generator=gen)

    def residual(y, a):
        return a["A"] @ y - a["b"] + 0.1 * torch.sin(y).sum()

    solver = optx.LevenbergMarquardt(1e-9, 1e-9)
    args, in_dims = {"A": A, "b": b}, {"A": 0, "b": 0}
    y0 = torch.zeros(4, 20, dtype=torch.float64, device=DEV)
    y, _, state, result, _ = _interactive(solver, lambda y, a: (residual(y, a), None), y0, args, {"args_in_dims": in_dims})
    sol = optx.least_squares(residual, solver, y0, args, args_in_dims=in_dims, throw=False)
    assert torch.equal(sol.value, y) and torch.equal(sol.result, result) and (result == optx.RESULTS.successful).all()
    info = state.f_info
    r = torch.func.vmap(residual, in_dims=(0, in_dims))(y, args)
    J = torch.func.vmap(torch.func.jacfwd(residual), in_dims=(0, in_dims))(y, args)
    assert torch.allclose(info.gram, J.transpose(1, 2) @ J, rtol=1e-9, atol=1e-10)
    assert torch.allclose(info.compute_grad(), (J * r[:, :, None]).sum(1), rtol=1e-8, atol=1e-10)
    assert torch.allclose(info.as_min(), 0.5 * (r * r).sum(1), rtol=1e-12)
    assert bool(state.terminate.all()) and (state.search_state > 0).all()
    with pytest.raises(ValueError, match="empty"):
        solver.init(lambda y, a: (residual(y, a), None), y0[:0], {"A": A[:0], "b": b[:0]}, {"args_in_dims": in_dims}, None, None, frozenset())
