"""Pytree-valued `y` (optimistix_b200/_pytree.py): layout on the CPU, solves on the GPU.

The reference takes any pytree as `y0` and its fixtures use them (tests/helpers.py:316-320 `bowl` on a pytree, the
minimisation zoo's inits at :540-562 are nested lists/tuples/dicts of arrays).
"""
import numpy as np
import pytest
import torch

import optimistix_b200 as optx
from optimistix_b200 import _pytree


def _tree(batch=None, device="cpu", dtype=torch.float64):
    lead = () if batch is None else (batch,)
    g = torch.Generator().manual_seed(3)
    mk = lambda *s: torch.randn(lead + tuple(s), generator=g, dtype=dtype).to(device)
    return {"a": mk(2, 3), "b": (mk(), [mk(4)])}


def test_is_array_like():
    assert _pytree.is_array_like(torch.zeros(3)) and _pytree.is_array_like(np.zeros((2, 3))) and _pytree.is_array_like([1.0, 2.0])
    assert not _pytree.is_array_like(torch.tensor(1.0)) and not _pytree.is_array_like(1.0)
    assert not _pytree.is_array_like({"a": torch.zeros(3)}) and not _pytree.is_array_like((torch.zeros(3), torch.zeros(2)))
    assert not _pytree.is_array_like([torch.tensor(1.0), torch.tensor(2.0)])


@pytest.mark.parametrize("batch", [None, 5])
def test_ravel_unravel_round_trip(batch):
    y0 = _tree(batch)
    rav = _pytree.Raveled(y0, batched=batch is not None)
    assert rav.n == 6 + 1 + 4 and rav.sizes == [6, 1, 4]
    assert tuple(rav.flat0.shape) == ((11,) if batch is None else (5, 11))
    back = rav.unravel(rav.flat0)
    assert torch.equal(back["a"], y0["a"]) and torch.equal(back["b"][0], y0["b"][0]) and torch.equal(back["b"][1][0], y0["b"][1][0])
    # leaf order == flattening order (the order jax.flatten_util.ravel_pytree uses)
    lead = slice(None) if batch is None else (slice(None), slice(None))
    first = y0["a"].reshape(-1) if batch is None else y0["a"].reshape(5, -1)
    assert torch.equal(rav.flat0[..., :6], first)


def test_mixed_dtypes_promote_and_integers_become_inexact():
    rav = _pytree.Raveled((torch.zeros(2, dtype=torch.float32), torch.ones(3, dtype=torch.float64), torch.tensor(2)), batched=False)
    assert rav.flat0.dtype == torch.float64 and rav.n == 6
    assert _pytree.Raveled((torch.tensor(2), torch.tensor([1, 2])), batched=False).flat0.dtype == torch.get_default_dtype()


def test_batch_size_mismatch_and_missing_axis():
    with pytest.raises(ValueError, match="batch size"):
        _pytree.Raveled((torch.zeros(3, 2), torch.zeros(4, 2)), batched=True)
    with pytest.raises(ValueError, match="leading batch axis"):
        _pytree.Raveled((torch.zeros(3, 2), torch.tensor(1.0)), batched=True)


def test_wrapped_fn_under_torch_func_transforms():
    """The wrapper is called under vmap(jacfwd(.)) by the generic path: unravel must be slicing / reshape only."""
    y0 = _tree()
    rav = _pytree.Raveled(y0, batched=False)

    def fn(y, args):  # residual pytree
        return {"r1": y["a"] ** 2 * args, "r2": (y["b"][0] * y["b"][1][0],)}

    g = _pytree.wrap_fn(fn, rav, "vector", has_aux=False)
    Y = rav.flat0[None].repeat(3, 1) * torch.arange(1, 4, dtype=torch.float64)[:, None]
    r = torch.func.vmap(lambda y: g(y, 2.0))(Y)
    J = torch.func.vmap(torch.func.jacfwd(lambda y: g(y, 2.0)))(Y)
    assert tuple(r.shape) == (3, 10) and tuple(J.shape) == (3, 10, 11)
    a = Y[1, :6]
    assert torch.allclose(r[1, :6], 2.0 * a * a) and torch.allclose(torch.diagonal(J[1, :6, :6]), 4.0 * a)
    # minimise: scalar output; has_aux passes the auxiliary output through
    gs = _pytree.wrap_fn(lambda y, a: ((y["a"] ** 2).sum() + y["b"][0] ** 2, {"n": y["b"][1][0].sum()}), rav, "scalar", has_aux=True)
    val, aux = gs(rav.flat0, None)
    assert val.ndim == 0 and torch.allclose(aux["n"], y0["b"][1][0].sum())
    with pytest.raises(ValueError, match="single scalar"):
        _pytree.wrap_fn(lambda y, a: y["a"], rav, "scalar", has_aux=False)(rav.flat0, None)
    # fixed_point: output must have y's structure
    gf = _pytree.wrap_fn(lambda y, a: {"a": 0.5 * y["a"], "b": (0.5 * y["b"][0], [0.5 * y["b"][1][0]])}, rav, "same", has_aux=False)
    assert torch.allclose(gf(rav.flat0, None), 0.5 * rav.flat0)
    with pytest.raises(ValueError, match="same structure"):
        _pytree.wrap_fn(lambda y, a: {"a": y["a"]}, rav, "same", has_aux=False)(rav.flat0, None)


def test_registered_family_rejects_pytree_y0():
    with pytest.raises(TypeError, match="flat parameter vector"):
        optx.least_squares(optx.models.gauss2, optx.LevenbergMarquardt(1e-4, 1e-4), {"a": torch.zeros(8)}, None)


# ------------------------------------------------------------------------------------------------ GPU: solves
def _bowl(y, args):  # helpers.py:316-320: y is a pytree, raveled inside
    flat = torch.cat([x.reshape(-1) for x in (y["a"], y["b"][0], y["b"][1][0])])
    return flat @ args @ flat


@pytest.mark.gpu
def test_minimise_bowl_on_a_pytree_bfgs():
    y0 = _tree(device="cuda")
    g = torch.Generator().manual_seed(0)
    M = torch.randn(11, 11, generator=g, dtype=torch.float64)
    A = (M @ M.T + torch.eye(11, dtype=torch.float64)).cuda()
    sol = optx.minimise(_bowl, optx.BFGS(1e-9, 1e-9), y0, A, max_steps=512)
    assert set(sol.value.keys()) == {"a", "b"} and tuple(sol.value["a"].shape) == (2, 3) and sol.value["b"][0].ndim == 0
    flat = torch.cat([sol.value["a"].reshape(-1), sol.value["b"][0].reshape(-1), sol.value["b"][1][0]])
    assert float(flat.abs().max()) < 1e-5  # minimum of a positive-definite bowl is 0 (tests/test_minimise.py:30-62 tolerance 1e-4..1e-5)
    assert int(sol.result) == optx.RESULTS.successful


@pytest.mark.gpu
@pytest.mark.parametrize("solver", [optx.LevenbergMarquardt(1e-10, 1e-10), optx.GaussNewton(1e-10, 1e-10)])
def test_least_squares_pytree_y_and_pytree_residuals_batched(solver):
    """Batched pytree y (y0_in_dims=0), per-problem args, residual pytree; equals the same problem posed on flat vectors."""
    B = 7
    y0 = _tree(B, device="cuda")
    tgt = torch.linspace(0.5, 1.5, B, dtype=torch.float64, device="cuda")

    def fn(y, t):
        return {"p": y["a"] ** 3 + y["a"] - 0.1 * t, "q": (y["b"][0] - t, y["b"][1][0] ** 3 + y["b"][1][0] - t)}

    sol = optx.least_squares(fn, solver, y0, tgt, args_in_dims=0, y0_in_dims=0, throw=False)
    rav = _pytree.Raveled(y0, batched=True)

    def fn_flat(v, t):
        return torch.cat([(v[:6] ** 3 + v[:6] - 0.1 * t), (v[6:7] - t), v[7:] ** 3 + v[7:] - t])

    ref = optx.least_squares(fn_flat, solver, rav.flat0, tgt, args_in_dims=0, throw=False)
    assert tuple(sol.value["a"].shape) == (B, 2, 3) and tuple(sol.value["b"][0].shape) == (B,) and tuple(sol.value["b"][1][0].shape) == (B, 4)
    assert torch.equal(rav.unravel(ref.value)["a"], sol.value["a"]) and torch.equal(ref.value[:, 6], sol.value["b"][0])
    assert torch.equal(ref.stats["num_steps"], sol.stats["num_steps"]) and torch.equal(ref.result, sol.result)
    assert bool((sol.result == optx.RESULTS.successful).all())
    assert torch.allclose(sol.value["a"] ** 3 + sol.value["a"], (0.1 * tgt)[:, None, None].expand(B, 2, 3), atol=1e-8)


@pytest.mark.gpu
def test_root_find_and_fixed_point_scalar_and_tuple_y():
    # scalar y (0-d), as the reference's tests/test_fixed_point.py / test_root_find.py use for `trivial`-style problems
    sol = optx.root_find(lambda y, a: y ** 3 + y - a, optx.Newton(1e-10, 1e-10), torch.tensor(0.5, dtype=torch.float64, device="cuda"),
                         torch.tensor(2.0, dtype=torch.float64, device="cuda"))
    assert sol.value.ndim == 0 and abs(float(sol.value) - 1.0) < 1e-9
    # fixed point of a contraction on a tuple: y = (cos(y1), 0.5 * y0 + 1)
    y0 = (torch.zeros(3, dtype=torch.float64, device="cuda"), torch.ones(3, dtype=torch.float64, device="cuda"))
    sol = optx.fixed_point(lambda y, a: (torch.cos(y[1]), 0.5 * y[0] + 1.0), optx.Newton(1e-10, 1e-10), y0)
    a, b = sol.value
    assert torch.allclose(a, torch.cos(b), atol=1e-9) and torch.allclose(b, 0.5 * a + 1.0, atol=1e-9)
    # has_aux with a pytree y
    sol = optx.root_find(lambda y, a: ((y[0] - 1.0, y[1] * 2.0 - y[0]), {"s": y[0].sum()}), optx.Newton(1e-10, 1e-10), y0, has_aux=True)
    assert torch.allclose(sol.value[0], torch.ones(3, dtype=torch.float64, device="cuda")) and abs(float(sol.aux["s"]) - 3.0) < 1e-8
