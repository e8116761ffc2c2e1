"""`optimistix.compat.minimize`: the drop-in for the deprecated `jax.scipy.optimize.minimize` (reference:
optimistix/compat/_impl.py:14-125), on the CUDA BFGS path.

    res = optimistix_b200.compat.minimize(fun, x0, args=(), method="bfgs", tol=None, options=None)

`fun(x, *args) -> scalar` with `x` a 1-D torch tensor; derivatives come from torch autograd, every other part of
the BFGS iteration runs in `b200opt_bfgs_step` (optimistix_b200/csrc/generic_step.cu).
"""
from __future__ import annotations

from typing import Any, Callable, Mapping, NamedTuple

from ._entry import minimise
from ._solution import RESULTS
from ._solvers import BFGS, max_norm


class OptimizeResults(NamedTuple):
    """compat/_impl.py:14-40: x, success, status (0 converged, 1 max BFGS iterations reached, 3 other failure), fun,
    jac, hess_inv, nfev, njev, nit."""

    x: Any
    success: Any
    status: Any
    fun: Any
    jac: Any
    hess_inv: Any
    nfev: Any
    njev: Any
    nit: Any


def minimize(fun: Callable, x0, args: tuple = (), *, method: str, tol: float | None = None,
             options: Mapping[str, Any] | None = None) -> OptimizeResults:
    import torch

    if method.lower() != "bfgs":
        raise ValueError(f"Method {method} not recognized")  # compat/_impl.py:87-88
    if not isinstance(x0, torch.Tensor) or x0.ndim != 1:
        raise ValueError("x0 must be a 1-dimensional array")  # :89-90
    if not isinstance(args, tuple):
        # TypeError, not ValueError, for compatibility with the old jax.scipy.optimize.minimize (:91-95)
        raise TypeError(f"args argument to `optimistix.compat.minimize` must be a tuple, got {args}")
    if tol is None:
        tol = 1e-5  # :96-97
    options = {} if options is None else dict(options)
    max_steps = options.pop("maxiter", None)  # None = unlimited (:103)
    options.pop("norm", max_norm)  # accepted and ignored, as in the reference (:104, :111)
    if len(options) != 0:
        raise ValueError(f"Unsupported options: {set(options.keys())}")  # :105-106

    def wrapped_fn(y, a):
        return fun(y, *a)

    solver = BFGS(rtol=tol, atol=tol, norm=max_norm)
    sol = minimise(wrapped_fn, solver, x0, args, max_steps=max_steps, throw=False, return_hessian_inv=True)
    ok = sol.result == RESULTS.successful
    status = torch.where(ok, 0, torch.where(sol.result == RESULTS.nonlinear_max_steps_reached, 1, 3))  # :113-117
    nacc = sol.stats["num_accepted_steps"][0]
    return OptimizeResults(x=sol.value, success=ok, status=status, fun=sol.state["f"][0], jac=sol.state["grad"][0],
                           hess_inv=sol.state["hessian_inv"][0], nfev=sol.stats["num_steps"], njev=nacc, nit=nacc)
