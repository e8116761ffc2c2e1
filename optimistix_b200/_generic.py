"""Generic path for an arbitrary user `fn`: the per-iteration CUDA step kernel fed caller-computed r and J.

This is what makes the library a drop-in beyond the registered model families (SURVEY section 8f.1;
BASELINE.json north_star: "for arbitrary user fn, a per-iteration step kernel is fed XLA-computed J and r").
The reference obtains (r, J) with `jax.linearize` / `jax.jacrev` (optimistix/_solver/gauss_newton.py:158-181);
here `torch.func.jacfwd` / `jacrev` plays that role (options["jac"] = "fwd" | "bwd"), batched with
`torch.func.vmap`, and everything else of the step - search, accept/reject, Gram, damped Cholesky solve,
termination, result codes - runs in `b200opt_lsq_step` (optimistix_b200/csrc/lsq_step.cu).

`fn(y, args)` takes ONE problem: y of shape [n], returns residuals [m] (any shape; it is raveled).
`args_in_dims` says which leaves of `args` carry the batch axis (torch.func.vmap in_dims for `args`;
default None = shared by all problems), mirroring `jax.vmap(..., in_axes=...)` at the reference's call site.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._adjoint import attach_if_needed
from ._protocol import LsqSession, MinSession, RootSession
from ._solution import RESULTS, Solution


def _torch():
    import torch

    return torch


def _run_iterations(first_body, body, active, max_steps, graph_ok=True):
    """Drive the per-iteration loop: `first_body()` / `body()` evaluate the user's function (and derivatives) at the current
    point and enqueue ONE step kernel; `active` is the device counter the kernel leaves behind.

    The first iteration runs eagerly (lazy initialisation, attribute setting).  The rest of the loop is launch bound on the
    host - torch.func dispatches ~30 small kernels per iteration - so one iteration is captured into a CUDA graph and
    replayed (static shapes, the carry lives in fixed buffers, frozen problems are no-ops inside the step kernel).  If the
    user's function cannot be captured (it synchronises, allocates outside the pool, ...) the loop falls back to eager
    launches.  B200OPT_GENERIC_GRAPH=0 disables capture.  Returns the number of step-kernel launches."""
    import os

    t = _torch()
    first_body()
    launches = 1
    if launches >= max_steps or int(active.item()) == 0:
        return launches
    graph = None
    if graph_ok and os.environ.get("B200OPT_GENERIC_GRAPH", "1") != "0" and max_steps > 3:  # a verbose solve prints from the host every step
        try:
            body()  # second eager iteration: the steady-state form of the body (Chord: no Jacobian) warms up too
            launches += 1
            if launches >= max_steps or int(active.item()) == 0:
                return launches
            g = t.cuda.CUDAGraph()
            with t.cuda.graph(g):
                body()
            graph = g
        except Exception:  # noqa: BLE001 - not capturable: eager loop below
            graph = None
            t.cuda.synchronize()
    while launches < max_steps:
        if graph is not None:
            graph.replay()
        else:
            body()
        launches += 1
        if int(active.item()) == 0:  # every lane of the vmapped while_loop has stopped (_iterate.py:230-232)
            break
    return launches


def _empty_solution(Y0, max_steps):
    """A batch of zero problems: nothing to launch (and torch.func.vmap refuses an empty batch axis)."""
    t = _torch()
    z = t.zeros((0,), dtype=t.int32, device=Y0.device)
    return Solution(value=Y0.clone(), result=z, aux=None, stats={"num_steps": z.clone(), "max_steps": (None if max_steps is None or max_steps >= 2**31 - 1 else max_steps)}, state=None)


def _throw_if_failed(res):
    rh = res.cpu().numpy()
    bad = np.flatnonzero(rh != RESULTS.successful)
    if bad.size:
        code = int(rh[bad[0]])
        raise RuntimeError(f"{bad.size} of {rh.size} solves failed; first failure (problem {int(bad[0])}): "
                           f"{RESULTS.name(code)}: {RESULTS[code]}")


def _max_steps_stat(max_steps):
    return None if max_steps is None or max_steps >= 2**31 - 1 else max_steps


def _drive(sess, max_steps_k):
    """The driver loop of _iterate.py:194-268 around a session (= the solver protocol's `step`), then the max-steps result."""
    first = getattr(sess, "first_body", sess.body)
    launches = _run_iterations(first, sess.body, sess.active, max_steps_k, graph_ok=sess.verbose is None) if max_steps_k > 0 else 0
    sess.launches += launches
    sess.finish()
    return launches


def least_squares_generic(fn, solver, y0, args, *, max_steps, throw, root_near_zero=False, options=None, args_in_dims=None,
                          adjoint=None, implicit=None):
    """LM / GN for an arbitrary residual function (device tensors only: the Jacobian is evaluated by torch on the GPU)."""
    sess = LsqSession(fn, solver, y0, args, options, args_in_dims=args_in_dims, max_steps=max_steps, root_near_zero=root_near_zero)
    if sess.empty:
        return _empty_solution(sess.Y0, max_steps)
    launches = _drive(sess, sess.max_steps_k)
    if throw:
        _throw_if_failed(sess.res)
    stats = {"num_steps": sess.num_steps, "max_steps": _max_steps_stat(max_steps), "step_kernel_launches": launches}
    kind, ifn = implicit or ("root" if root_near_zero else "lstsq", fn)
    y = attach_if_needed(adjoint, kind, ifn, sess.y, args, args_in_dims)
    return Solution(value=sess._sq(y), result=sess.result, aux=None, stats=stats, state=sess.state)


def root_find_generic(fn, solver, y0, args, options, *, max_steps, throw, args_in_dims=None, adjoint=None):
    """Newton / Chord for an arbitrary `fn(y, args) -> [n]`: torch evaluates f and the n x n Jacobian (the role of
    jax.linearize / jacfwd at newton_chord.py:116-119; Chord: once, at y0, :80-87), `b200opt_root_step` does the LU
    solve with partial pivoting, the update, clipping, termination and result codes."""
    sess = RootSession(fn, solver, y0, args, options, args_in_dims=args_in_dims, max_steps=max_steps)
    if sess.empty:
        return _empty_solution(sess.Y0, max_steps)
    launches = _drive(sess, sess.max_steps_k)
    if throw:
        _throw_if_failed(sess.res)
    stats = {"num_steps": sess.num_steps, "max_steps": _max_steps_stat(max_steps), "step_kernel_launches": launches}
    y = attach_if_needed(adjoint, "root", fn, sess.y, args, args_in_dims)
    return Solution(value=sess._sq(y), result=sess.result, aux=None, stats=stats, state=sess.state)


def minimise_generic(fn, solver, y0, args, *, max_steps, throw, args_in_dims=None, root_fn=None, return_hessian_inv=False,
                     adjoint=None, implicit=None, autodiff_mode="bwd"):
    """BFGS (inverse Hessian, backtracking Armijo) for an arbitrary scalar `fn(y, args)`: torch evaluates f and its
    gradient (jax.linearize + linear_transpose at quasi_newton.py:214-227), `b200opt_bfgs_step` does the rest.

    root_fn: when minimising 0.5 |root_fn|^2 on behalf of root_find (_root_find.py:45-80), the extra
    `norm(root_fn(y)) < atol` termination condition is evaluated on the pre-step y.
    max_steps=None (compat.minimize without maxiter, compat/_impl.py:103): run until every problem stops."""
    t = _torch()
    sess = MinSession(fn, solver, y0, args, None, args_in_dims=args_in_dims, max_steps=max_steps, root_fn=root_fn, autodiff_mode=autodiff_mode)
    if sess.empty:
        return _empty_solution(sess.Y0, max_steps)
    launches = _drive(sess, sess.max_steps_k)
    if throw:
        _throw_if_failed(sess.res)
    n, B, state = sess.n, sess.B, sess.state
    stats = {"num_steps": sess.num_steps, "max_steps": _max_steps_stat(max_steps), "step_kernel_launches": launches,
             "num_accepted_steps": state[:, n * n + 2 * n + 4].to(t.int32)}
    st = {"f": state[:, n * n + 2 * n], "grad": state[:, n * n : n * n + n]}
    if return_hessian_inv:
        st["hessian_inv" if sess.use_inverse else "hessian"] = state[:, : n * n].reshape(B, n, n)
    kind, ifn = implicit or (("root", root_fn) if root_fn is not None else ("min", fn))
    y = attach_if_needed(adjoint, kind, ifn, sess.y, args, args_in_dims)
    return Solution(value=sess._sq(y), result=sess.result, aux=None, stats=stats, state=st)
