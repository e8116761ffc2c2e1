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

import ctypes as C

import numpy as np

from . import _lib
from ._adjoint import attach_if_needed
from ._solution import RESULTS, Solution


def _torch():
    import torch

    return torch


def _batched_r_and_J(fn, args, args_in_dims, jac_mode, single):
    t = _torch()

    jacf = t.func.jacfwd if jac_mode == "fwd" else t.func.jacrev

    if args_in_dims is None:
        def one(y):
            r = fn(y, args)
            return r.reshape(-1)

        def value_and_aux(z):
            r = one(z)  # ONE evaluation of the user function per Jacobian pass: the residual rides along as `aux`
            return r, r

        def both(y):
            J, r = jacf(value_and_aux, has_aux=True)(y)
            return r, J

        return t.func.vmap(both)

    def one_a(y, a):
        return fn(y, a).reshape(-1)

    def both_a(y, a):
        def value_and_aux(z):
            r = one_a(z, a)
            return r, r

        J, r = jacf(value_and_aux, has_aux=True)(y)
        return r, J

    vm = t.func.vmap(both_a, in_dims=(0, args_in_dims))
    return lambda Y: vm(Y, args)


def least_squares_generic(fn, solver, y0, args, *, max_steps, throw, root_near_zero=False, options=None, args_in_dims=None,
                          adjoint=None, implicit=None):
    """LM / GN for an arbitrary residual function (device tensors only: the Jacobian is evaluated by torch on the GPU)."""
    from ._solvers import LevenbergMarquardt

    t = _torch()
    lib = _lib.load()
    y0 = t.as_tensor(y0)
    if not y0.is_cuda:
        if not t.cuda.is_available():
            raise _lib.B200OptError("the generic path needs a CUDA device (there is no CPU fallback)")
        y0 = y0.cuda()
    if not y0.dtype.is_floating_point:
        y0 = y0.to(t.get_default_dtype())
    if y0.dtype not in (t.float32, t.float64):
        raise TypeError(f"unsupported dtype {y0.dtype}")
    single = y0.ndim == 1
    Y0 = (y0[None] if single else y0).contiguous()
    B, n = Y0.shape
    if n > 64:
        raise NotImplementedError("generic least-squares step kernels support n <= 64 parameters (n <= 16: Gram in registers, one warp "
                                  "per problem; n <= 64: Gram in shared memory, one CTA per problem)")
    dtype_code = _lib.F32 if y0.dtype == t.float32 else _lib.F64
    if B == 0:
        return _empty_solution(Y0, max_steps)
    jac_mode = (options or {}).get("jac", "fwd")
    rj = _batched_r_and_J(fn, _detached(args), args_in_dims, jac_mode, single)
    r, J = rj(Y0)
    m = int(r.shape[1])
    es = 4 if dtype_code == _lib.F32 else 8
    stride = lib.b200opt_lsq_step_state_size(dtype_code, n, m)
    if stride <= 0:
        _lib.check(-1, "b200opt_lsq_step_state_size")
    state = t.empty((B, stride // es), dtype=y0.dtype, device=Y0.device)
    y = t.empty_like(Y0)
    y_eval = t.empty_like(Y0)
    steps = t.empty((B,), dtype=t.int32, device=Y0.device)
    res = t.empty((B,), dtype=t.int32, device=Y0.device)
    active = t.zeros((1,), dtype=t.int32, device=Y0.device)
    stream = t.cuda.current_stream(Y0.device).cuda_stream
    solver_id = _lib.SOLVER_LM if isinstance(solver, LevenbergMarquardt) else _lib.SOLVER_GN
    opts = _lib.make_options(solver.rtol, solver.atol, solver.norm, max_steps, root_near_zero=root_near_zero)
    _lib.check(lib.b200opt_lsq_step_init(dtype_code, B, n, m, Y0.data_ptr(), y.data_ptr(), y_eval.data_ptr(), state.data_ptr(),
                                         steps.data_ptr(), res.data_ptr(), stream), "b200opt_lsq_step_init")
    first = {"rJ": (r, J)}

    def body():
        rr, JJ = first.pop("rJ", None) or rj(y_eval)
        rr, JJ = rr.contiguous(), JJ.contiguous()
        st = t.cuda.current_stream(Y0.device).cuda_stream  # the capture stream while a graph is being recorded
        _lib.check(lib.b200opt_lsq_step(solver_id, dtype_code, B, n, m, rr.data_ptr(), JJ.data_ptr(), y.data_ptr(), y_eval.data_ptr(),
                                        state.data_ptr(), C.byref(opts), steps.data_ptr(), res.data_ptr(), active.data_ptr(), st),
                   "b200opt_lsq_step")
        body.keep = (rr, JJ)  # the eager path must keep its inputs alive until the kernel has run

    launches = _run_iterations(body, body, active, max_steps) if max_steps > 0 else 0
    _lib.check(lib.b200opt_lsq_step_finish(dtype_code, B, n, m, state.data_ptr(), res.data_ptr(), stream), "b200opt_lsq_step_finish")
    if throw:
        rh = res.cpu().numpy()
        bad = np.flatnonzero(rh != RESULTS.successful)
        if bad.size:
            code = int(rh[bad[0]])
            raise RuntimeError(f"{bad.size} of {rh.size} solves failed; first failure (problem {int(bad[0])}): "
                               f"{RESULTS.name(code)}: {RESULTS[code]}")
    stats = {"num_steps": steps, "max_steps": (None if max_steps is None or max_steps >= 2**31 - 1 else max_steps), "step_kernel_launches": launches}
    kind, ifn = implicit or ("root" if root_near_zero else "lstsq", fn)
    y = attach_if_needed(adjoint, kind, ifn, y, args, args_in_dims)
    value, result = y, res
    if single:
        value, result = y[0], res[0]
        stats["num_steps"] = steps[0]
    return Solution(value=value, result=result, aux=None, stats=stats, state=state)


def _prep_y0(y0):
    t = _torch()
    y0 = t.as_tensor(y0)
    if not y0.is_cuda:
        if not t.cuda.is_available():
            raise _lib.B200OptError("the generic path needs a CUDA device (there is no CPU fallback)")
        y0 = y0.cuda()
    if not y0.dtype.is_floating_point:
        y0 = y0.to(t.get_default_dtype())
    if y0.dtype not in (t.float32, t.float64):
        raise TypeError(f"unsupported dtype {y0.dtype}")
    single = y0.ndim == 1
    if y0.ndim not in (1, 2):
        raise ValueError("y0 must have shape [n] (one problem) or [batch, n]")
    Y0 = (y0[None] if single else y0).contiguous()
    return Y0, single, (_lib.F32 if y0.dtype == t.float32 else _lib.F64)


def _run_iterations(first_body, body, active, max_steps):
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
    if os.environ.get("B200OPT_GENERIC_GRAPH", "1") != "0" and max_steps > 3:
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


def _detached(args):
    """args with every tensor leaf detached: the forward iteration is never recorded by autograd (ImplicitAdjoint
    differentiates the optimality condition at the solution instead, _adjoint.py)."""
    t = _torch()
    from torch.utils._pytree import tree_map

    return tree_map(lambda x: x.detach() if isinstance(x, t.Tensor) else x, args)


def _vmapped(f, args, args_in_dims):
    """vmap f(y, args) over the batch axis of y (and of the leaves of args selected by args_in_dims)."""
    t = _torch()
    if args_in_dims is None:
        vm = t.func.vmap(lambda y: f(y, args))
        return vm
    vm = t.func.vmap(f, in_dims=(0, args_in_dims))
    return lambda Y: vm(Y, args)


def root_find_generic(fn, solver, y0, args, options, *, max_steps, throw, args_in_dims=None, adjoint=None):
    """Newton / Chord for an arbitrary `fn(y, args) -> [n]`: torch evaluates f and the n x n Jacobian (the role of
    jax.linearize / jacfwd at newton_chord.py:116-119; Chord: once, at y0, :80-87), `b200opt_root_step` does the LU
    solve with partial pivoting, the update, clipping, termination and result codes."""
    from ._solvers import Chord

    t = _torch()
    lib = _lib.load()
    Y0, single, dtype_code = _prep_y0(y0)
    B, n = Y0.shape
    if n > 32:
        raise NotImplementedError("generic Newton/Chord step kernel supports n <= 32 unknowns (row-per-lane LU)")
    dev = Y0.device
    if B == 0:
        return _empty_solution(Y0, max_steps)

    def one(y, a):
        out = fn(y, a).reshape(-1)
        if out.shape[0] != n:
            raise ValueError("root_find: fn must map [n] -> [n]")
        return out

    args_fwd = _detached(args)
    f_only = _vmapped(one, args_fwd, args_in_dims)
    f_and_J = _vmapped(lambda y, a: tuple(reversed(t.func.jacfwd(lambda z: (one(z, a), one(z, a)), has_aux=True)(y))), args_fwd,
                       args_in_dims)
    lower = upper = None
    if options:
        if options.get("lower") is not None:
            lower = t.broadcast_to(t.as_tensor(options["lower"], dtype=Y0.dtype, device=dev), (n,)).contiguous()
        if options.get("upper") is not None:
            upper = t.broadcast_to(t.as_tensor(options["upper"], dtype=Y0.dtype, device=dev), (n,)).contiguous()
    es = 4 if dtype_code == _lib.F32 else 8
    stride = lib.b200opt_root_step_state_size(dtype_code, n)
    state = t.empty((B, stride // es), dtype=Y0.dtype, device=dev)
    y = t.empty_like(Y0)
    steps = t.empty((B,), dtype=t.int32, device=dev)
    res = t.empty((B,), dtype=t.int32, device=dev)
    active = t.zeros((1,), dtype=t.int32, device=dev)
    stream = t.cuda.current_stream(dev).cuda_stream
    opts = _lib.make_options(solver.rtol, solver.atol, solver.norm, max_steps, cauchy_termination=solver.cauchy_termination,
                             kappa=solver.kappa)
    _lib.check(lib.b200opt_root_step_init(dtype_code, B, n, Y0.data_ptr(), y.data_ptr(), state.data_ptr(), steps.data_ptr(),
                                          res.data_ptr(), stream), "b200opt_root_step_init")
    chord = isinstance(solver, Chord)
    hold = {}

    def launch(f, J):
        f, J = f.contiguous(), J.contiguous()
        st = t.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.b200opt_root_step(dtype_code, B, n, f.data_ptr(), J.data_ptr(), lower.data_ptr() if lower is not None else None,
                                         upper.data_ptr() if upper is not None else None, y.data_ptr(), state.data_ptr(),
                                         C.byref(opts), steps.data_ptr(), res.data_ptr(), active.data_ptr(), st),
                   "b200opt_root_step")
        hold["fJ"] = (f, J)

    def first_body():
        f, J = f_and_J(y)
        if chord:
            hold["J_fixed"] = J.contiguous()  # newton_chord.py:80-87: linearised once, at y0
        launch(f, J)

    def body():
        if chord:
            launch(f_only(y), hold["J_fixed"])
        else:
            launch(*f_and_J(y))

    launches = _run_iterations(first_body, body, active, max_steps) if max_steps > 0 else 0
    _lib.check(lib.b200opt_root_step_finish(dtype_code, B, n, state.data_ptr(), res.data_ptr(), stream), "b200opt_root_step_finish")
    if throw:
        _throw_if_failed(res)
    stats = {"num_steps": steps[0] if single else steps, "max_steps": (None if max_steps is None or max_steps >= 2**31 - 1 else max_steps), "step_kernel_launches": launches}
    y = attach_if_needed(adjoint, "root", fn, y, args, args_in_dims)
    return Solution(value=y[0] if single else y, result=res[0] if single else res, aux=None, stats=stats, state=state)


def minimise_generic(fn, solver, y0, args, *, max_steps, throw, args_in_dims=None, root_fn=None, return_hessian_inv=False,
                     adjoint=None, implicit=None, autodiff_mode="bwd"):
    """BFGS (inverse Hessian, backtracking Armijo) for an arbitrary scalar `fn(y, args)`: torch evaluates f and its
    gradient (jax.linearize + linear_transpose at quasi_newton.py:214-227), `b200opt_bfgs_step` does the rest.

    root_fn: when minimising 0.5 |root_fn|^2 on behalf of root_find (_root_find.py:45-80), the extra
    `norm(root_fn(y)) < atol` termination condition is evaluated on the pre-step y."""
    t = _torch()
    lib = _lib.load()
    Y0, single, dtype_code = _prep_y0(y0)
    B, n = Y0.shape
    dev = Y0.device
    if B == 0:
        return _empty_solution(Y0, max_steps)

    def one(y, a):
        out = fn(y, a)
        if out.numel() != 1:
            raise ValueError("minimise function must output a single scalar.")  # _minimise.py:97-104
        return out.reshape(())

    args_fwd = _detached(args)
    if autodiff_mode == "fwd":  # options["autodiff_mode"] (_misc.py:245-258): gradient by forward-mode JVPs
        gv = _vmapped(lambda y, a: t.func.jacfwd(lambda z: (one(z, a), one(z, a)), has_aux=True)(y), args_fwd, args_in_dims)
    else:
        gv = _vmapped(lambda y, a: t.func.grad_and_value(lambda z: one(z, a))(y), args_fwd, args_in_dims)
    rootv = None
    if root_fn is not None:
        norm = solver.norm
        rootv = _vmapped(lambda y, a: norm(root_fn(y, a).reshape(-1)), args_fwd, args_in_dims)
    es = 4 if dtype_code == _lib.F32 else 8
    nbytes = lib.b200opt_bfgs_step_state_size(dtype_code, n)
    if nbytes <= 0:
        _lib.check(-1, "b200opt_bfgs_step_state_size")
    state = t.empty((B, nbytes // es), dtype=Y0.dtype, device=dev)
    y = t.empty_like(Y0)
    y_eval = t.empty_like(Y0)
    steps = t.empty((B,), dtype=t.int32, device=dev)
    res = t.empty((B,), dtype=t.int32, device=dev)
    active = t.zeros((1,), dtype=t.int32, device=dev)
    stream = t.cuda.current_stream(dev).cuda_stream
    unbounded = max_steps is None  # compat.minimize without maxiter (compat/_impl.py:103): run until every problem stops
    max_steps_k = (2**31 - 1) if unbounded else max_steps
    opts = _lib.make_options(solver.rtol, solver.atol, solver.norm, max_steps_k)
    update = _lib.UPDATE_DFP if getattr(solver, "_update", "bfgs") == "dfp" else _lib.UPDATE_BFGS
    use_inverse = 1 if getattr(solver, "use_inverse", True) else 0
    if not use_inverse and n > 64:
        raise NotImplementedError("use_inverse=False factorises the Hessian approximation in shared memory: n <= 64")
    _lib.check(lib.b200opt_bfgs_step_init(dtype_code, B, n, Y0.data_ptr(), y.data_ptr(), y_eval.data_ptr(), state.data_ptr(),
                                          steps.data_ptr(), res.data_ptr(), stream), "b200opt_bfgs_step_init")
    hold = {}

    def body():
        g, f = gv(y_eval)
        g, f = g.contiguous(), f.contiguous()
        rn = rootv(y).to(Y0.dtype).contiguous() if rootv is not None else None
        st = t.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.b200opt_bfgs_step(update, use_inverse, dtype_code, B, n, f.data_ptr(), g.data_ptr(),
                                         rn.data_ptr() if rn is not None else None, y.data_ptr(), y_eval.data_ptr(), state.data_ptr(),
                                         C.byref(opts), steps.data_ptr(), res.data_ptr(), active.data_ptr(), st), "b200opt_bfgs_step")
        hold["in"] = (g, f, rn)

    launches = _run_iterations(body, body, active, max_steps_k) if max_steps_k > 0 else 0
    _lib.check(lib.b200opt_bfgs_step_finish(dtype_code, B, n, state.data_ptr(), res.data_ptr(), stream), "b200opt_bfgs_step_finish")
    if throw:
        _throw_if_failed(res)
    stats = {"num_steps": steps[0] if single else steps, "max_steps": (None if max_steps is None or max_steps >= 2**31 - 1 else max_steps), "step_kernel_launches": launches,
             "num_accepted_steps": state[:, n * n + 2 * n + 4].to(t.int32)}
    st = {"f": state[:, n * n + 2 * n], "grad": state[:, n * n : n * n + n]}
    if return_hessian_inv:
        st["hessian_inv" if use_inverse else "hessian"] = state[:, : n * n].reshape(B, n, n)
    kind, ifn = implicit or (("root", root_fn) if root_fn is not None else ("min", fn))
    y = attach_if_needed(adjoint, kind, ifn, y, args, args_in_dims)
    return Solution(value=y[0] if single else y, result=res[0] if single else res, aux=None, stats=stats, state=st)
