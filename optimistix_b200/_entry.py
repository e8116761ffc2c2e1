"""Entry points: least_squares / minimise / root_find / fixed_point.

Signatures follow the reference (optimistix/_least_squares.py:44-58, _minimise.py:37-51,
_root_find.py:119-133, _fixed_point.py:50-66):
    f(fn, solver, y0, args=None, options=None, *, has_aux=False, max_steps=256,
      adjoint=ImplicitAdjoint(), throw=True, tags=frozenset()) -> Solution

What differs, by construction: the reference is batched from outside with `jax.vmap`; here the batch
axis is explicit.  `y0` of shape [n] is one problem, `[B, n]` is B independent problems whose
per-problem args carry the same leading axis.  CUDA tensors in -> CUDA tensors out, asynchronously on
torch's current stream (unless `throw=True`, which has to look at the result codes).  NumPy / CPU
tensors in -> NumPy out through the `_host` C entry points (chunked H2D / solve / D2H pipeline).

There is no CPU implementation behind these functions: without the CUDA library they raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, _pytree
from ._models import KIND_LSQ, KIND_MIN, KIND_ROOT, RegisteredModel
from ._solution import RESULTS, Solution
from ._solvers import (
    BFGS,
    DFP,
    AbstractLeastSquaresSolver,
    AbstractMinimiser,
    AbstractRootFinder,
    BacktrackingArmijo,
    Chord,
    ClassicalTrustRegion,
    GaussNewton,
    ImplicitAdjoint,
    LearningRate,
    LevenbergMarquardt,
    Newton,
)

_NP_DTYPES = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64}


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _torch():
    import torch

    return torch


class _Buf:
    """Uniform view of a torch-CUDA tensor or a NumPy array as (pointer, keepalive)."""

    def __init__(self, device: bool, dtype_code: int):
        self.device = device
        self.dtype_code = dtype_code
        if device:
            t = _torch()
            self.tdtype = t.float32 if dtype_code == _lib.F32 else t.float64
        self.ndtype = np.float32 if dtype_code == _lib.F32 else np.float64

    def arr(self, x, dev=None):
        if self.device:
            t = _torch()
            x = t.as_tensor(x, dtype=self.tdtype, device=dev)
            return x.contiguous()
        if _is_torch(x):
            x = x.detach().cpu().numpy()
        return np.ascontiguousarray(x, dtype=self.ndtype)

    def empty(self, shape, like, dtype=None):
        t = _torch()
        if self.device:
            return t.empty(shape, dtype=dtype or self.tdtype, device=like.device)
        return _host_empty(shape, dtype or self.ndtype)

    def empty_i32(self, shape, like):
        t = _torch()
        if self.device:
            return t.empty(shape, dtype=t.int32, device=like.device)
        return _host_empty(shape, np.int32)

    @staticmethod
    def ptr(x):
        if x is None:
            return None
        if _is_torch(x):
            return x.data_ptr()
        return x.ctypes.data


def _data16_kind(x):
    """16-bit per-problem data (float16 / uint16 / int16) is handed to the kernels as it is (B200OPT_DATA_*): half the bytes per
    problem; every value is converted exactly to the compute dtype when the row is loaded."""
    name = str(getattr(x, "dtype", "")).replace("torch.", "")
    return {"float16": _lib.DATA_F16, "uint16": _lib.DATA_U16, "int16": _lib.DATA_I16}.get(name, _lib.DATA_AS_COMPUTE)


def _np_ndim(x):
    return x.ndim if hasattr(x, "ndim") else np.asarray(x).ndim


def _as_cuda(x):
    t = _torch()
    return x if (_is_torch(x) and x.is_cuda) else t.as_tensor(x).cuda()


def _host_empty(shape, dtype):
    """Host output / staging array.  Page-locked through torch's caching host allocator when a CUDA device is present:
    the D2H copies of the C host pipeline then land directly in it (no bounce buffer, no first-touch page faults on a
    fresh pageable array - those cost more than the solve itself for the small-problem workloads)."""
    t = _torch()
    if t.cuda.is_available():
        try:
            tdt = {np.dtype(np.float32): t.float32, np.dtype(np.float64): t.float64, np.dtype(np.int32): t.int32}[np.dtype(dtype)]
            return t.empty(shape, dtype=tdt, pin_memory=True).numpy()
        except Exception:  # noqa: BLE001 - out of lockable memory: a pageable array works too (the pipeline bounces)
            pass
    return np.empty(shape, dtype=dtype)


def _prepare_y0(y0):
    if _is_torch(y0):
        t = _torch()
        if not y0.dtype.is_floating_point:
            y0 = y0.to(t.get_default_dtype())  # inexact_asarray, _misc.py:281-285
        if y0.dtype not in (t.float32, t.float64):
            raise TypeError(f"unsupported dtype {y0.dtype}: the CUDA path computes in float32 or float64")
        code = _lib.F32 if y0.dtype == t.float32 else _lib.F64
        device = y0.is_cuda
    else:
        y0 = np.asarray(y0)
        if not np.issubdtype(y0.dtype, np.floating):
            y0 = y0.astype(np.float64)
        if y0.dtype not in _NP_DTYPES:
            raise TypeError(f"unsupported dtype {y0.dtype}")
        code = _NP_DTYPES[y0.dtype]
        device = False
    single = y0.ndim == 1
    if y0.ndim not in (1, 2):
        raise ValueError("y0 must have shape [n] (one problem) or [batch, n]")
    buf = _Buf(device, code)
    y0 = buf.arr(y0, dev=y0.device if device else None)
    if single:
        y0 = y0[None]
    return y0, buf, single


def _finish(buf, single, throw, value, num_steps, result, max_steps, extra_stats=None, state=None):
    if throw:
        # reference: result.error_if(sol, result != successful) (_iterate.py:357-358) - a runtime error if
        # ANY lane failed.  Needs the codes on the host, so this synchronises.
        res_host = result.cpu().numpy() if _is_torch(result) else result
        bad = np.flatnonzero(res_host != RESULTS.successful)
        if bad.size:
            code = int(res_host[bad[0]])
            raise RuntimeError(
                f"{bad.size} of {res_host.size} solves failed; first failure (problem {int(bad[0])}): "
                f"{RESULTS.name(code)}: {RESULTS[code]}"
            )
    stats = {"num_steps": num_steps, "max_steps": (None if max_steps is None or max_steps >= 2**31 - 1 else max_steps)}
    if extra_stats:
        stats.update(extra_stats)
    if single:
        value, result = value[0], result[0]
        stats = {k: (v[0] if hasattr(v, "__len__") else v) for k, v in stats.items()}
    return Solution(value=value, result=result, aux=None, stats=stats, state=state)


def _attach_registered(adjoint, kind, model, a, y_out, buf):
    """ImplicitAdjoint for a registered family: the persistent kernel produced y_out; its gradient w.r.t. the tensors in
    `args` follows from the family's torch evaluator (`RegisteredModel.torch_fn`) by the implicit function theorem
    (_adjoint.py).  Device tensors only; y0 never carries a gradient (the solution does not depend on it, _ad.py)."""
    if not buf.device or model.torch_fn is None:
        return y_out
    from ._adjoint import attach_if_needed, needs_grad

    if not needs_grad(a):
        return y_out
    in_dims = {k: (None if i < model.n_shared else 0) for i, k in enumerate(model.arg_names)}
    return attach_if_needed(adjoint, kind, lambda y, ad: model.torch_fn(y, ad), y_out, a, in_dims)


def _on_device_of(buf, like):
    """Kernels launch on the CURRENT device: make it the one that holds the buffers (a caller with tensors on cuda:1 need
    not have switched devices, exactly as with torch's own operators)."""
    import contextlib

    if not buf.device:
        return contextlib.nullcontext()
    return _torch().cuda.device(like.device)


def _stream_ptr(buf, like):
    if not buf.device:
        return None
    return _torch().cuda.current_stream(like.device).cuda_stream


_UNBOUNDED = 2**31 - 1  # max_steps=None (_iterate.py:194-268 with no bound): the kernels count steps in int32


def _solve_with_aux(entry, fn, solver, y0, args, options, *, args_in_dims=None, **kw):
    """has_aux=True (reference: fn returns (out, aux), Solution.aux carries the auxiliary output): solve with the first
    output, then evaluate the auxiliary output once more at the returned value."""
    if isinstance(fn, RegisteredModel):
        raise NotImplementedError("has_aux=True: registered model families return no auxiliary output")
    t = _torch()
    sol = entry(lambda y, a: fn(y, a)[0], solver, y0, args, options, has_aux=False, args_in_dims=args_in_dims, **kw)
    v = sol.value.detach()
    if v.ndim == 1:
        aux = fn(v, args)[1]
    elif v.shape[0] == 0:
        aux = None
    elif args_in_dims is None:
        aux = t.func.vmap(lambda y: fn(y, args)[1])(v)
    else:
        aux = t.func.vmap(lambda y, a: fn(y, a)[1], in_dims=(0, args_in_dims))(v, args)
    return Solution(value=sol.value, result=sol.result, aux=aux, stats=sol.stats, state=sol.state)


def _check_common(fn, options, has_aux, adjoint):
    if has_aux:
        raise NotImplementedError("has_aux=True: registered model families return no auxiliary output")
    if options:
        unknown = set(options) - {"jac", "autodiff_mode", "lower", "upper"}
        if unknown:
            raise ValueError(f"unknown options {sorted(unknown)}")
        if options.get("jac", "fwd") not in ("fwd", "bwd"):
            raise ValueError("Only `jac='fwd'` or `jac='bwd'` are valid.")  # gauss_newton.py:180
        if options.get("autodiff_mode", "bwd") not in ("fwd", "bwd"):
            raise ValueError("Only `autodiff_mode='fwd'` or `autodiff_mode='bwd'` are valid.")  # _misc.py:255-258


def _check_defaults(solver):
    search = getattr(solver, "search", None)
    if isinstance(solver, LevenbergMarquardt) and search != ClassicalTrustRegion():
        raise NotImplementedError("the CUDA LM kernel implements ClassicalTrustRegion with the reference defaults")
    if isinstance(solver, GaussNewton) and search != LearningRate(1.0):
        raise NotImplementedError("the CUDA GN kernel implements LearningRate(1.0)")
    if isinstance(solver, (BFGS, DFP)) and search != BacktrackingArmijo():
        raise NotImplementedError("the CUDA quasi-Newton kernels implement BacktrackingArmijo with the reference defaults")


# ------------------------------------------------------------------------------------------ least squares
def least_squares(fn, solver, y0, args=None, options=None, *, has_aux=False, max_steps: int | None = 256,
                  adjoint=ImplicitAdjoint(), throw: bool = True, tags=frozenset(), _root_near_zero=False, chunk: int = 0,
                  args_in_dims=None, y0_in_dims=None):
    """Solve (a batch of) nonlinear least-squares problems. Reference: optimistix/_least_squares.py:44-135."""
    if y0_in_dims is not None or not _pytree.is_array_like(y0):  # pytree-valued y (and residuals): ravel on the host side
        return _pytree.dispatch(least_squares, "vector", fn, solver, y0, args, options, y0_in_dims, has_aux,
                                dict(max_steps=max_steps, adjoint=adjoint, throw=throw, tags=tags, _root_near_zero=_root_near_zero,
                                     chunk=chunk, args_in_dims=args_in_dims))
    if has_aux:
        return _solve_with_aux(least_squares, fn, solver, y0, args, options, max_steps=max_steps, adjoint=adjoint, throw=throw,
                               tags=tags, _root_near_zero=_root_near_zero, chunk=chunk, args_in_dims=args_in_dims)
    if isinstance(solver, AbstractMinimiser):
        # _least_squares.py:36-41,102-114: minimise 0.5 * sum(residual^2)
        if isinstance(fn, RegisteredModel):
            raise NotImplementedError("least_squares through a minimiser takes a Python callable (generic path)")
        _check_common(fn, options, has_aux, adjoint)
        return minimise(lambda y, a: 0.5 * (fn(y, a).reshape(-1) ** 2).sum(), solver, y0, args, options, max_steps=max_steps,
                        adjoint=adjoint, throw=throw, args_in_dims=args_in_dims, _implicit=("lstsq", fn))
    if not isinstance(solver, (LevenbergMarquardt, GaussNewton)):
        raise TypeError(f"least_squares: unsupported solver {type(solver).__name__}")
    _check_common(fn, options, has_aux, adjoint)
    _check_defaults(solver)
    if max_steps is None:
        max_steps = _UNBOUNDED
    if isinstance(fn, RegisteredModel) and getattr(solver, "verbose", None) is not None and fn.name != "linear_rows":
        # a verbose solve reports every step to the host: it runs through the per-iteration step kernel, with the family's
        # torch evaluator standing in for the device functions
        if fn.torch_fn is None:
            raise NotImplementedError(f"verbose: model family {fn.name} has no torch evaluator")
        tfn, names, nshared = fn.torch_fn, fn.arg_names, fn.n_shared
        a0 = fn.split_args(args)
        in_dims = {k: (None if i < nshared else 0) for i, k in enumerate(names)} if _np_ndim(y0) == 2 else None
        return least_squares(lambda y, ad: tfn(y, ad), solver, y0, {k: _as_cuda(a0[k]) for k in names}, options, max_steps=max_steps,
                             adjoint=adjoint, throw=throw, _root_near_zero=_root_near_zero, args_in_dims=in_dims)
    if not isinstance(fn, RegisteredModel):
        from ._generic import least_squares_generic

        return least_squares_generic(fn, solver, y0, args, max_steps=max_steps, throw=throw, root_near_zero=_root_near_zero,
                                     options=options, args_in_dims=args_in_dims, adjoint=adjoint)
    if fn.kind != KIND_LSQ:
        raise TypeError(f"model {fn.name} is not a least-squares family")
    if fn.name == "linear_rows":
        # one large row-sharded problem: tensor-core Gram + (optional) NCCL all-reduce (BASELINE configs[4])
        if not isinstance(solver, GaussNewton):
            raise NotImplementedError("linear_rows runs under GaussNewton (BASELINE configs[4])")
        from ._gn_large import gauss_newton_large

        a = fn.split_args(args)
        return gauss_newton_large(a["J"], a["d"], y0, solver, max_steps=max_steps, throw=throw)
    lib = _lib.load()
    y0, buf, single = _prepare_y0(y0)
    B, n = y0.shape
    a = fn.split_args(args)
    dev = y0.device if buf.device else None
    shared = buf.arr(a[fn.arg_names[0]], dev).reshape(-1)
    data_kind = _data16_kind(a[fn.arg_names[1]])
    if data_kind != _lib.DATA_AS_COMPUTE:
        prob = a[fn.arg_names[1]]
        if buf.device:
            t = _torch()
            prob = (prob if _is_torch(prob) else t.as_tensor(prob)).to(dev).contiguous()
        else:
            prob = np.ascontiguousarray(prob.detach().cpu().numpy() if _is_torch(prob) else prob)
    else:
        prob = buf.arr(a[fn.arg_names[1]], dev)
    if single and prob.ndim == 1:
        prob = prob[None]
    m = int(shared.shape[0])
    if prob.ndim != 2 or prob.shape[0] != B or prob.shape[1] != m:
        raise ValueError(f"{fn.name}: per-problem data must have shape [{B}, {m}], got {tuple(prob.shape)}")
    solver_id = _lib.SOLVER_LM if isinstance(solver, LevenbergMarquardt) else _lib.SOLVER_GN
    opts = _lib.make_options(solver.rtol, solver.atol, solver.norm, max_steps, root_near_zero=_root_near_zero,
                             linear_solver=solver.descent.linear_solver, data_dtype=data_kind)
    y_out = buf.empty((B, n), y0)
    steps = buf.empty_i32((B,), y0)
    res = buf.empty_i32((B,), y0)
    extra = {}
    if buf.device:
        step_size = buf.empty((B,), y0)
        nacc = buf.empty_i32((B,), y0)
        scratch = buf.empty_i32((2,), y0)  # options.scratch: the launch's problem-queue counter (caching allocator, stream-ordered reuse)
        opts.scratch = scratch.data_ptr()
        with _on_device_of(buf, y0):
            code = lib.b200opt_least_squares(solver_id, fn.model_id, buf.dtype_code, B, n, m, buf.ptr(y0), buf.ptr(shared),
                                             buf.ptr(prob), m, C.byref(opts), buf.ptr(y_out), buf.ptr(steps), buf.ptr(res),
                                             buf.ptr(step_size), buf.ptr(nacc), _stream_ptr(buf, y0))
        extra = {"step_size": step_size, "num_accepted_steps": nacc}
    else:
        t = _torch()
        device = t.cuda.current_device() if t.cuda.is_available() else 0
        code = lib.b200opt_least_squares_host(device, solver_id, fn.model_id, buf.dtype_code, B, n, m, buf.ptr(y0),
                                              buf.ptr(shared), m, buf.ptr(prob), m, C.byref(opts), buf.ptr(y_out),
                                              buf.ptr(steps), buf.ptr(res), int(chunk))
    _lib.check(code, "b200opt_least_squares")
    y_out = _attach_registered(adjoint, "root" if _root_near_zero else "lstsq", fn, a, y_out, buf)
    return _finish(buf, single, throw, y_out, steps, res, max_steps, extra)


# ------------------------------------------------------------------------------------------ root find
def root_find(fn, solver, y0, args=None, options=None, *, has_aux=False, max_steps: int | None = 256,
              adjoint=ImplicitAdjoint(), throw: bool = True, tags=frozenset(), chunk: int = 0, args_in_dims=None, y0_in_dims=None):
    """Solve (a batch of) root-finding problems. Reference: optimistix/_root_find.py:119-241."""
    if y0_in_dims is not None or not _pytree.is_array_like(y0):
        return _pytree.dispatch(root_find, "vector", fn, solver, y0, args, options, y0_in_dims, has_aux,
                                dict(max_steps=max_steps, adjoint=adjoint, throw=throw, tags=tags, chunk=chunk, args_in_dims=args_in_dims))
    if has_aux:
        return _solve_with_aux(root_find, fn, solver, y0, args, options, max_steps=max_steps, adjoint=adjoint, throw=throw,
                               tags=tags, chunk=chunk, args_in_dims=args_in_dims)
    _check_common(fn, options, has_aux, adjoint)
    if max_steps is None:
        max_steps = _UNBOUNDED
    if isinstance(solver, AbstractLeastSquaresSolver):
        # _root_find.py:200-225: least-squares solver wrapped in _LstsqToRoot (extra `norm(f) < atol` test)
        return least_squares(fn, solver, y0, args, options, max_steps=max_steps, adjoint=adjoint, throw=throw, _root_near_zero=True,
                             chunk=chunk, args_in_dims=args_in_dims)
    if isinstance(solver, AbstractMinimiser):
        # _root_find.py:39-41,45-80,176-199: minimise norm(fn), terminate only once norm(fn) < atol as well
        if isinstance(fn, RegisteredModel):
            raise NotImplementedError("root_find through a minimiser takes a Python callable (generic path)")
        _check_defaults(solver)
        from ._generic import minimise_generic

        norm = solver.norm
        return minimise_generic(lambda y, a: norm(fn(y, a).reshape(-1)), solver, y0, args, max_steps=max_steps,
                                throw=throw, args_in_dims=args_in_dims, root_fn=fn, adjoint=adjoint)
    if not isinstance(solver, (Newton, Chord)):
        raise TypeError(f"root_find: unsupported solver {type(solver).__name__}")
    if not isinstance(fn, RegisteredModel):
        from ._generic import root_find_generic

        return root_find_generic(fn, solver, y0, args, options, max_steps=max_steps, throw=throw, args_in_dims=args_in_dims,
                                 adjoint=adjoint)
    if fn.kind != KIND_ROOT:
        raise TypeError(f"model {fn.name} is not a root-finding family")
    lib = _lib.load()
    y0, buf, single = _prepare_y0(y0)
    B, n = y0.shape
    a = fn.split_args(args)
    dev = y0.device if buf.device else None
    shared, prob, stride = _pack_root_args(fn, a, buf, dev, B, n, single)
    lower = upper = None
    if options:
        if options.get("lower") is not None:
            lower = _bound(buf, options["lower"], n, dev)
        if options.get("upper") is not None:
            upper = _bound(buf, options["upper"], n, dev)
    solver_id = _lib.SOLVER_NEWTON if isinstance(solver, Newton) else _lib.SOLVER_CHORD
    opts = _lib.make_options(solver.rtol, solver.atol, solver.norm, max_steps,
                             cauchy_termination=solver.cauchy_termination, kappa=solver.kappa)
    y_out = buf.empty((B, n), y0)
    steps = buf.empty_i32((B,), y0)
    res = buf.empty_i32((B,), y0)
    if buf.device:
        with _on_device_of(buf, y0):
            code = lib.b200opt_root_find(solver_id, fn.model_id, buf.dtype_code, B, n, buf.ptr(y0), buf.ptr(shared), buf.ptr(prob),
                                         stride, buf.ptr(lower), buf.ptr(upper), C.byref(opts), buf.ptr(y_out), buf.ptr(steps),
                                         buf.ptr(res), _stream_ptr(buf, y0))
    else:
        if lower is not None or upper is not None:
            raise NotImplementedError("lower/upper bounds are only wired on the device-pointer entry point")
        t = _torch()
        device = t.cuda.current_device() if t.cuda.is_available() else 0
        if isinstance(prob, _HostParts):
            k = len(prob.arrays)
            ptrs = (C.c_void_p * k)(*[buf.ptr(x) for x in prob.arrays])
            elems = (C.c_int32 * k)(*prob.elems)
            code = lib.b200opt_root_find_host_parts(device, solver_id, fn.model_id, buf.dtype_code, B, n, buf.ptr(y0), buf.ptr(shared),
                                                    int(shared.shape[0]), ptrs, elems, k, stride, C.byref(opts), buf.ptr(y_out),
                                                    buf.ptr(steps), buf.ptr(res), int(chunk))
        else:
            code = lib.b200opt_root_find_host(device, solver_id, fn.model_id, buf.dtype_code, B, n, buf.ptr(y0), buf.ptr(shared),
                                              int(shared.shape[0]), buf.ptr(prob), stride, C.byref(opts), buf.ptr(y_out),
                                              buf.ptr(steps), buf.ptr(res), int(chunk))
    _lib.check(code, "b200opt_root_find")
    y_out = _attach_registered(adjoint, "root", fn, a, y_out, buf)
    return _finish(buf, single, throw, y_out, steps, res, max_steps)


def _bound(buf, b, n, dev):
    if buf.device:
        t = _torch()
        return t.broadcast_to(t.as_tensor(b, dtype=buf.tdtype, device=dev), (n,)).contiguous()
    return np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=buf.ndtype), (n,)))


class _HostParts:
    """Per-problem arguments as separate host arrays (part p: [B, elems[p]]) for b200opt_root_find_host_parts."""

    def __init__(self, arrays, elems):
        self.arrays, self.elems = arrays, elems


def _pack_root_args(fn, a, buf, dev, B, n, single):
    """implicit_euler: shared = [A | B] (2 n^2), per problem = [y_prev (n) | dt | pad] with an even stride.
    Run-time families (models.register_root_family): args = (shared [S], prob [B, P]) are passed through."""
    if fn.model_id >= 1000:
        prob = buf.arr(a["prob"], dev)
        if single and prob.ndim == 1:
            prob = prob[None]
        if prob.ndim != 2 or prob.shape[0] != B:
            raise ValueError(f"{fn.name}: per-problem args must have shape [{B}, P], got {tuple(prob.shape)}")
        sh = a["shared"]
        if sh is None:
            sh = [0.0]  # the C entry point wants a non-NULL pointer even when the family stages nothing
        shared = buf.arr(sh, dev).reshape(-1)
        return shared, prob, int(prob.shape[1])
    if fn.name != "implicit_euler":
        raise NotImplementedError(fn.name)
    A, Bm = buf.arr(a["A"], dev), buf.arr(a["B"], dev)
    if tuple(A.shape) != (n, n) or tuple(Bm.shape) != (n, n):
        raise ValueError(f"implicit_euler: A and B must be [{n}, {n}]")
    y_prev, dt = buf.arr(a["y_prev"], dev), buf.arr(a["dt"], dev)
    if single:
        y_prev, dt = y_prev.reshape(1, n), dt.reshape(1)
    if tuple(y_prev.shape) != (B, n) or tuple(dt.shape) != (B,):
        raise ValueError(f"implicit_euler: y_prev must be [{B}, {n}] and dt [{B}]")
    stride = n + 2
    if buf.device:
        t = _torch()
        shared = t.cat([A.reshape(-1), Bm.reshape(-1)])
        prob = t.empty((B, stride), dtype=buf.tdtype, device=dev)
        prob[:, :n] = y_prev
        prob[:, n] = dt
        prob[:, n + 1] = 0
    else:
        # host buffers: the two per-problem arrays go to the C pipeline as they are (b200opt_root_find_host_parts lays them out
        # as [y_prev | dt | pad] rows on the device); packing them here was 71 MB of host copies per call for configs[2]
        shared = np.concatenate([A.reshape(-1), Bm.reshape(-1)])
        prob = _HostParts([np.ascontiguousarray(y_prev), np.ascontiguousarray(dt)], [n, 1])
    return shared, prob, stride


# ------------------------------------------------------------------------------------------ fixed point
def fixed_point(fn, solver, y0, args=None, options=None, *, has_aux=False, max_steps: int | None = 256,
                adjoint=ImplicitAdjoint(), throw: bool = True, tags=frozenset(), args_in_dims=None, y0_in_dims=None):
    """Find a fixed point. Reference: optimistix/_fixed_point.py:50-154: with a root finder / least-squares
    solver / minimiser this is `root_find` on `fn(y, args) - y` (:35-47, :115-129)."""
    if (y0_in_dims is not None or not _pytree.is_array_like(y0)) and not isinstance(fn, RegisteredModel):
        return _pytree.dispatch(fixed_point, "same", fn, solver, y0, args, options, y0_in_dims, has_aux,
                                dict(max_steps=max_steps, adjoint=adjoint, throw=throw, tags=tags, args_in_dims=args_in_dims))
    if isinstance(fn, RegisteredModel):
        raise TypeError("registered model families are residual/root/objective functions, not fixed-point maps; "
                        "use root_find, or pass a Python callable")
    if not isinstance(solver, (AbstractRootFinder, AbstractLeastSquaresSolver, AbstractMinimiser)):
        raise NotImplementedError("FixedPointIteration is out of scope (SURVEY section 2); use Newton/Chord/LM")

    def _to_root_fn(y, a):  # _fixed_point.py:35-47
        out = fn(y, a)
        aux = None
        if has_aux:
            out, aux = out
        if out.shape != y.shape:
            raise ValueError("The input and output of `fixed_point_fn` must have the same structure")
        return (out - y, aux) if has_aux else out - y

    return root_find(_to_root_fn, solver, y0, args, options, has_aux=has_aux, max_steps=max_steps, adjoint=adjoint, throw=throw,
                     args_in_dims=args_in_dims)


# ------------------------------------------------------------------------------------------ minimise
def minimise(fn, solver, y0, args=None, options=None, *, has_aux=False, max_steps: int | None = 256,
             adjoint=ImplicitAdjoint(), throw: bool = True, tags=frozenset(), chunk: int = 0, return_hessian_inv=False,
             args_in_dims=None, _implicit=None, y0_in_dims=None):
    """Minimise (a batch of) scalar functions. Reference: optimistix/_minimise.py:37-119."""
    if y0_in_dims is not None or not _pytree.is_array_like(y0):
        return _pytree.dispatch(minimise, "scalar", fn, solver, y0, args, options, y0_in_dims, has_aux,
                                dict(max_steps=max_steps, adjoint=adjoint, throw=throw, tags=tags, chunk=chunk,
                                     return_hessian_inv=return_hessian_inv, args_in_dims=args_in_dims, _implicit=_implicit))
    if has_aux:
        return _solve_with_aux(minimise, fn, solver, y0, args, options, max_steps=max_steps, adjoint=adjoint, throw=throw, tags=tags,
                               chunk=chunk, return_hessian_inv=return_hessian_inv, args_in_dims=args_in_dims, _implicit=_implicit)
    _check_common(fn, options, has_aux, adjoint)
    if not isinstance(solver, (BFGS, DFP)):
        raise TypeError(f"minimise: unsupported solver {type(solver).__name__}")
    _check_defaults(solver)
    if not isinstance(fn, RegisteredModel):
        from ._generic import minimise_generic

        return minimise_generic(fn, solver, y0, args, max_steps=max_steps, throw=throw, args_in_dims=args_in_dims,
                                return_hessian_inv=return_hessian_inv, adjoint=adjoint, implicit=_implicit,
                                autodiff_mode=(options or {}).get("autodiff_mode", "bwd"))
    if fn.kind != KIND_MIN:
        raise TypeError(f"model {fn.name} is not a minimisation family")
    if isinstance(solver, DFP) or not solver.use_inverse:
        raise NotImplementedError("registered minimisation families run BFGS(use_inverse=True) in the persistent kernel; "
                                  "DFP / use_inverse=False take a Python callable (generic step kernel)")
    if max_steps is None:
        max_steps = _UNBOUNDED
    lib = _lib.load()
    y0, buf, single = _prepare_y0(y0)
    B, n = y0.shape
    opts = _lib.make_options(solver.rtol, solver.atol, solver.norm, max_steps)
    y_out = buf.empty((B, n), y0)
    steps = buf.empty_i32((B,), y0)
    res = buf.empty_i32((B,), y0)
    extra, state = {}, None
    if buf.device:
        f_out = buf.empty((B,), y0)
        nacc = buf.empty_i32((B,), y0)
        hinv = buf.empty((B, n, n), y0) if return_hessian_inv else None
        scratch = buf.empty_i32((2,), y0)  # options.scratch (see least_squares)
        opts.scratch = scratch.data_ptr()
        with _on_device_of(buf, y0):
            code = lib.b200opt_minimise(_lib.SOLVER_BFGS, fn.model_id, buf.dtype_code, B, n, buf.ptr(y0), None, None, 0,
                                        C.byref(opts), buf.ptr(y_out), buf.ptr(steps), buf.ptr(res), buf.ptr(f_out),
                                        buf.ptr(hinv), buf.ptr(nacc), _stream_ptr(buf, y0))
        extra = {"num_accepted_steps": nacc}
        state = {"f": f_out, "hessian_inv": hinv}
    else:
        t = _torch()
        device = t.cuda.current_device() if t.cuda.is_available() else 0
        code = lib.b200opt_minimise_host(device, _lib.SOLVER_BFGS, fn.model_id, buf.dtype_code, B, n, buf.ptr(y0), None, 0,
                                         None, 0, C.byref(opts), buf.ptr(y_out), buf.ptr(steps), buf.ptr(res), int(chunk))
    _lib.check(code, "b200opt_minimise")
    return _finish(buf, single, throw, y_out, steps, res, max_steps, extra, state)
