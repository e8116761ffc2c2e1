"""ImplicitAdjoint over the CUDA solves (reference: optimistix/_adjoint.py, optimistix/_ad.py:26-118).

The reference differentiates a solution y*(args) with the implicit function theorem: with `rewrite(y, args)` the
optimality residual that vanishes at y* (the root function itself, `_root_find.py:22-26`; the gradient of
0.5 |r|^2, `_least_squares.py:20-28`; the gradient of f, `_minimise.py:20-28`),

    d y* = -(d rewrite / d y)^-1  (d rewrite / d args) d args                       (_ad.py:104-110)

Here the same rule is attached to the (detached) CUDA solution as a `torch.autograd.Function`, in reverse mode:
for a cotangent v of y*, solve (d rewrite / d y)^T w = v with `b200opt_lu_solve` (one warp per problem, LU with
partial pivoting - lx.AutoLinearSolver(well_posed=None) on a square operator) and return -w^T d rewrite / d args,
the latter as a torch VJP of the user's function.  Nothing of the iteration itself is differentiated, exactly as in
the reference.  n <= 32 (the in-warp LU).
"""
from __future__ import annotations

from . import _lib


def _torch():
    import torch

    return torch


def needs_grad(args) -> bool:
    t = _torch()
    from torch.utils._pytree import tree_leaves

    return t.is_grad_enabled() and any(isinstance(x, t.Tensor) and x.requires_grad for x in tree_leaves(args))


def lu_solve(A, b):
    """Batched A x = b on the GPU through the C ABI (A [B, n, n], b [B, n])."""
    t = _torch()
    lib = _lib.load()
    B, n = b.shape
    nmax = 160 if b.dtype == t.float64 else 224
    if n > nmax:
        raise NotImplementedError(f"ImplicitAdjoint: the batched LU solve kernels support n <= {nmax} unknowns in this dtype")
    A = A.contiguous()
    b = b.contiguous()
    x = t.empty_like(b)
    status = t.empty((B,), dtype=t.int32, device=b.device)
    code = _lib.F32 if b.dtype == t.float32 else _lib.F64
    _lib.check(lib.b200opt_lu_solve(code, B, n, A.data_ptr(), b.data_ptr(), x.data_ptr(), status.data_ptr(),
                                    t.cuda.current_stream(b.device).cuda_stream), "b200opt_lu_solve")
    return x, status


def implicit_attach(rewrite, Y, args, args_in_dims):
    """Return Y (batched solution [B, n], detached) as a tensor whose gradient w.r.t. the differentiable leaves of `args`
    follows the implicit function theorem for `rewrite(y, args) == 0`."""
    t = _torch()
    from torch.utils._pytree import tree_flatten, tree_unflatten

    leaves, spec = tree_flatten(args)
    didx = [i for i, x in enumerate(leaves) if isinstance(x, t.Tensor) and x.requires_grad]
    if not didx:
        return Y
    def batched_rewrite(Yb, *dleaves):
        full = list(leaves)
        for i, x in zip(didx, dleaves):
            full[i] = x
        a = tree_unflatten(full, spec)
        if args_in_dims is None:
            return t.func.vmap(lambda y: rewrite(y, a))(Yb)
        return t.func.vmap(rewrite, in_dims=(0, args_in_dims))(Yb, a)

    class _Implicit(t.autograd.Function):
        @staticmethod
        def forward(ctx, Ystar, *dleaves):
            ctx.save_for_backward(Ystar, *dleaves)
            return Ystar.clone()

        @staticmethod
        def backward(ctx, v):
            Ystar, *dleaves = ctx.saved_tensors
            with t.enable_grad():
                dl = [x.detach() for x in dleaves]
                # d rewrite / d y at the solution, per problem
                full = list(leaves)
                for i, x in zip(didx, dl):
                    full[i] = x
                a = tree_unflatten(full, spec)
                if args_in_dims is None:
                    J = t.func.vmap(t.func.jacfwd(lambda y: rewrite(y, a)))(Ystar)
                else:
                    J = t.func.vmap(t.func.jacfwd(rewrite), in_dims=(0, args_in_dims))(Ystar, a)
                w, status = lu_solve(J.transpose(1, 2), v.contiguous())
                # lx.linear_solve inside implicit_jvp (_ad.py:110) raises on a singular operator; a silent NaN gradient
                # would be worse than the synchronisation this check costs
                nbad = int((status != 0).sum().item())
                if nbad:
                    raise RuntimeError(f"ImplicitAdjoint: d(optimality residual)/dy is singular at the solution of {nbad} of "
                                       f"{status.numel()} problems; the implicit gradient does not exist there")
                _, vjp_fn = t.func.vjp(lambda *ls: batched_rewrite(Ystar, *ls), *dl)
                grads = vjp_fn(-w)
            return (None, *grads)

    return _Implicit.apply(Y.detach(), *[leaves[i] for i in didx])


def make_rewrite(kind, fn):
    """The optimality residual of each problem class, as a function of ONE problem (y [n], args)."""
    t = _torch()
    if kind == "root":  # _root_find.py:22-26
        return lambda y, a: fn(y, a).reshape(-1)
    if kind == "lstsq":  # _least_squares.py:20-28
        return lambda y, a: t.func.grad(lambda z: 0.5 * (fn(z, a).reshape(-1) ** 2).sum())(y)
    if kind == "min":  # _minimise.py:20-28
        return lambda y, a: t.func.grad(lambda z: fn(z, a).reshape(()))(y)
    raise ValueError(kind)


def attach_if_needed(adjoint, kind, fn, Y, args, args_in_dims):
    """Called by the generic entry points on the batched solution Y [B, n]."""
    from ._solvers import ImplicitAdjoint

    if not needs_grad(args):
        return Y
    if not isinstance(adjoint, ImplicitAdjoint):
        raise NotImplementedError("only ImplicitAdjoint is available: the iterations run inside CUDA kernels and are not "
                                  "recorded for backpropagation (RecursiveCheckpointAdjoint is out of scope)")
    return implicit_attach(make_rewrite(kind, fn), Y, args, args_in_dims)
