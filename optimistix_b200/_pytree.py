"""Pytree-valued `y` for the entry points (Python-callable `fn` only).

The reference accepts any pytree of arrays as `y0` (`y0: Y`, optimistix/_least_squares.py:44-58, _minimise.py:37-51,
_root_find.py:119-133, _fixed_point.py:50-66) and ravels it wherever it needs a vector (lineax's `tree_dot` / norms,
optimistix/_misc.py:15-22; its own fixtures do the same by hand, tests/helpers.py:318,331).  The CUDA step kernels work
on flat `[n]` vectors, so the ravel happens once, here, on the host side of the boundary:

    y0 (pytree) --ravel--> Y0 [n] or [B, n] --entry point--> Solution(value [.., n]) --unravel--> Solution(value pytree)

and `fn` is wrapped so that the user's callable still sees its own structure.  Outputs of `fn` may be pytrees too
(residual pytrees for least_squares / root_find, a `y`-shaped pytree for fixed_point, `(out, aux)` with has_aux=True);
they are raveled leaf by leaf in flattening order, which is the order `jax.flatten_util.ravel_pytree` uses.

Batching: the reference is batched from outside with `jax.vmap`; here the batch axis is explicit.  For a pytree `y0`,
`y0_in_dims=None` (default) means one problem, `y0_in_dims=0` means every leaf carries a leading batch axis of the same
length B (the `in_axes=0` of the reference's call site).

Pure host-side bookkeeping: no arithmetic happens here, so there is nothing for the oracle to check beyond the layout
(tests/test_pytree.py).
"""
from __future__ import annotations

import math
from typing import Any, Callable

import numpy as np


def _torch():
    import torch

    return torch


def _tree():
    import torch.utils._pytree as pt

    return pt


def is_array_like(x) -> bool:
    """What the entry points take as a flat `y0` already: tensors, arrays, scalars, and (nested) lists of numbers."""
    if x is None:
        return False
    if type(x).__module__.startswith("torch") and hasattr(x, "dtype"):
        return x.ndim > 0  # a 0-d y is a pytree with one scalar leaf (the flat path wants [n] or [batch, n])
    if isinstance(x, np.ndarray):
        return x.ndim > 0
    if isinstance(x, (np.generic, int, float)):
        return False
    if isinstance(x, (list, tuple)):
        try:
            a = np.asarray(x)
        except Exception:  # noqa: BLE001 - ragged / mixed content: a pytree
            return False
        return a.dtype != object and not any(type(e).__module__.startswith("torch") or isinstance(e, np.ndarray) for e in x)
    return False


class Raveled:
    """The layout of one problem's `y`: tree structure + per-leaf shapes, with ravel / unravel in both batch modes."""

    def __init__(self, y0, batched: bool):
        t, pt = _torch(), _tree()
        leaves, self.spec = pt.tree_flatten(y0)
        if not leaves:
            raise ValueError("y0 has no array leaves")
        dev = None
        for leaf in leaves:
            if type(leaf).__module__.startswith("torch") and leaf.is_cuda:
                dev = leaf.device
                break
        tens = [t.as_tensor(leaf) for leaf in leaves]
        # jnp.result_type over the leaves (_misc.py:281-285 makes every leaf inexact first)
        dtype = None
        for x in tens:
            d = x.dtype if x.dtype.is_floating_point else t.get_default_dtype()
            dtype = d if dtype is None else t.promote_types(dtype, d)
        tens = [x.to(device=dev, dtype=dtype) if dev is not None else x.to(dtype) for x in tens]
        self.batched = bool(batched)
        if self.batched:
            if any(x.ndim == 0 for x in tens):
                raise ValueError("y0_in_dims=0: every leaf of y0 needs a leading batch axis")
            sizes = {int(x.shape[0]) for x in tens}
            if len(sizes) != 1:
                raise ValueError(f"y0_in_dims=0: leaves disagree on the batch size: {sorted(sizes)}")
            self.batch = sizes.pop()
            self.shapes = [tuple(x.shape[1:]) for x in tens]
            flat = [x.reshape(self.batch, -1) for x in tens]
        else:
            self.batch = None
            self.shapes = [tuple(x.shape) for x in tens]
            flat = [x.reshape(-1) for x in tens]
        self.sizes = [math.prod(s) for s in self.shapes]
        self.n = sum(self.sizes)
        self.flat0 = t.cat(flat, dim=-1)

    def unravel(self, v):
        """[.., n] -> pytree with leaves [.., *shape] (works under torch.func transforms: only slicing and reshape)."""
        pt = _tree()
        lead = tuple(v.shape[:-1])
        leaves, o = [], 0
        for shape, size in zip(self.shapes, self.sizes):
            leaves.append(v[..., o:o + size].reshape(lead + shape))
            o += size
        return pt.tree_unflatten(leaves, self.spec)


def ravel_output(out):
    """Ravel a (pytree of) tensor output of `fn` for ONE problem into a 1-D tensor, leaves in flattening order."""
    t, pt = _torch(), _tree()
    if type(out).__module__.startswith("torch") and hasattr(out, "dtype"):
        return out.reshape(-1)
    leaves, _ = pt.tree_flatten(out)
    if not leaves:
        raise ValueError("fn returned no array leaves")
    leaves = [t.as_tensor(x) for x in leaves]
    dtype = leaves[0].dtype
    for x in leaves[1:]:
        dtype = t.promote_types(dtype, x.dtype)
    dev = next((x.device for x in leaves if x.is_cuda), leaves[0].device)
    return t.cat([x.to(device=dev, dtype=dtype).reshape(-1) for x in leaves])


def wrap_fn(fn: Callable, rav: Raveled, kind: str, has_aux: bool) -> Callable:
    """`fn(y_tree, args)` -> `g(y_flat [n], args)` with the output raveled as the entry point `kind` needs it:
    "vector" (least_squares / root_find: residual pytree -> [m]), "scalar" (minimise), "same" (fixed_point: a pytree of
    y's structure -> [n])."""

    def g(y_flat, args):
        out = fn(rav.unravel(y_flat), args)
        aux = None
        if has_aux:
            out, aux = out
        if kind == "scalar":
            if not (type(out).__module__.startswith("torch") and out.numel() == 1):
                raise ValueError("minimise function must output a single scalar.")  # _minimise.py:97-104
            flat = out.reshape(())
        elif kind == "same":
            pt = _tree()
            leaves, spec = pt.tree_flatten(out)
            if spec != rav.spec or [tuple(x.shape) for x in leaves] != rav.shapes:
                raise ValueError("The input and output of `fixed_point_fn` must have the same structure")  # _fixed_point.py:40-46
            flat = ravel_output(out)
        else:
            flat = ravel_output(out)
        return (flat, aux) if has_aux else flat

    return g


def unravel_solution(sol, rav: Raveled):
    from ._solution import Solution

    return Solution(value=rav.unravel(sol.value), result=sol.result, aux=sol.aux, stats=sol.stats, state=sol.state)


def dispatch(entry: Callable, kind: str, fn, solver, y0, args, options, y0_in_dims, has_aux: bool, kw: dict[str, Any]):
    """Run `entry` (one of the four entry points, flat-`y` implementation) on a pytree `y0`."""
    from ._models import RegisteredModel

    if isinstance(fn, RegisteredModel):
        raise TypeError(f"registered model family {fn.name!r} takes a flat parameter vector y0 of shape [n] or [batch, n]")
    if y0_in_dims not in (None, 0):
        raise ValueError("y0_in_dims must be None (one problem) or 0 (leading batch axis on every leaf)")
    if options and (options.get("lower") is not None or options.get("upper") is not None):
        raise NotImplementedError("options['lower'/'upper'] with a pytree y0: pass flat bounds with a flat y0")
    rav = Raveled(y0, batched=y0_in_dims == 0)
    sol = entry(wrap_fn(fn, rav, kind, has_aux), solver, rav.flat0, args, options, has_aux=has_aux, **kw)
    return unravel_solution(sol, rav)
