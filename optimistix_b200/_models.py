"""Residual-model plugin registry, host side.

A `RegisteredModel` stands where the reference takes the user callable `fn(y, args)`
(optimistix/_least_squares.py:44-58 etc.).  Passing one to `least_squares` / `root_find` /
`minimise` / `fixed_point` selects the persistent CUDA kernel whose device functions implement
that family (optimistix_b200/csrc/models.cuh); an arbitrary Python callable goes to the generic
per-iteration step kernel instead.  Each model also carries `torch_fn`, the same function
written with torch ops, so that `fn(y, args)` can still be *evaluated* like the reference's fn
(used by tests and by users who want residuals at the solution); it is never used to solve.
"""
from __future__ import annotations

import dataclasses
from typing import Callable

KIND_LSQ, KIND_ROOT, KIND_MIN = 0, 1, 2


@dataclasses.dataclass(frozen=True)
class RegisteredModel:
    name: str
    model_id: int
    kind: int
    n: int  # 0 = any
    arg_names: tuple  # names in `args`, shared ones first
    n_shared: int  # how many of arg_names are shared by all problems
    torch_fn: Callable | None = None
    doc: str = ""

    def __call__(self, y, args):
        if self.torch_fn is None:
            raise NotImplementedError(f"model {self.name} has no host-side evaluator")
        return self.torch_fn(y, self.split_args(args))

    def split_args(self, args) -> dict:
        if args is None:
            args = ()
        if isinstance(args, dict):
            missing = [k for k in self.arg_names if k not in args]
            if missing:
                raise TypeError(f"model {self.name}: args is missing {missing}; expected {self.arg_names}")
            return {k: args[k] for k in self.arg_names}
        if len(args) != len(self.arg_names):
            raise TypeError(f"model {self.name}: expected args {self.arg_names}, got {len(args)} items")
        return dict(zip(self.arg_names, args))


def _exp_decay(y, a):
    import torch

    t, d = a["t"], a["d"]
    return y[..., 0:1] * torch.exp(-y[..., 1:2] * t) + y[..., 2:3] - d


def _gauss2(y, a):
    import torch

    x, d = a["x"], a["d"]
    z1 = (x - y[..., 1:2]) / y[..., 2:3]
    z2 = (x - y[..., 4:5]) / y[..., 5:6]
    return (y[..., 0:1] * torch.exp(-0.5 * z1 * z1) + y[..., 3:4] * torch.exp(-0.5 * z2 * z2)
            + y[..., 6:7] + y[..., 7:8] * x - d)


def _implicit_euler(y, a):
    A, B, y_prev, dt = a["A"], a["B"], a["y_prev"], a["dt"]
    f = y @ A.T + y * (y @ B.T)
    return y - y_prev - dt[..., None] * f


def _rosenbrock(y, a):
    p = y[..., 1:] - y[..., :-1] ** 2
    q = 1.0 - y[..., :-1]
    return (100.0 * p * p + q * q).sum(-1)


exp_decay = RegisteredModel("exp_decay", 0, KIND_LSQ, 3, ("t", "d"), 1, _exp_decay,
                            "r_i = a*exp(-b*t_i) + c - d_i, y = [a, b, c]")
gauss2 = RegisteredModel("gauss2", 1, KIND_LSQ, 8, ("x", "d"), 1, _gauss2,
                         "two Gaussians + linear baseline, y = [A1, mu1, s1, A2, mu2, s2, b0, b1]")
implicit_euler = RegisteredModel("implicit_euler", 2, KIND_ROOT, 32, ("A", "B", "y_prev", "dt"), 2, _implicit_euler,
                                 "F(y) = y - y_prev - dt*(A y + y*(B y))")
rosenbrock = RegisteredModel("rosenbrock", 3, KIND_MIN, 0, (), 0, _rosenbrock, "standard chained Rosenbrock")
linear_rows = RegisteredModel("linear_rows", 4, KIND_LSQ, 0, ("J", "d"), 0, None,
                              "r = J y - d, J [m, n] materialised row-major in HBM, 4 <= n <= 256 (single large problem)")

_REGISTRY = {m.name: m for m in (exp_decay, gauss2, implicit_euler, rosenbrock, linear_rows)}


def register_curve_family(name: str, n: int, source: str, *, naux: int = 1, torch_fn: Callable | None = None,
                          doc: str = "") -> RegisteredModel:
    """Register a least-squares "curve" family at run time (the user side of the plugin registry).

    `source` is CUDA C++: the members `struct Pre`, `prepare`, `residual`, `jac_row` of a model struct exactly like the
    built-in ones in optimistix_b200/csrc/models.cuh (`T` is the scalar type, `N = n`, `NAUX = naux`, `row.x` the abscissa,
    `row.d` the datum).  `b200opt_register_curve_family` compiles it with NVRTC for sm_100a against the library's own kernel
    headers; the returned model then runs `least_squares(model, LevenbergMarquardt | GaussNewton, y0, (x, d))` in the same
    persistent one-warp-per-problem kernel as the built-in families, in fp32 and fp64, on device or host buffers.
    `torch_fn(y, {"x": ..., "d": ...})` is the optional host-side evaluator (tests, ImplicitAdjoint).  Needs a CUDA device."""
    from . import _lib

    if name in _REGISTRY:
        raise ValueError(f"a model family named {name!r} is already registered")
    lib = _lib.load()
    model_id = lib.b200opt_register_curve_family(name.encode(), int(n), int(naux), source.encode(), _lib.CSRC_DIR.encode())
    if model_id < 0:
        _lib.check(model_id, "b200opt_register_curve_family")
    m = RegisteredModel(name, int(model_id), KIND_LSQ, int(n), ("x", "d"), 1, torch_fn, doc or "run-time compiled curve family")
    _REGISTRY[name] = m
    return m


def register_root_family(name: str, n: int, source: str, *, shared_elems: int = 0, torch_fn: Callable | None = None,
                         doc: str = "") -> RegisteredModel:
    """Register a root-finding family at run time: `source` defines the device function
    `eval_row(sh, ys, c, lane, f, row, want_jac)` (lane `lane` evaluates F_lane(y) and row `lane` of the Jacobian; `sh` =
    the `shared_elems` shared arguments staged in shared memory, `ys` = the iterate, `c.prob` = this problem's arguments).
    The returned model runs `root_find(model, Newton | Chord, y0, (shared, prob))` in the persistent one-warp-per-problem
    kernel (in-warp LU with partial pivoting); `shared` is a flat array of `shared_elems` values, `prob` is
    `[batch, anything]`.  n must be a multiple of 4, at most 32.  `torch_fn(y, {"shared": ..., "prob": ...})` is the optional
    host-side evaluator.  Needs a CUDA device."""
    from . import _lib

    if name in _REGISTRY:
        raise ValueError(f"a model family named {name!r} is already registered")
    lib = _lib.load()
    model_id = lib.b200opt_register_root_family(name.encode(), int(n), int(shared_elems), source.encode(), _lib.CSRC_DIR.encode())
    if model_id < 0:
        _lib.check(model_id, "b200opt_register_root_family")
    m = RegisteredModel(name, int(model_id), KIND_ROOT, int(n), ("shared", "prob"), 1, torch_fn, doc or "run-time compiled root family")
    _REGISTRY[name] = m
    return m


def get(name: str) -> RegisteredModel:
    return _REGISTRY[name]


def names():
    return tuple(_REGISTRY)


def check_against_library(lib) -> None:
    """The Python table must describe exactly what is compiled into libb200opt.so."""
    import ctypes as C

    from ._lib import ModelInfo

    assert lib.b200opt_model_count() == len(_REGISTRY)
    for m in _REGISTRY.values():
        info = ModelInfo()
        assert lib.b200opt_model_info_get(m.model_id, C.byref(info)) == 0
        assert info.name.decode() == m.name and info.kind == m.kind and info.n == m.n, (m.name, info.name, info.kind, info.n)
        assert lib.b200opt_model_id(m.name.encode()) == m.model_id
