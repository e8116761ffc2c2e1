"""The reference's own test problems (tests/helpers.py of patrick-kidger/optimistix), restated.

Each problem is a plain function of a flat 1-D torch tensor ``y`` (the reference ravel()s its
pytrees too, helpers.py:318,331) so that the residual / Jacobian / gradient can be produced
by ``torch.func`` on the CPU for the oracle and on the GPU for the generic step-kernel path.
Cites are to /root/reference/tests/helpers.py unless noted.  ``jr.normal(PRNGKey(17))`` data
(helpers.py:477-480,535-538,789-790) is replaced by a NumPy seed, as SURVEY.md section 8c allows.
"""
from __future__ import annotations

import numpy as np
import torch

_rng = np.random.default_rng(17)


def batched_lsq(fn):
    """fn(y[n], args) -> r[m]  ==>  model(Y[B,n], args) -> (r[B,m], J[B,m,n]) in NumPy."""

    def model(Y, args):
        Yt = torch.as_tensor(np.ascontiguousarray(Y))
        f = lambda y: fn(y, args)
        r = torch.func.vmap(f)(Yt)
        J = torch.func.vmap(torch.func.jacfwd(f))(Yt)
        return r.numpy().astype(Y.dtype), J.numpy().astype(Y.dtype)

    return model


def batched_min(fn):
    """fn(y[n], args) -> scalar  ==>  model(Y, args) -> (f[B], grad[B,n])."""

    def model(Y, args):
        Yt = torch.as_tensor(np.ascontiguousarray(Y))
        f = lambda y: fn(y, args)
        val = torch.func.vmap(f)(Yt)
        g = torch.func.vmap(torch.func.grad(f))(Yt)
        return val.numpy().astype(Y.dtype), g.numpy().astype(Y.dtype)

    return model


def T(x):
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


# ---------------------------------------------------------------- least squares (helpers.py:486-505)
def diagonal_quadratic_bowl(y, w):  # :323-326
    return y * y * (0.1 + w)


def linear_rosenbrock(y, const):  # :329-334  (NOT the Rosenbrock function - SURVEY F7)
    return torch.cat([100.0 * (y[1:] - y[:-1]), const - y[:-1]])


LEAST_SQUARES = [
    # (name, fn, minimum of sum r^2, y0, args)
    ("diagonal_quadratic_bowl", diagonal_quadratic_bowl, 0.0, 0.05 * np.ones(20), T(_rng.standard_normal(20) ** 2)),
    ("rosenbrock_2", linear_rosenbrock, 0.0, np.array([1.5, 1.5]), T(1.0)),
    ("rosenbrock_20", linear_rosenbrock, 0.0, 1.45 * np.ones(20), T(1.0)),
]


# ---------------------------------------------------------------- minimisation (helpers.py:540-562)
def bowl(y, matrix):  # :316-320
    return y @ matrix @ y


def himmelblau(y, args):  # :337-343
    c1, c2 = args
    return (y[0] ** 2 + y[1] - c1) ** 2 + (y[0] + y[1] ** 2 - c2) ** 2


def matyas(y, args):  # :346-352
    c1, c2 = args
    return c1 * (y[0] ** 2 + y[1] ** 2) - c2 * y[0] * y[1]


def beale(y, args):  # :365-372
    c1, c2, c3 = args
    a, b = y[0], y[1]
    return (c1 - a + a * b) ** 2 + (c2 - a + a * b**2) ** 2 + (c3 - a + a * b**3) ** 2


def square_minus_one(y, args):  # :413-415
    return torch.sum(y * y) - 1.0


_M = _rng.standard_normal((20, 20))
MINIMISATION = [
    ("bowl", bowl, 0.0, 0.05 * np.ones(20), T(_M.T @ _M)),
    ("himmelblau", himmelblau, 0.0, np.array([2.0, 2.5]), (11.0, 7.0)),
    ("matyas", matyas, 0.0, np.array([6.0, 6.0]), (0.26, 0.48)),
    ("beale", beale, 0.0, np.array([2.0, 0.0]), (1.5, 2.25, 2.625)),
    ("square_minus_one", square_minus_one, -1.0, np.array([1.0]), None),
]


# ---------------------------------------------------------------- fixed point / root find (helpers.py:840-871)
def _cos(y, c):  # :679-681
    return torch.cos(c * y)


def _exponential(y, c):  # :674-676
    return torch.exp(-c * y)


def robertson(y, k):  # :658-670
    k1, k2, k3 = k
    f0 = -k1 * y[0] + k3 * y[1] * y[2]
    f1 = k1 * y[0] - k2 * y[1] ** 2 - k3 * y[1] * y[2]
    f2 = k2 * y[1] ** 2
    return torch.stack([f0, f1, f2])


def _midpoint_y(f, dt):  # :605-615 with y0 = 0
    return lambda y, a: dt * f(0.5 * y, a)


def _midpoint_f(f, dt):  # :618-628
    return lambda y, a: f(dt * y, a)


def _midpoint_k(f, dt):  # :631-645
    return lambda y, a: f(y, a) * dt


_lin = lambda x, M: M @ x
_ROB = (0.04, 3e7, 1e4)
_MAT = T(np.random.default_rng(19).standard_normal((28, 28)))
_ones28 = 0.5 * np.ones(28)
FIXED_POINT = [
    # (name, fn (fixed-point form), y0, args)
    ("cos", _cos, _ones28, T(1.0)),
    ("exponential", _exponential, _ones28, T(1.0)),
    ("midpoint_y_linear", _midpoint_y(_lin, 1 / 16), _ones28, _MAT),
    ("midpoint_f_linear", _midpoint_f(_lin, 1 / 16), _ones28, _MAT),
    ("midpoint_k_linear", _midpoint_k(_lin, 1 / 16), _ones28, _MAT),
    ("midpoint_y_robertson", _midpoint_y(robertson, 1e-8), np.ones(3), _ROB),
    ("midpoint_f_robertson", _midpoint_f(robertson, 1e-8), np.ones(3), _ROB),
    ("midpoint_k_robertson", _midpoint_k(robertson, 1e-8), np.ones(3), _ROB),
]


def as_root(fn):
    """fixed_point -> root_find conversion, _fixed_point.py:35-47 / test_root_find.py:38-40."""
    return lambda y, a: fn(y, a) - y


# tests/test_root_find.py:160-193
def newton_bounded(y, _):
    ya, yb, yc, yd, ye = y[0], y[1], y[2], y[3], y[4]
    foo = ya + yb + yc + yd + ye
    bar = ya - yb + yc - yd + ye
    return torch.stack([foo * bar, ya, yc**2, torch.tanh(yd), ye + 1])


# tests/test_root_find.py:196-213
def root_via_min(y, _):
    return torch.stack([torch.tanh(y[0] + 0.1), torch.tanh(y[1] - 0.5), torch.tanh(y[2] * 2)])


def bad_root_via_min(y, _):
    return torch.stack([torch.ones_like(y[0]), torch.tanh(y[1] - 0.5), torch.tanh(y[2] * 2)])
