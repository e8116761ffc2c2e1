"""Oracle-side (NumPy, batched) restatement of the registered residual-model families.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each family mirrors one device function family
in ``optimistix_b200/csrc/models.cuh``.  The reference has no notion of a "model family": the
nearest concept is the user callable ``fn(y, args)`` whose Jacobian is obtained by
``jax.linearize`` (gauss_newton.py:158-181) - here the Jacobian is written analytically, which
is what ``jax.linearize`` evaluates to.

Conventions: ``y`` is [B, n]; residual models return (r [B, m], J [B, m, n]); root models
return (F [B, n], J [B, n, n]); minimisation models return (f [B], g [B, n]).
Arithmetic is done in ``y.dtype``.
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------- least squares families
def exp_decay(y, args):
    t, d = args["t"], args["d"]
    a, b, c = y[:, 0:1], y[:, 1:2], y[:, 2:3]
    e = np.exp(-b * t[None, :])
    r = a * e + c - d
    J = np.empty(r.shape + (3,), y.dtype)
    J[..., 0] = e
    J[..., 1] = -a * t[None, :] * e
    J[..., 2] = 1.0
    return r, J


def gauss2(y, args):
    x, d = args["x"], args["d"]
    one = y.dtype.type(1.0)
    half = y.dtype.type(0.5)
    A1, mu1, s1 = y[:, 0:1], y[:, 1:2], y[:, 2:3]
    A2, mu2, s2 = y[:, 3:4], y[:, 4:5], y[:, 5:6]
    b0, b1 = y[:, 6:7], y[:, 7:8]
    i1, i2 = one / s1, one / s2
    z1 = (x[None, :] - mu1) * i1
    z2 = (x[None, :] - mu2) * i2
    e1 = np.exp(-half * z1 * z1)
    e2 = np.exp(-half * z2 * z2)
    r = A1 * e1 + A2 * e2 + b0 + b1 * x[None, :] - d
    J = np.empty(r.shape + (8,), y.dtype)
    J[..., 0] = e1
    J[..., 1] = A1 * e1 * z1 * i1
    J[..., 2] = A1 * e1 * z1 * z1 * i1
    J[..., 3] = e2
    J[..., 4] = A2 * e2 * z2 * i2
    J[..., 5] = A2 * e2 * z2 * z2 * i2
    J[..., 6] = 1.0
    J[..., 7] = x[None, :]
    return r, J


def linear_rows(y, args):
    """r = J y - d with an explicit, y-independent Jacobian (configs[4])."""
    J, d = args["J"], args["d"]
    r = np.einsum("mn,bn->bm", J, y) - d[None, :]
    return r, np.broadcast_to(J, (y.shape[0],) + J.shape)


# ---------------------------------------------------------------- root-finding families
def implicit_euler(y, args):
    """F(y) = y - y_prev - dt*(A y + y*(B y));  J = I - dt*(A + diag(B y) + diag(y) B)."""
    A, Bm, y_prev, dt = args["A"], args["B"], args["y_prev"], args["dt"]
    Ay = y @ A.T
    By = y @ Bm.T
    F = y - y_prev - dt[:, None] * (Ay + y * By)
    n = y.shape[1]
    eye = np.eye(n, dtype=y.dtype)
    Jf = A[None] + By[:, :, None] * eye[None] + y[:, :, None] * Bm[None]
    J = eye[None] - dt[:, None, None] * Jf
    return F, J


# ---------------------------------------------------------------- minimisation families
def rosenbrock(y, args=None):
    """Standard chained Rosenbrock: sum_i 100 (y_{i+1}-y_i^2)^2 + (1-y_i)^2."""
    a = y[:, 1:] - y[:, :-1] ** 2
    b = 1.0 - y[:, :-1]
    f = np.sum(100.0 * a * a + b * b, axis=1)
    g = np.zeros_like(y)
    g[:, :-1] += -400.0 * a * y[:, :-1] - 2.0 * b
    g[:, 1:] += 200.0 * a
    return f, g


FAMILIES = {
    "exp_decay": ("lsq", exp_decay),
    "gauss2": ("lsq", gauss2),
    "linear_rows": ("lsq", linear_rows),
    "implicit_euler": ("root", implicit_euler),
    "rosenbrock": ("min", rosenbrock),
}
