"""Seeded synthetic workloads for the five BASELINE.json configs.

These are *inputs only* (NumPy, host side). They are shared by the product path
(bench.py, smoke) and by the test oracle, so they live in the package and not
under ``oracle/``.  Nothing here solves anything.

Parameter layouts
-----------------
exp_decay   (n=3):  y = [a, b, c],                 r_i = a*exp(-b*t_i) + c - d_i
gauss2      (n=8):  y = [A1, mu1, s1, A2, mu2, s2, b0, b1],
                    r_i = A1*exp(-(x_i-mu1)^2/(2 s1^2)) + A2*exp(-(x_i-mu2)^2/(2 s2^2))
                          + b0 + b1*x_i - d_i
implicit_euler (n=32): F(y) = y - y_prev - dt*(A y + y*(B y))
rosenbrock  (n=64): f(y) = sum_i 100*(y_{i+1}-y_i^2)^2 + (1-y_i)^2   (standard chained form;
                    NOT tests/helpers.py:329-334 of the reference, which is linear - SURVEY F7)
"""
from __future__ import annotations

import numpy as np


def exp_decay_batch(batch: int = 4096, m: int = 64, seed: int = 0, dtype=np.float64):
    """configs[0]: LM exponential-decay curve fits (n=3, m=64, fp64)."""
    rng = np.random.default_rng(seed)
    t = np.linspace(0.0, 4.0, m)
    a = rng.uniform(1.0, 3.0, batch)
    b = rng.uniform(0.5, 2.0, batch)
    c = rng.uniform(-0.5, 0.5, batch)
    truth = np.stack([a, b, c], axis=1)
    model = a[:, None] * np.exp(-b[:, None] * t[None, :]) + c[:, None]
    d = model + 0.01 * rng.standard_normal((batch, m))
    y0 = truth * (1.0 + 0.2 * rng.uniform(-1.0, 1.0, (batch, 3)))
    return dict(t=t.astype(dtype), d=d.astype(dtype), y0=y0.astype(dtype), truth=truth.astype(dtype))


def gauss2_batch(batch: int = 1 << 20, m: int = 256, seed: int = 1, dtype=np.float32, chunk: int = 1 << 16):
    """configs[1]: LM two-Gaussian + linear-baseline fits (n=8, m=256, fp32)."""
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, m)
    truth = np.empty((batch, 8), np.float64)
    truth[:, 0] = rng.uniform(0.5, 2.0, batch)
    truth[:, 1] = rng.uniform(0.2, 0.4, batch)
    truth[:, 2] = rng.uniform(0.03, 0.08, batch)
    truth[:, 3] = rng.uniform(0.5, 2.0, batch)
    truth[:, 4] = rng.uniform(0.6, 0.8, batch)
    truth[:, 5] = rng.uniform(0.03, 0.08, batch)
    truth[:, 6] = rng.uniform(-0.2, 0.2, batch)
    truth[:, 7] = rng.uniform(-0.2, 0.2, batch)
    y0 = truth * (1.0 + 0.1 * rng.uniform(-1.0, 1.0, (batch, 8)))
    d = np.empty((batch, m), dtype)
    for s in range(0, batch, chunk):  # chunked so that 1M x 256 never needs fp64 temporaries of full size
        p = truth[s : s + chunk]
        z1 = (x[None, :] - p[:, 1:2]) / p[:, 2:3]
        z2 = (x[None, :] - p[:, 4:5]) / p[:, 5:6]
        f = p[:, 0:1] * np.exp(-0.5 * z1 * z1) + p[:, 3:4] * np.exp(-0.5 * z2 * z2) + p[:, 6:7] + p[:, 7:8] * x[None, :]
        f += 0.01 * rng.standard_normal(f.shape)
        d[s : s + chunk] = f.astype(dtype)
    return dict(x=x.astype(dtype), d=d, y0=y0.astype(dtype), truth=truth.astype(dtype))


def implicit_euler_batch(batch: int = 1 << 18, n: int = 32, seed: int = 2, dtype=np.float64):
    """configs[2]: Newton root finds, implicit-Euler step of a stiff quadratic ODE (n=32, fp64)."""
    rng = np.random.default_rng(seed)
    A = -np.diag(np.logspace(0.0, 6.0, n)) + 0.1 * rng.standard_normal((n, n))
    B = 0.01 * rng.standard_normal((n, n))
    y_prev = rng.uniform(0.0, 1.0, (batch, n))
    dt = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), batch))
    return dict(A=A.astype(dtype), B=B.astype(dtype), y_prev=y_prev.astype(dtype), dt=dt.astype(dtype),
                y0=y_prev.astype(dtype).copy())


def rosenbrock_batch(batch: int = 1 << 16, n: int = 64, seed: int = 3, dtype=np.float64):
    """configs[3]: BFGS minimisations of the standard chained Rosenbrock function (n=64, fp64)."""
    rng = np.random.default_rng(seed)
    y0 = 1.0 + 0.3 * rng.standard_normal((batch, n))
    return dict(y0=y0.astype(dtype))


def linear_lsq_rows(m: int, n: int = 256, seed: int = 4, dtype=np.float32, rank: int = 0):
    """configs[4]: one row shard of a well-conditioned linear-in-parameters least-squares problem.

    r(y) = J y - d with J ~ N(0,1)/sqrt(n)-ish columns of graded scale (kappa(J) ~ 1e2 by construction).
    Small sizes only on the host; bench.py generates the full-size shard on the device with the same recipe.
    """
    rng = np.random.default_rng(seed * 1000 + rank)
    scale = np.logspace(0.0, -2.0, n)
    J = rng.standard_normal((m, n)) * scale[None, :]
    y_true = np.random.default_rng(seed).standard_normal(n) / scale
    d = J @ y_true + 0.01 * rng.standard_normal(m)
    return dict(J=J.astype(dtype), d=d.astype(dtype), y_true=y_true.astype(np.float64))
