"""Run-time registered model families (b200opt_register_curve_family: NVRTC-compiled device functions in the persistent
LM / GN kernel) against the generic step-kernel path on the same function and against the oracle.  GPU."""
import numpy as np
import pytest
import torch

import optimistix_b200 as optx
from optimistix_b200 import _lib
from oracle import solvers as S

from .parity import compare, to_np

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# Lorentzian peak + offset: r_i = A / (1 + ((x_i - x0) / w)^2) + c - d_i,  y = [A, x0, w, c]
LORENTZ_SRC = r"""
  struct Pre { T iw; };
  __device__ static __forceinline__ Pre prepare(const T (&y)[N]) { return Pre{T(1) / y[2]}; }
  __device__ static __forceinline__ T residual(const T (&y)[N], const Pre& p, const CurveRow<T>& row, T (&aux)[NAUX]) {
    const T z = (row.x - y[1]) * p.iw;
    const T q = T(1) / (T(1) + z * z);
    aux[0] = q;
    return y[0] * q + y[3] - row.d;
  }
  __device__ static __forceinline__ void jac_row(const T (&y)[N], const Pre& p, const CurveRow<T>& row, const T (&aux)[NAUX], T (&j)[N]) {
    const T z = (row.x - y[1]) * p.iw;
    const T q = aux[0];
    const T dq = T(2) * y[0] * q * q * z * p.iw;
    j[0] = q;
    j[1] = dq;
    j[2] = dq * z;
    j[3] = T(1);
  }
"""


def _lorentz_torch(y, a):
    x, d = a["x"], a["d"]
    z = (x - y[..., 1:2]) / y[..., 2:3]
    return y[..., 0:1] / (1 + z * z) + y[..., 3:4] - d


def _lorentz_numpy(Y, args):
    x, d = args["x"], args["d"]
    A, x0, w, c = (Y[:, k : k + 1] for k in range(4))
    z = (x[None, :] - x0) / w
    q = 1.0 / (1.0 + z * z)
    r = A * q + c - d
    dq = 2.0 * A * q * q * z / w
    J = np.stack([q, dq, dq * z, np.ones_like(q)], axis=2)
    return r.astype(Y.dtype), J.astype(Y.dtype)


_model = None


def _family():
    global _model
    if _model is None:
        _model = optx.models.register_curve_family("lorentz", 4, LORENTZ_SRC, naux=1, torch_fn=_lorentz_torch)
    return _model


def _data(B, m, dtype, seed=0):
    rng = np.random.default_rng(seed)
    x = np.linspace(-1.0, 1.0, m)
    truth = np.stack([rng.uniform(1, 3, B), rng.uniform(-0.3, 0.3, B), rng.uniform(0.1, 0.3, B), rng.uniform(-0.5, 0.5, B)], 1)
    z = (x[None, :] - truth[:, 1:2]) / truth[:, 2:3]
    d = truth[:, 0:1] / (1 + z * z) + truth[:, 3:4] + 0.01 * rng.standard_normal((B, m))
    y0 = truth * (1 + 0.1 * rng.uniform(-1, 1, truth.shape))
    return x.astype(dtype), d.astype(dtype), y0.astype(dtype), truth


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-8), (np.float32, 1e-4)])
@pytest.mark.parametrize("m", [64, 100])
def test_registered_family_vs_oracle_and_generic(dtype, tol, m):
    fam = _family()
    assert fam.model_id >= 1000 and _lib.load().b200opt_model_id(b"lorentz") == fam.model_id
    B = 512
    x, d, y0, truth = _data(B, m, dtype)
    xs, ds, ys = (torch.as_tensor(a, device=DEV) for a in (x, d, y0))
    sol = optx.least_squares(fam, optx.LevenbergMarquardt(tol, tol), ys, (xs, ds), throw=False)
    ref = S.levenberg_marquardt(_lorentz_numpy, y0, dict(x=x, d=d), rtol=tol, atol=tol, linsolve="gram")
    s = compare(sol, ref, tol, tol, f"jit lorentz {np.dtype(dtype).name} m={m} vs oracle")
    assert s["result_match"] >= 0.995 and s["within_tol"] >= (1.0 if dtype == np.float64 else 0.99)
    # fp32: the last accept/reject decisions sit in the rounding noise of the loss, counts differ by a step
    assert abs(s["mean_steps"][0] - s["mean_steps"][1]) / s["mean_steps"][1] < (0.05 if dtype == np.float64 else 0.10)
    assert np.median(np.abs(to_np(sol.value) - truth), 0).max() < 0.02

    def fn(y, a):
        xx, dd = a
        z = (xx - y[1]) / y[2]
        return y[0] / (1 + z * z) + y[3] - dd

    gen = optx.least_squares(fn, optx.LevenbergMarquardt(tol, tol), ys, (xs, ds), args_in_dims=(None, 0), throw=False)
    scaled = ((sol.value - gen.value).abs() / (tol + tol * gen.value.abs())).amax(1)
    assert float((scaled <= 1.0).float().mean()) >= (1.0 if dtype == np.float64 else 0.99)
    # Gauss-Newton through the same compiled module, and the host entry point
    gn = optx.least_squares(fam, optx.GaussNewton(tol, tol), ys, (xs, ds), throw=False)
    gref = S.gauss_newton(_lorentz_numpy, y0, dict(x=x, d=d), rtol=tol, atol=tol, linsolve="gram")
    assert compare(gn, gref, tol, tol, "jit lorentz GN vs oracle")["within_tol"] >= 0.99
    host = optx.least_squares(fam, optx.LevenbergMarquardt(tol, tol), y0, (x, d), throw=False)
    assert np.array_equal(host.value, to_np(sol.value)) and np.array_equal(host.result, to_np(sol.result))


def test_registered_family_prefix_parity_fp64():
    fam = _family()
    x, d, y0, _ = _data(256, 64, np.float64, seed=3)
    xs, ds, ys = (torch.as_tensor(a, device=DEV) for a in (x, d, y0))
    for k in (1, 2, 3, 4):
        sol = optx.least_squares(fam, optx.LevenbergMarquardt(1e-8, 1e-8), ys, (xs, ds), throw=False, max_steps=k)
        ref = S.levenberg_marquardt(_lorentz_numpy, y0, dict(x=x, d=d), rtol=1e-8, atol=1e-8, linsolve="gram", max_steps=k)
        safe = ref["margin"] > 1e3
        assert np.array_equal(to_np(sol.stats["num_steps"])[safe], ref["num_steps"][safe])
        assert np.array_equal(to_np(sol.stats["num_accepted_steps"])[safe], ref["num_accepted"][safe])
        assert np.allclose(to_np(sol.value)[safe], ref["value"][safe], rtol=1e-9, atol=0)


def test_registered_family_gradient():
    fam = _family()
    x, d, y0, _ = _data(8, 64, np.float64, seed=5)
    xs, ys = torch.as_tensor(x, device=DEV), torch.as_tensor(y0, device=DEV)
    ds = torch.as_tensor(d, device=DEV).requires_grad_(True)
    sol = optx.least_squares(fam, optx.LevenbergMarquardt(1e-12, 1e-12), ys, (xs, ds), throw=False)
    sol.value[:, 1].sum().backward()  # peak positions w.r.t. the data, by ImplicitAdjoint through the registered torch_fn
    h = 1e-6
    with torch.no_grad():
        dp, dm = ds.detach().clone(), ds.detach().clone()
        dp[3, 20] += h
        dm[3, 20] -= h
        fd = (optx.least_squares(fam, optx.LevenbergMarquardt(1e-12, 1e-12), ys, (xs, dp), throw=False).value[3, 1]
              - optx.least_squares(fam, optx.LevenbergMarquardt(1e-12, 1e-12), ys, (xs, dm), throw=False).value[3, 1]) / (2 * h)
    assert abs(float(ds.grad[3, 20]) - float(fd)) <= 1e-4 * max(1.0, abs(float(fd)))


def test_registration_errors_are_loud():
    with pytest.raises(_lib.B200OptError, match="compilation failed"):
        optx.models.register_curve_family("broken", 2, "this is not C++;")
    with pytest.raises(ValueError):
        _family()
        optx.models.register_curve_family("lorentz", 4, LORENTZ_SRC)
    with pytest.raises(_lib.B200OptError):
        optx.models.register_curve_family("too_big", 17, LORENTZ_SRC)
