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


# ---- run-time ROOT family: F_i(y) = y_i^3 + sum_j S_ij y_j - b_i, S shared (staged in shared memory), b per problem
CUBIC_SRC = r"""
  __device__ static __forceinline__ void eval_row(const T* sh, const T* ys, const Ctx& c, int lane, T& f, T (&row)[N], bool want_jac) {
    const T yi = ys[lane];
    T acc = yi * yi * yi - c.prob[lane];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const T s = sh[lane * N + j];
      acc += s * ys[j];
      if (want_jac) row[j] = s + (j == lane ? T(3) * yi * yi : T(0));
    }
    f = acc;
  }
"""
_cubic = None


def _cubic_family():
    global _cubic
    if _cubic is None:
        _cubic = optx.models.register_root_family(
            "cubic8", 8, CUBIC_SRC, shared_elems=64,
            torch_fn=lambda y, a: y**3 + y @ a["shared"].reshape(8, 8).T - a["prob"])
    return _cubic


def _cubic_numpy(Smat):
    def model(Y, args):
        b = args
        F = Y**3 + Y @ Smat.T - b
        J = Smat[None, :, :] + 3.0 * Y[:, :, None] ** 2 * np.eye(Y.shape[1])[None]
        return F.astype(Y.dtype), J.astype(Y.dtype)

    return model


@pytest.mark.parametrize("chord", (False, True), ids=("newton", "chord"))
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-10), (np.float32, 1e-5)])
def test_registered_root_family_vs_oracle(dtype, tol, chord):
    fam = _cubic_family()
    rng = np.random.default_rng(7)
    B = 1000
    Smat = (np.eye(8) * 3.0 + 0.3 * rng.standard_normal((8, 8))).astype(dtype)
    b = rng.uniform(-2, 2, (B, 8)).astype(dtype)
    y0 = np.zeros((B, 8), dtype)
    solver = (optx.Chord if chord else optx.Newton)(tol, tol)
    sol = optx.root_find(fam, solver, torch.as_tensor(y0, device=DEV), (torch.as_tensor(Smat.reshape(-1), device=DEV), torch.as_tensor(b, device=DEV)),
                         throw=False, max_steps=512)
    ref = S.newton_chord(_cubic_numpy(Smat), y0, b, rtol=tol, atol=tol, chord=chord, max_steps=512)
    assert (to_np(sol.result) == ref["result"]).all()
    if dtype == np.float64:
        assert np.array_equal(to_np(sol.stats["num_steps"]), ref["num_steps"])  # identical iteration counts
        assert np.allclose(to_np(sol.value), ref["value"], rtol=1e-10, atol=1e-12)
    else:
        assert np.mean(to_np(sol.stats["num_steps"]) == ref["num_steps"]) >= 0.95
        assert np.allclose(to_np(sol.value), ref["value"], rtol=1e-3, atol=1e-4)
    ok = ref["result"] == 0  # Chord (Jacobian frozen at y0 = 0) does not converge for every right-hand side
    assert ok.mean() > (0.5 if chord else 0.99)
    F = to_np(sol.value) ** 3 + to_np(sol.value) @ Smat.T - b
    assert np.abs(F[ok]).max() < (1e-8 if dtype == np.float64 else 1e-3)
    # host entry point gives the same bits; the generic step-kernel path the same roots
    host = optx.root_find(fam, solver, y0, (Smat.reshape(-1), b), throw=False, max_steps=512)
    assert np.array_equal(host.value, to_np(sol.value))
    if dtype == np.float64 and not chord:
        St = torch.as_tensor(Smat, device=DEV)
        gen = optx.root_find(lambda y, a: y**3 + a[0] @ y - a[1], solver, torch.as_tensor(y0, device=DEV), (St, torch.as_tensor(b, device=DEV)),
                             args_in_dims=(None, 0), throw=False, max_steps=512)
        assert torch.equal(gen.stats["num_steps"], sol.stats["num_steps"])
        assert torch.allclose(gen.value, sol.value, rtol=1e-10, atol=1e-12)
