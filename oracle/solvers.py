"""NumPy restatement of the reference solvers on the hot path, batched over problems.

TEST INFRASTRUCTURE (see oracle/__init__.py) - never imported by ``optimistix_b200``.

Every function cites the reference lines it restates (paths relative to /root/reference).
Batching follows ``jax.vmap`` of ``lax.while_loop``: the body is evaluated for every lane,
the carry of finished lanes is frozen, so each lane's value / num_steps / result equal the
unbatched solve.

Linear solves (lineax, not vendored, unpinned ">=0.0.6" - behaviour restated from its
published algorithm):
  * ``lx.QR``  -> reduced Householder QR + back substitution  (``linsolve="qr"``, reference-faithful)
  * Gram + Cholesky of ``J^T J + lambda I``                    (``linsolve="gram"``, what the CUDA kernel does)
  * ``lx.AutoLinearSolver(well_posed=None)`` on a square operator -> LU with partial pivoting
  * ``linear_solve(throw=False)``: a non-finite solution is reported as ``singular``.
"""
from __future__ import annotations

import numpy as np

# optimistix/_solution.py:11-20 (RESULTS extends lx.RESULTS); compare by NAME across implementations.
RESULT_NAMES = (
    "successful",
    "max_steps_reached",  # lineax
    "singular",  # lineax
    "breakdown",  # lineax
    "stagnation",  # lineax
    "nonlinear_max_steps_reached",
    "nonlinear_divergence",
    "nonfinite",
)
RES = {name: i for i, name in enumerate(RESULT_NAMES)}
OK, SINGULAR = RES["successful"], RES["singular"]
MAXSTEPS, DIVERGED, NONFINITE = RES["nonlinear_max_steps_reached"], RES["nonlinear_divergence"], RES["nonfinite"]


# ------------------------------------------------------------------ norms (_misc.py:15-45, lineax.internal)
def max_norm(x):
    return np.max(np.abs(x), axis=-1) if x.shape[-1] else np.zeros(x.shape[:-1], x.dtype)


def two_norm(x):
    return np.sqrt(np.sum(x * x, axis=-1))


def rms_norm(x):
    return np.sqrt(np.mean(x * x, axis=-1))


NORMS = {"max": max_norm, "two": two_norm, "rms": rms_norm}


def cauchy_termination(rtol, atol, norm, y, y_diff, f, f_diff):
    """_misc.py:291-310."""
    y_scale = atol + rtol * np.abs(y)
    f_scale = atol + rtol * np.abs(f)
    y_conv = norm(np.abs(y_diff) / y_scale) < 1
    f_conv = norm(np.abs(f_diff) / f_scale) < 1
    return y_conv & f_conv


def _allfinite(y):
    return np.all(np.isfinite(y), axis=-1)  # _misc.py:85-87


# ------------------------------------------------------------------ linear algebra helpers
def _back_substitute(R, b):
    n = R.shape[-1]
    x = np.zeros_like(b)
    for i in range(n - 1, -1, -1):
        s = b[:, i] - np.einsum("bj,bj->b", R[:, i, i + 1 :], x[:, i + 1 :])
        x[:, i] = s / R[:, i, i]
    return x


def _lstsq_qr(A, b):
    """lx.QR(): reduced QR then triangular solve; non-finite => singular (lineax, unverified)."""
    A = np.where(np.isfinite(A), A, np.nan)  # LAPACK geqrf may trap on inf; NaN propagates the same way
    Q, R = np.linalg.qr(A)
    x = _back_substitute(R, np.einsum("bmn,bm->bn", Q, b))
    res = np.where(_allfinite(x), OK, SINGULAR)
    return x, res


def _solve_gram_cholesky(J, r, lam):
    """What the CUDA LM kernel does: (J^T J + lam I) p = J^T r by Cholesky (BASELINE.json north_star)."""
    G = np.einsum("bmi,bmj->bij", J, J)
    g = np.einsum("bmi,bm->bi", J, r)
    return _solve_spd(G, g, lam)


def _solve_spd(G, g, lam):
    B, n, _ = G.shape
    A = G + lam[:, None, None] * np.eye(n, dtype=G.dtype)[None]
    L = np.zeros_like(A)
    for j in range(n):
        dj = A[:, j, j] - np.einsum("bk,bk->b", L[:, j, :j], L[:, j, :j])
        ljj = np.sqrt(dj)
        L[:, j, j] = ljj
        for i in range(j + 1, n):
            L[:, i, j] = (A[:, i, j] - np.einsum("bk,bk->b", L[:, i, :j], L[:, j, :j])) / ljj
    z = np.zeros_like(g)
    for i in range(n):
        z[:, i] = (g[:, i] - np.einsum("bk,bk->b", L[:, i, :i], z[:, :i])) / L[:, i, i]
    x = _back_substitute(np.swapaxes(L, 1, 2), z)
    res = np.where(_allfinite(x), OK, SINGULAR)
    return x, res


def _solve_lu(A, b):
    """LU with partial pivoting (getrf/getrs semantics: first maximal |pivot|); non-finite => singular."""
    A = A.copy()
    x = b.copy()
    B, n, _ = A.shape
    idx = np.arange(B)
    for k in range(n):
        p = k + np.argmax(np.abs(A[:, k:, k]), axis=1)
        rk, rp = A[idx, k].copy(), A[idx, p].copy()
        A[idx, k], A[idx, p] = rp, rk
        xk, xp = x[idx, k].copy(), x[idx, p].copy()
        x[idx, k], x[idx, p] = xp, xk
        l = A[:, k + 1 :, k] / A[:, k, k][:, None]
        A[:, k + 1 :, k + 1 :] -= l[:, :, None] * A[:, k, None, k + 1 :]
        x[:, k + 1 :] -= l * x[:, k, None]
        A[:, k + 1 :, k] = 0
    x = _back_substitute(A, x)
    res = np.where(_allfinite(x), OK, SINGULAR)
    return x, res


def _freeze(active, new, old):
    """vmap-of-while_loop carry masking: finished lanes keep their old carry."""
    a = active.reshape(active.shape + (1,) * (new.ndim - 1))
    return np.where(a, new, old)


# ------------------------------------------------------------------ Gauss-Newton family (LM, GN)
def gauss_newton_family(
    model, y0, args, *, rtol, atol, norm="max", max_steps=256, search="trust_region", linsolve="qr", trace=None,
    root_near_zero=False, pred="residual",
):
    """LevenbergMarquardt (search="trust_region") / GaussNewton (search="learning_rate").

    Restates: driver _iterate.py:194-268; AbstractGaussNewton.init/step gauss_newton.py:216-334;
    ClassicalTrustRegion trust_region.py:73-106,159-186 (constants :125-128); LearningRate
    learning_rate.py:24-35; DampedNewtonDescent + damped_newton_step levenberg_marquardt.py:23-76,
    108-133; NewtonDescent gauss_newton.py:64-78,107-127; cauchy_termination _misc.py:291-310.

    ``root_near_zero``: restate ``_LstsqToRoot`` (root_find / fixed_point through a least-squares
    solver, _root_find.py:45-80): additionally require ``norm(residual at the pre-step y) < atol``.

    ``linsolve`` / ``pred`` select the arithmetic, not the algorithm:
      * ``linsolve="qr", pred="residual"`` - the reference's own operations (lx.QR on the augmented system,
        predicted reduction ``0.5(|J d + r|^2 - |r|^2)`` from the two squared norms).  The parity oracle.
      * ``linsolve="gram", pred="gram"``   - the CUDA kernels' operations ("kernel arithmetic"): Cholesky of
        ``J^T J + lambda I`` and ``|J d + r|^2`` formed through the Gram identity ``r.r + (2 d.g + d.G d)`` and
        then rounded / subtracted like the reference's (csrc/common.cuh ``predicted_reduction``).
      * ``pred="gram_exact"``              - ``d.g + 0.5 d.G d`` without the rounding to the spacing of ``|r|^2``:
        what the round-1 kernels did; kept to document the +2-3 % iteration-count bias it causes
        (tests/test_oracle_pred_modes.py).

    Returns dict(value, num_steps, result, step_size, num_accepted).
    """
    if pred not in ("residual", "gram", "gram_exact"):
        raise ValueError(pred)
    nrm = NORMS[norm]
    y0 = np.array(y0)
    dt = y0.dtype
    B, n = y0.shape
    fi = np.finfo(dt)
    r0, J0 = model(y0, args)
    m = r0.shape[1]
    # gauss_newton.py:229-241 / trust_region.py:75 : zero-filled f_info, step_size 1, first_step True
    y = y0.copy()
    y_eval = y0.copy()
    r = np.zeros((B, m), dt)
    J = np.zeros((B, m, n), dt)
    newton = np.zeros((B, n), dt)  # NewtonDescent state (GN only)
    newton_res = np.full(B, OK)
    delta = np.ones(B, dt)
    first = np.ones(B, bool)
    terminate = np.zeros(B, bool)
    # _iterate.py:214-219: terminate once before the loop; non-finite y0 => nonfinite, loop never runs
    result = np.where(_allfinite(y0), OK, NONFINITE)
    num_steps = np.zeros(B, np.int32)
    num_accepted = np.zeros(B, np.int32)
    lm = search == "trust_region"
    # Not part of the reference: the smallest distance of any accept/good/terminate comparison from its
    # threshold, in units of the rounding noise of the quantities compared.  A problem whose margin is
    # large takes the same branches under ANY faithful floating-point evaluation order; one whose margin
    # is O(1) has a coin-flip iteration count (see DESIGN.md "iteration-count parity").
    margin = np.full(B, np.inf)

    with np.errstate(all="ignore"):
        for _ in range(max_steps):
            active = (~terminate) & (result == OK)  # _iterate.py:230-232
            if not active.any():
                break
            y_old = y
            r_e, J_e = model(y_eval, args)  # gauss_newton.py:253 (every step, accepted or not)
            d = y_eval - y
            if lm:
                # trust_region.py:86-104,182-186
                rtr = np.sum(r * r, axis=1)
                jd = np.einsum("bmn,bn->bm", J, d) + r
                if pred == "residual":
                    pred_v = dt.type(0.5) * (np.sum(jd * jd, axis=1) - rtr)
                else:
                    Gm = np.einsum("bmi,bmj->bij", J, J)
                    gv = np.einsum("bmi,bm->bi", J, r)
                    lin = np.einsum("bi,bi->b", d, gv)
                    quad = np.einsum("bi,bij,bj->b", d, Gm, d)
                    if pred == "gram":
                        pred_v = dt.type(0.5) * ((rtr + (dt.type(2) * lin + quad)) - rtr)
                    else:
                        pred_v = lin + dt.type(0.5) * quad
                f_min = dt.type(0.5) * rtr
                f_min_eval = dt.type(0.5) * np.sum(r_e * r_e, axis=1)
                df = f_min_eval - f_min
                accept = df <= dt.type(0.01) * pred_v
                good = (df < dt.type(0.99) * pred_v) & (pred_v < 0)
                good = good & ~first
                accept = accept | first
                mul = np.where(good, dt.type(3.5), dt.type(1.0))
                mul = np.where(accept, mul, dt.type(0.25))
                new_delta = mul * delta
                noise = fi.eps * np.maximum(f_min, f_min_eval) + fi.tiny
                mg = np.minimum(np.abs(df - dt.type(0.01) * pred_v), np.abs(df - dt.type(0.99) * pred_v)) / noise
                mg = np.where(first, np.inf, mg)
                margin = np.where(active, np.minimum(margin, mg), margin)
            else:
                accept = np.ones(B, bool)  # learning_rate.py:24-35
                new_delta = np.ones(B, dt)
            # gauss_newton.py:276-296 accepted/rejected branches
            term_acc = cauchy_termination(rtol, atol, nrm, y_eval, d, r_e, r_e - r)
            new_terminate = np.where(accept, term_acc, False)
            yn_ = nrm(np.abs(d) / (atol + rtol * np.abs(y_eval)))
            fn_ = nrm(np.abs(r_e - r) / (atol + rtol * np.abs(r_e)))
            mt = np.minimum(np.abs(yn_ - 1), np.abs(fn_ - 1)) / (64 * fi.eps)
            margin = np.where(active & accept, np.minimum(margin, mt), margin)
            y_new = _freeze(accept, y_eval, y)
            r_new = _freeze(accept, r_e, r)
            J_new = _freeze(accept, J_e, J)
            if lm:
                # levenberg_marquardt.py:55-57,66-69,75
                pred_ok = new_delta > fi.eps
                safe = np.where(pred_ok, new_delta, dt.type(1))
                lam = np.where(pred_ok, dt.type(1) / safe, fi.max)
                if linsolve == "qr":
                    sq = np.sqrt(lam)
                    A = np.concatenate([J_new, sq[:, None, None] * np.eye(n, dtype=dt)[None]], axis=1)
                    b = np.concatenate([r_new, np.zeros((B, n), dt)], axis=1)
                    p, lin_res = _lstsq_qr(A, b)
                else:
                    p, lin_res = _solve_gram_cholesky(J_new, r_new, lam)
                y_eval_new = y_new - p  # levenberg_marquardt.py:132, gauss_newton.py:315
                new_result = lin_res  # gauss_newton.py:316-318 (search result is always successful)
            else:
                # NewtonDescent.query only in the accepted branch (gauss_newton.py:277), always here
                if linsolve == "qr":
                    p, lin_res = _lstsq_qr(J_e, r_e)  # gauss_newton.py:64-78
                else:
                    p, lin_res = _solve_gram_cholesky(J_e, r_e, np.zeros(B, dt))
                newton_new, newton_res_new = p, lin_res
                y_eval_new = y_new - new_delta[:, None] * newton_new  # gauss_newton.py:126-127
                new_result = newton_res_new
            if root_near_zero:
                # prev_aux (gauss_newton.py:320): residual at the pre-step accepted y (r_e on the first step)
                f_prev = _freeze(first, r_e, r)
                new_terminate = new_terminate & (nrm(f_prev) < atol)
            # _iterate.py:216,238: non-finite check on the carry's y from BEFORE this step
            new_result = np.where(_allfinite(y_old), new_result, NONFINITE)
            if trace is not None:
                trace.append(dict(active=active.copy(), accept=accept.copy(), delta=new_delta.copy(), y_eval=y_eval_new.copy()))
            # commit, masked by `active`
            y = _freeze(active, y_new, y)
            y_eval = _freeze(active, y_eval_new, y_eval)
            r = _freeze(active, r_new, r)
            J = _freeze(active, J_new, J)
            delta = np.where(active, new_delta, delta)
            first = np.where(active, False, first)
            terminate = np.where(active, new_terminate, terminate)
            result = np.where(active, new_result, result)
            num_accepted = num_accepted + (active & accept)
            num_steps = num_steps + active
    # _iterate.py:253-257
    result = np.where((result == OK) & ~terminate, MAXSTEPS, result)
    return dict(value=y, num_steps=num_steps, result=result, step_size=delta, num_accepted=num_accepted, margin=margin)


def levenberg_marquardt(model, y0, args, **kw):
    return gauss_newton_family(model, y0, args, search="trust_region", **kw)


def gauss_newton(model, y0, args, **kw):
    return gauss_newton_family(model, y0, args, search="learning_rate", **kw)


# ------------------------------------------------------------------ Newton / Chord
def newton_chord(
    model, y0, args, *, rtol, atol, norm="max", max_steps=256, chord=False, cauchy=True, kappa=1e-2,
    lower=None, upper=None,
):
    """Newton / Chord root finders.

    Restates newton_chord.py:25-37 (_small/_diverged/_converged), :69-101 (init), :103-152 (step),
    :154-196 (terminate) under the driver _iterate.py:194-268.  ``model(y,args) -> (F, J)``.
    """
    nrm = NORMS[norm]
    y = np.array(y0)
    dt = y.dtype
    B, n = y.shape
    fval = np.full((B, n), np.inf, dt)  # :88-91 (f=+inf when cauchy_termination)
    diff = np.full((B, n), np.inf, dt)
    diffsize = np.full(B, np.inf, dt)
    diffsize_prev = np.ones(B, dt)
    lin_result = np.full(B, OK)
    step = np.zeros(B, np.int32)
    num_steps = np.zeros(B, np.int32)
    precision = np.finfo(dt).precision
    small_tol = 10.0 ** (2 - precision)

    if chord:
        _, J_fixed = model(y, args)  # :80-87 Jacobian linearised once at y0

    def terminate_fn(y_, fval_, diff_, diffsize_, diffsize_prev_, lin_result_, step_):
        if cauchy:
            term = cauchy_termination(rtol, atol, nrm, y_, diff_, np.zeros_like(fval_), fval_)  # :160-176
            term_res = np.full(B, OK)
        else:
            at_least_two = step_ >= 2  # :179-192
            rate = diffsize_ / diffsize_prev_
            factor = diffsize_ * rate / (1 - rate)
            small = diffsize_ < small_tol
            diverged = ~np.isfinite(rate) | (rate > 2)
            converged = (factor > 0) & (factor < kappa)
            term = at_least_two & (small | diverged | converged)
            term_res = np.where(at_least_two & ~small & diverged, DIVERGED, OK)
        fail = lin_result_ != OK  # :193-195
        res = np.where(fail, lin_result_, term_res)
        return fail | term, res

    with np.errstate(all="ignore"):
        terminate, result = terminate_fn(y, fval, diff, diffsize, diffsize_prev, lin_result, step)
        result = np.where(_allfinite(y), result, NONFINITE)
        for _ in range(max_steps):
            active = (~terminate) & (result == OK)
            if not active.any():
                break
            y_old = y
            fx, Jx = model(y, args)  # :116-119
            sol, lres = _solve_lu(J_fixed if chord else Jx, fx)  # :120 / :125-127
            new_y = y - sol  # :129
            d = sol
            if lower is not None:
                new_y = np.maximum(new_y, lower)  # :130-131
            if upper is not None:
                new_y = np.minimum(new_y, upper)  # :132-133
            if lower is not None or upper is not None:
                d = y - new_y  # :134-135
            scale = atol + rtol * np.abs(new_y)  # :136
            dsz = nrm(d / scale)  # :137-138
            n_fval, n_diff, n_diffsize, n_prev = fx, d, dsz, diffsize
            n_step = step + 1
            n_term, n_res = terminate_fn(y_old, n_fval, n_diff, n_diffsize, n_prev, lres, n_step)  # _iterate.py:238
            n_res = np.where(_allfinite(y_old), n_res, NONFINITE)
            y = _freeze(active, new_y, y)
            fval = _freeze(active, n_fval, fval)
            diff = _freeze(active, n_diff, diff)
            diffsize = np.where(active, n_diffsize, diffsize)
            diffsize_prev = np.where(active, n_prev, diffsize_prev)
            lin_result = np.where(active, lres, lin_result)
            step = np.where(active, n_step, step)
            terminate = np.where(active, n_term, terminate)
            result = np.where(active, n_res, result)
            num_steps = num_steps + active
    result = np.where((result == OK) & ~terminate, MAXSTEPS, result)
    return dict(value=y, num_steps=num_steps, result=result)


def _solve_cholesky(H, g):
    """lx.linear_solve(H, g, lx.Cholesky(), throw=False): lower factor, two triangular solves; a non-positive pivot
    gives NaN (jnp.linalg.cholesky semantics) and a non-finite solution is reported as `singular`."""
    B, n, _ = H.shape
    x = np.full((B, n), np.nan, H.dtype)
    for b in range(B):
        try:
            L = np.linalg.cholesky(H[b])
        except np.linalg.LinAlgError:
            continue
        z = np.zeros(n, H.dtype)
        for i in range(n):
            z[i] = (g[b, i] - L[i, :i] @ z[:i]) / L[i, i]
        xb = np.zeros(n, H.dtype)
        for i in range(n - 1, -1, -1):
            xb[i] = (z[i] - L[i + 1 :, i] @ xb[i + 1 :]) / L[i, i]
        x[b] = xb
    res = np.where(np.isfinite(x).all(axis=1), OK, SINGULAR)
    return x, res


# ------------------------------------------------------------------ BFGS
def bfgs(model, y0, args=None, *, rtol, atol, norm="max", max_steps=256, near_zero_residual=None, update="bfgs", use_inverse=True):
    """BFGS / DFP (``update``) with BacktrackingArmijo, on the inverse Hessian (``use_inverse=True``, the default) or on
    the Hessian approximation itself with a Cholesky solve for the direction (quasi_newton.py:387-399, :487-573;
    NewtonDescent(linear_solver=lx.Cholesky()) :455, gauss_newton.py:64-78: non-finite solution => singular).

    Restates quasi_newton.py:176-202 (init), :204-300 (step), :334-340 (init_hessian), :342-415
    (update_hessian, inverse branch :371-385); backtracking.py:29-31,51-53,79-99; NewtonDescent
    gauss_newton.py:60-62,126-127; _iterate.py:194-268.  ``model(y,args) -> (f, grad)``.

    ``near_zero_residual``: optional callable y -> residual vector, restating the extra
    ``_ToRoot.terminate`` condition ``norm(f) < atol`` of root_find-via-minimiser (_root_find.py:45-80).
    """
    nrm = NORMS[norm]
    y = np.array(y0)
    dt = y.dtype
    B, n = y.shape
    eps = np.finfo(dt).eps
    y_eval = y.copy()
    f = np.zeros(B, dt)
    g = np.zeros((B, n), dt)
    Hinv = np.broadcast_to(np.eye(n, dtype=dt), (B, n, n)).copy()  # :41-75
    newton = np.zeros((B, n), dt)
    alpha = np.ones(B, dt)  # backtracking.py:51-53
    first = np.ones(B, bool)
    terminate = np.zeros(B, bool)
    result = np.where(_allfinite(y), OK, NONFINITE)
    num_steps = np.zeros(B, np.int32)
    num_accepted = np.zeros(B, np.int32)
    froot = None
    if near_zero_residual is not None:
        froot = np.full(B, np.inf, dt)
    dres = np.full(B, OK)

    with np.errstate(all="ignore"):
        for _ in range(max_steps):
            active = (~terminate) & (result == OK)
            if not active.any():
                break
            y_old = y
            f_e, g_e = model(y_eval, args)  # :214-216 (+ lin_to_grad :227, only used when accepted)
            d = y_eval - y
            pred = np.einsum("bn,bn->b", g, d)  # backtracking.py:79-80
            df = f_e - f
            armijo = df <= dt.type(0.1) * pred  # :87
            has_red = pred <= 0  # :88
            accept = first | (armijo & has_red)  # :90
            n_alpha = np.where(accept, dt.type(1.0), dt.type(0.5) * alpha)  # :91-93
            # accepted branch: quasi_newton.py:226-257 / update_hessian :342-415
            q = g_e - g
            inner = np.einsum("bn,bn->b", q, d)
            upd = inner > eps  # :360
            # the four updates are two formulas with the roles of s = d and q exchanged (:371-399, :516-547)
            u, v = (d, q) if use_inverse else (q, d)
            Hq = np.einsum("bij,bj->bi", Hinv, v)
            mvp_inner = np.einsum("bn,bn->b", v, Hq)
            if (update == "bfgs") == bool(use_inverse):
                term1 = ((inner + mvp_inner)[:, None, None] * (u[:, :, None] * u[:, None, :])) / (inner**2)[:, None, None]
                term2 = (Hq[:, :, None] * u[:, None, :] + u[:, :, None] * Hq[:, None, :]) / inner[:, None, None]
            else:
                term1 = (u[:, :, None] * u[:, None, :]) / inner[:, None, None]
                term2 = (Hq[:, :, None] * Hq[:, None, :]) / mvp_inner[:, None, None]
            H_upd = Hinv + term1 - term2  # :375-382
            H_acc = _freeze(upd, H_upd, Hinv)
            if use_inverse:
                newton_acc = np.einsum("bij,bj->bi", H_acc, g_e)  # gauss_newton.py:60-61
                dres_acc = np.full(B, OK)
            else:
                newton_acc, dres_acc = _solve_cholesky(H_acc, g_e)  # gauss_newton.py:64-78
            term_acc = cauchy_termination(rtol, atol, nrm, y_eval, d, f_e[:, None], df[:, None])  # :244-246
            term_acc = term_acc & ~first  # :247-249
            n_terminate = np.where(accept, term_acc, False)
            y_new = _freeze(accept, y_eval, y)
            f_new = np.where(accept, f_e, f)
            g_new = _freeze(accept, g_e, g)
            H_new = _freeze(accept, H_acc, Hinv)
            newton_new = _freeze(accept, newton_acc, newton)
            dres = np.where(accept, dres_acc, dres)
            y_eval_new = y_new - n_alpha[:, None] * newton_new  # gauss_newton.py:126-127, quasi_newton.py:281-282
            n_result = dres.copy()  # quasi_newton.py:280-283
            if froot is not None:
                # _ToRoot.step stores the solver's returned aux, which is prev_aux = the residual at the
                # carry's pre-step y (quasi_newton.py:285: where(first_step, aux, state.aux); on the first
                # step y_old == y_eval == y0).  terminate = solver terminate & norm(f) < atol
                # (_root_find.py:64-73)
                fr = nrm(near_zero_residual(y_old))
                froot = np.where(active, fr, froot)
                n_terminate = n_terminate & (fr < atol)
            n_result = np.where(_allfinite(y_old), n_result, NONFINITE)
            y = _freeze(active, y_new, y)
            y_eval = _freeze(active, y_eval_new, y_eval)
            f = np.where(active, f_new, f)
            g = _freeze(active, g_new, g)
            Hinv = _freeze(active, H_new, Hinv)
            newton = _freeze(active, newton_new, newton)
            alpha = np.where(active, n_alpha, alpha)
            first = np.where(active, False, first)
            terminate = np.where(active, n_terminate, terminate)
            result = np.where(active, n_result, result)
            num_accepted = num_accepted + (active & accept)
            num_steps = num_steps + active
    result = np.where((result == OK) & ~terminate, MAXSTEPS, result)
    return dict(value=y, num_steps=num_steps, result=result, hessian_inv=Hinv, f=f, grad=g, num_accepted=num_accepted)
