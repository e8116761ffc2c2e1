/* Included twice by lm_oracle.c with REAL = double / float and SUFFIX = _f64 / _f32.
 *
 * TEST INFRASTRUCTURE: a scalar C restatement of the reference's LevenbergMarquardt step for the
 * "curve" model families, one problem at a time, parallel over problems with OpenMP.  Used as the
 * CPU baseline (bench.py cpu_baseline / --impl reference) and cross-checked against oracle/solvers.py.
 * Cites are to /root/reference/optimistix.  Never linked into the product library.
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* model: r_i and Jacobian row (what jax.linearize evaluates to, gauss_newton.py:166-169) */
static void FN(model_eval)(int model, int n, int m, const REAL* y, const REAL* x, const REAL* d, REAL* r, REAL* J /* m x n row-major */) {
  if (model == 0) { /* exp_decay */
    for (int i = 0; i < m; ++i) {
      REAL e = EXP(-y[1] * x[i]);
      r[i] = y[0] * e + y[2] - d[i];
      J[i * n + 0] = e;
      J[i * n + 1] = -y[0] * x[i] * e;
      J[i * n + 2] = (REAL)1;
    }
  } else { /* gauss2 */
    REAL i1 = (REAL)1 / y[2], i2 = (REAL)1 / y[5];
    for (int i = 0; i < m; ++i) {
      REAL z1 = (x[i] - y[1]) * i1, z2 = (x[i] - y[4]) * i2;
      REAL e1 = EXP((REAL)-0.5 * z1 * z1), e2 = EXP((REAL)-0.5 * z2 * z2);
      r[i] = y[0] * e1 + y[3] * e2 + y[6] + y[7] * x[i] - d[i];
      REAL* j = J + i * n;
      j[0] = e1;
      j[1] = y[0] * e1 * z1 * i1;
      j[2] = y[0] * e1 * z1 * z1 * i1;
      j[3] = e2;
      j[4] = y[3] * e2 * z2 * i2;
      j[5] = y[3] * e2 * z2 * z2 * i2;
      j[6] = (REAL)1;
      j[7] = x[i];
    }
  }
}

/* lx.QR on the augmented system [J; sqrt(lam) I] p = [r; 0] (levenberg_marquardt.py:66-75):
 * Householder QR (LAPACK geqrf convention) + back substitution. A is (m+n) x n column-major scratch. */
static void FN(solve_qr)(int n, int m, const REAL* J, const REAL* r, REAL lam, REAL* A, REAL* b, REAL* p) {
  const int M = m + n;
  const REAL sq = SQRT(lam);
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < m; ++i) A[j * M + i] = J[i * n + j];
    for (int i = 0; i < n; ++i) A[j * M + m + i] = (i == j) ? sq : (REAL)0;
  }
  for (int i = 0; i < m; ++i) b[i] = r[i];
  for (int i = 0; i < n; ++i) b[m + i] = (REAL)0;
  for (int k = 0; k < n; ++k) {
    REAL* a = A + k * M;
    REAL nrm = 0;
    for (int i = k; i < M; ++i) nrm += a[i] * a[i];
    nrm = SQRT(nrm);
    if (nrm == (REAL)0) continue; /* R_kk = 0 -> division by zero below -> non-finite -> singular */
    REAL alpha = a[k] > 0 ? -nrm : nrm;
    REAL v0 = a[k] - alpha;
    /* v = (v0, a[k+1..]); H = I - 2 v v^T / (v^T v) */
    REAL vtv = v0 * v0;
    for (int i = k + 1; i < M; ++i) vtv += a[i] * a[i];
    if (vtv == (REAL)0) continue;
    REAL beta = (REAL)2 / vtv;
    for (int j = k + 1; j < n; ++j) {
      REAL* c = A + j * M;
      REAL s = v0 * c[k];
      for (int i = k + 1; i < M; ++i) s += a[i] * c[i];
      s *= beta;
      c[k] -= s * v0;
      for (int i = k + 1; i < M; ++i) c[i] -= s * a[i];
    }
    REAL s = v0 * b[k];
    for (int i = k + 1; i < M; ++i) s += a[i] * b[i];
    s *= beta;
    b[k] -= s * v0;
    for (int i = k + 1; i < M; ++i) b[i] -= s * a[i];
    a[k] = alpha;
  }
  for (int i = n - 1; i >= 0; --i) {
    REAL s = b[i];
    for (int j = i + 1; j < n; ++j) s -= A[j * M + i] * p[j];
    p[i] = s / A[i * M + i];
  }
}

/* Gram + Cholesky, the arithmetic of the CUDA kernel */
static void FN(solve_gram)(int n, int m, const REAL* J, const REAL* r, REAL lam, REAL* G, REAL* p) {
  REAL g[16], z[16];
  for (int a = 0; a < n; ++a) {
    for (int c = 0; c <= a; ++c) {
      REAL s = 0;
      for (int i = 0; i < m; ++i) s += J[i * n + a] * J[i * n + c];
      G[a * n + c] = s;
    }
    REAL s = 0;
    for (int i = 0; i < m; ++i) s += J[i * n + a] * r[i];
    g[a] = s;
    G[a * n + a] += lam;
  }
  for (int j = 0; j < n; ++j) {
    REAL s = G[j * n + j];
    for (int k = 0; k < j; ++k) s -= G[j * n + k] * G[j * n + k];
    REAL l = SQRT(s);
    G[j * n + j] = l;
    for (int i = j + 1; i < n; ++i) {
      REAL t = G[i * n + j];
      for (int k = 0; k < j; ++k) t -= G[i * n + k] * G[j * n + k];
      G[i * n + j] = t / l;
    }
  }
  for (int i = 0; i < n; ++i) {
    REAL s = g[i];
    for (int k = 0; k < i; ++k) s -= G[i * n + k] * z[k];
    z[i] = s / G[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    REAL s = z[i];
    for (int k = i + 1; k < n; ++k) s -= G[k * n + i] * p[k];
    p[i] = s / G[i * n + i];
  }
}

static int FN(allfinite)(const REAL* v, int n) {
  for (int i = 0; i < n; ++i)
    if (!isfinite(v[i])) return 0;
  return 1;
}

/* one problem: driver _iterate.py:194-268 + AbstractGaussNewton.step gauss_newton.py:243-334 +
 * ClassicalTrustRegion trust_region.py:77-106,182-186 + damped_newton_step levenberg_marquardt.py:34-76 +
 * cauchy_termination (max_norm) _misc.py:291-310 */
static void FN(lm_one)(int model, int n, int m, const REAL* y0, const REAL* x, const REAL* d, REAL rtol, REAL atol,
                       int max_steps, int linsolve, REAL* scratch, REAL* y_out, int32_t* steps_out, int32_t* result_out) {
  REAL y[16], ye[16], dd[16], p[16];
  REAL* r = scratch;            /* m */
  REAL* re = r + m;             /* m */
  REAL* J = re + m;             /* m*n accepted */
  REAL* Je = J + (size_t)m * n; /* m*n trial */
  REAL* A = Je + (size_t)m * n; /* (m+n)*n or n*n */
  REAL* b = A + (size_t)(m + n) * n; /* m+n */
  for (int j = 0; j < n; ++j) y[j] = ye[j] = y0[j];
  memset(r, 0, sizeof(REAL) * m);
  memset(J, 0, sizeof(REAL) * (size_t)m * n);
  REAL delta = 1;
  int first = 1, terminate = 0, steps = 0;
  int result = FN(allfinite)(y, n) ? 0 : 7;
  while (!terminate && result == 0 && steps < max_steps) {
    int y_old_finite = FN(allfinite)(y, n);
    FN(model_eval)(model, n, m, ye, x, d, re, Je);
    for (int j = 0; j < n; ++j) dd[j] = ye[j] - y[j];
    REAL rtr = 0, jt = 0, ree = 0, pred_exact = 0;
    for (int i = 0; i < m; ++i) {
      REAL s = r[i];
      for (int j = 0; j < n; ++j) s += J[i * n + j] * dd[j];
      jt += s * s;
      rtr += r[i] * r[i];
      ree += re[i] * re[i];
    }
    if (linsolve >= 2) {
      /* kernel arithmetic: |J d + r|^2 through the Gram identity r.r + (2 d.g + d.G d), then rounded and subtracted
       * like the reference's (optimistix_b200/csrc/common.cuh predicted_reduction) */
      REAL lin = 0, quad = 0;
      for (int a = 0; a < n; ++a) {
        REAL ga = 0, gd = 0;
        for (int i = 0; i < m; ++i) ga += J[i * n + a] * r[i];
        for (int c = 0; c < n; ++c) {
          REAL gac = 0;
          for (int i = 0; i < m; ++i) gac += J[i * n + a] * J[i * n + c];
          gd += gac * dd[c];
        }
        lin += dd[a] * ga;
        quad += dd[a] * gd;
      }
      volatile REAL t = rtr + ((REAL)2 * lin + quad);
      jt = t;
      if (linsolve == 3) jt = rtr; /* see below */
      pred_exact = lin + (REAL)0.5 * quad;
    }
    REAL pred = (REAL)0.5 * (jt - rtr);
    if (linsolve == 3) pred = pred_exact; /* d.g + 0.5 d.G d without the rounding to the spacing of |r|^2 (round-1 kernels) */
    REAL df = (REAL)0.5 * ree - (REAL)0.5 * rtr;
    int accept = (df <= (REAL)0.01 * pred) || first;
    int good = (df < (REAL)0.99 * pred) && (pred < 0) && !first;
    delta *= accept ? (good ? (REAL)3.5 : (REAL)1) : (REAL)0.25;
    if (accept) {
      REAL yn = 0, fn = 0;
      for (int j = 0; j < n; ++j) {
        REAL q = FABS(dd[j]) / (atol + rtol * FABS(ye[j]));
        if (!(q <= yn)) yn = q; /* NaN-propagating max */
      }
      for (int i = 0; i < m; ++i) {
        REAL q = FABS(re[i] - r[i]) / (atol + rtol * FABS(re[i]));
        if (!(q <= fn)) fn = q;
      }
      terminate = (yn < 1) && (fn < 1);
      for (int j = 0; j < n; ++j) y[j] = ye[j];
      memcpy(r, re, sizeof(REAL) * m);
      memcpy(J, Je, sizeof(REAL) * (size_t)m * n);
    } else {
      terminate = 0;
    }
    REAL lam = (delta > EPS) ? (REAL)1 / delta : RMAX;
    if (linsolve == 0) FN(solve_qr)(n, m, J, r, lam, A, b, p);
    else FN(solve_gram)(n, m, J, r, lam, A, p);
    for (int j = 0; j < n; ++j) ye[j] = y[j] - p[j];
    result = FN(allfinite)(p, n) ? 0 : 2;
    if (!y_old_finite) result = 7;
    first = 0;
    ++steps;
  }
  if (result == 0 && !terminate) result = 5;
  for (int j = 0; j < n; ++j) y_out[j] = y[j];
  *steps_out = steps;
  *result_out = result;
}

static int FN(lm_batch)(int model, int64_t batch, int n, int m, const REAL* y0, const REAL* x, const REAL* d,
                        int64_t d_stride, double rtol, double atol, int max_steps, int linsolve, REAL* y_out,
                        int32_t* steps, int32_t* result, int nthreads) {
  if (n > 16) return -1;
  size_t per = (size_t)2 * m + (size_t)2 * m * n + (size_t)(m + n) * n + (m + n);
  int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
  REAL* all = (REAL*)malloc(sizeof(REAL) * per * nt);
  if (!all) return -2;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t b = 0; b < batch; ++b) {
    REAL* scratch = all + per * omp_get_thread_num();
    FN(lm_one)(model, n, m, y0 + b * n, x, d + b * d_stride, (REAL)rtol, (REAL)atol, max_steps, linsolve, scratch,
               y_out + b * n, steps + b, result + b);
  }
  free(all);
  return 0;
}
#undef FN
#undef CAT
#undef CAT_
