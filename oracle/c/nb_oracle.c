/* CPU oracle in C (TEST INFRASTRUCTURE - see oracle/__init__.py): Newton / Chord on the implicit-Euler family and BFGS
 * (inverse update, backtracking Armijo) on the chained Rosenbrock function, double precision, one problem at a time,
 * parallel over problems with OpenMP.  Scalar restatements of the same reference steps as oracle/solvers.py
 * (newton_chord.py:69-196, quasi_newton.py:176-300,342-415, backtracking.py:51-99, _iterate.py:194-268, _misc.py:291-310),
 * max_norm only; used as the multi-core CPU baseline of bench.py for BASELINE configs[2] and configs[3] and cross-checked
 * against the NumPy oracle in tests/test_golden_oracle.py.  Never linked into the product library. */
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_max_threads(void) { return 1; }
static int omp_get_thread_num(void) { return 0; }
#endif
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { OK = 0, SINGULAR = 2, MAXSTEPS = 5, DIVERGED = 6, NONFINITE = 7 };

static int all_finite(const double* v, int n) {
  for (int i = 0; i < n; ++i)
    if (!isfinite(v[i])) return 0;
  return 1;
}
/* jnp.max semantics: NaN propagates */
static double nanmax(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

/* LU with partial pivoting (first maximal entry, column scaled by the reciprocal pivot as LAPACK getf2), then both solves.
 * A is n x n row-major and is overwritten; returns 0 if the solution is finite. */
static int lu_solve(int n, double* A, const double* b, double* x, int* piv) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (fabs(A[i * n + k]) > best) best = fabs(A[i * n + k]), p = i;
    piv[k] = p;
    if (p != k)
      for (int j = 0; j < n; ++j) {
        double t = A[k * n + j];
        A[k * n + j] = A[p * n + j];
        A[p * n + j] = t;
      }
    const double r = 1.0 / A[k * n + k];
    for (int i = k + 1; i < n; ++i) {
      const double l = A[i * n + k] * r;
      A[i * n + k] = l;
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
    }
  }
  for (int i = 0; i < n; ++i) x[i] = b[i];
  for (int k = 0; k < n; ++k) {
    if (piv[k] != k) {
      double t = x[k];
      x[k] = x[piv[k]];
      x[piv[k]] = t;
    }
    for (int i = k + 1; i < n; ++i) x[i] -= A[i * n + k] * x[k];
  }
  for (int k = n - 1; k >= 0; --k) {
    x[k] /= A[k * n + k];
    for (int i = 0; i < k; ++i) x[i] -= A[i * n + k] * x[k];
  }
  return all_finite(x, n) ? 0 : 1;
}

/* implicit Euler: F(y) = y - y_prev - dt (A y + y*(B y)); J = I - dt (A + diag(B y) + diag(y) B) */
static void ie_eval(int n, const double* A, const double* B, const double* y_prev, double dt, const double* y, double* F, double* J) {
  for (int i = 0; i < n; ++i) {
    double Ay = 0, By = 0;
    for (int j = 0; j < n; ++j) Ay += A[i * n + j] * y[j], By += B[i * n + j] * y[j];
    F[i] = y[i] - y_prev[i] - dt * (Ay + y[i] * By);
    if (J)
      for (int j = 0; j < n; ++j) J[i * n + j] = (i == j ? 1.0 : 0.0) - dt * (A[i * n + j] + (i == j ? By : 0.0) + y[i] * B[i * n + j]);
  }
}

int oracle_newton_implicit_euler(int64_t batch, int n, const double* A, const double* B, const double* y_prev, const double* dt,
                                 const double* y0, double rtol, double atol, int max_steps, int chord, double* y_out, int32_t* steps_out,
                                 int32_t* result_out, int nthreads) {
  const int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
  const size_t per = (size_t)3 * n * n + 4 * n;
  double* all = (double*)malloc(sizeof(double) * per * nt);
  int* pivs = (int*)malloc(sizeof(int) * n * nt);
  if (!all || !pivs) return -1;
#pragma omp parallel for schedule(dynamic, 64) num_threads(nt)
  for (int64_t b = 0; b < batch; ++b) {
    double* w = all + per * omp_get_thread_num();
    double *J = w, *Jfix = J + n * n, *LU = Jfix + n * n, *F = LU + n * n, *y = F + n, *sol = y + n, *newy = sol + n;
    int* piv = pivs + n * omp_get_thread_num();
    memcpy(y, y0 + b * n, sizeof(double) * n);
    const double* yp = y_prev + b * n;
    int result = all_finite(y, n) ? OK : NONFINITE, terminate = 0, steps = 0;
    if (chord) ie_eval(n, A, B, yp, dt[b], y, F, Jfix); /* newton_chord.py:80-87 */
    while (!terminate && result == OK && steps < max_steps) {
      const int y_old_finite = all_finite(y, n);
      ie_eval(n, A, B, yp, dt[b], y, F, chord ? NULL : J);
      memcpy(LU, chord ? Jfix : J, sizeof(double) * n * n);
      const int fail = lu_solve(n, LU, F, sol, piv);
      double yn = 0, fn = 0;
      for (int i = 0; i < n; ++i) {
        newy[i] = y[i] - sol[i];
        yn = nanmax(yn, fabs(sol[i]) / (atol + rtol * fabs(y[i]))); /* cauchy_termination with the pre-step y (:160-176) */
        fn = nanmax(fn, fabs(F[i]) / atol);
      }
      terminate = (yn < 1.0 && fn < 1.0) || fail;
      result = fail ? SINGULAR : OK;
      if (!y_old_finite) result = NONFINITE;
      memcpy(y, newy, sizeof(double) * n);
      ++steps;
    }
    if (result == OK && !terminate) result = MAXSTEPS;
    memcpy(y_out + b * n, y, sizeof(double) * n);
    steps_out[b] = steps;
    result_out[b] = result;
  }
  free(all);
  free(pivs);
  return 0;
}

static double rosen(int n, const double* y, double* g) {
  double f = 0;
  for (int i = 0; i < n; ++i) g[i] = 0;
  for (int i = 0; i + 1 < n; ++i) {
    const double a = y[i + 1] - y[i] * y[i], b = 1.0 - y[i];
    f += 100.0 * a * a + b * b;
    g[i] += -400.0 * a * y[i] - 2.0 * b;
    g[i + 1] += 200.0 * a;
  }
  return f;
}

int oracle_bfgs_rosenbrock(int64_t batch, int n, const double* y0, double rtol, double atol, int max_steps, double* y_out,
                           int32_t* steps_out, int32_t* result_out, int nthreads) {
  const int nt = nthreads > 0 ? nthreads : omp_get_max_threads();
  const size_t per = (size_t)n * n + 8 * n;
  double* all = (double*)malloc(sizeof(double) * per * nt);
  if (!all) return -1;
  const double eps = 2.220446049250313e-16;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
  for (int64_t b = 0; b < batch; ++b) {
    double* w = all + per * omp_get_thread_num();
    double *H = w, *y = H + n * n, *ye = y + n, *g = ye + n, *ge = g + n, *nt_ = ge + n, *d = nt_ + n, *q = d + n, *Hq = q + n;
    for (int i = 0; i < n * n; ++i) H[i] = 0;
    for (int i = 0; i < n; ++i) H[i * n + i] = 1, y[i] = ye[i] = y0[b * n + i], g[i] = 0, nt_[i] = 0;
    double f = 0, alpha = 1;
    int first = 1, terminate = 0, steps = 0, result = all_finite(y, n) ? OK : NONFINITE;
    while (!terminate && result == OK && steps < max_steps) {
      const int y_old_finite = all_finite(y, n);
      const double fe = rosen(n, ye, ge);
      double pred = 0;
      for (int i = 0; i < n; ++i) d[i] = ye[i] - y[i], pred += g[i] * d[i];
      const double df = fe - f;
      const int accept = first || (df <= 0.1 * pred && pred <= 0);   /* backtracking.py:79-93 */
      alpha = accept ? 1.0 : 0.5 * alpha;
      if (accept) {
        double inner = 0;
        for (int i = 0; i < n; ++i) q[i] = ge[i] - g[i], inner += q[i] * d[i];
        if (inner > eps) { /* quasi_newton.py:360,371-385 */
          double mvp_inner = 0;
          for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int j = 0; j < n; ++j) s += H[i * n + j] * q[j];
            Hq[i] = s;
          }
          for (int i = 0; i < n; ++i) mvp_inner += q[i] * Hq[i];
          for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j)
              H[i * n + j] = H[i * n + j] + ((inner + mvp_inner) * (d[i] * d[j])) / (inner * inner) - (Hq[i] * d[j] + d[i] * Hq[j]) / inner;
        }
        for (int i = 0; i < n; ++i) {
          double s = 0;
          for (int j = 0; j < n; ++j) s += H[i * n + j] * ge[j];
          nt_[i] = s; /* gauss_newton.py:60-61 */
        }
        double yn = 0;
        for (int i = 0; i < n; ++i) yn = nanmax(yn, fabs(d[i]) / (atol + rtol * fabs(ye[i])));
        const double fnm = fabs(df) / (atol + rtol * fabs(fe));
        terminate = (yn < 1.0 && fnm < 1.0) && !first; /* quasi_newton.py:244-249 */
        for (int i = 0; i < n; ++i) y[i] = ye[i], g[i] = ge[i];
        f = fe;
      } else {
        terminate = 0;
      }
      for (int i = 0; i < n; ++i) ye[i] = y[i] - alpha * nt_[i];
      result = OK;
      if (!y_old_finite) result = NONFINITE;
      first = 0;
      ++steps;
    }
    if (result == OK && !terminate) result = MAXSTEPS;
    memcpy(y_out + b * n, y, sizeof(double) * n);
    steps_out[b] = steps;
    result_out[b] = result;
  }
  free(all);
  return 0;
}
