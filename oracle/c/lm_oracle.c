/* CPU oracle in C (TEST INFRASTRUCTURE - see oracle/__init__.py): LevenbergMarquardt for the curve families.
 * Build: make -C oracle/c   ->  oracle/c/liboracle.so   (gcc -O2 -fopenmp; no -ffast-math: IEEE semantics) */
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

#define REAL double
#define SUFFIX _f64
#define EXP exp
#define SQRT sqrt
#define FABS fabs
#define EPS 2.220446049250313e-16
#define RMAX 1.7976931348623157e+308
#include "lm_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef EXP
#undef SQRT
#undef FABS
#undef EPS
#undef RMAX

#define REAL float
#define SUFFIX _f32
#define EXP expf
#define SQRT sqrtf
#define FABS fabsf
#define EPS 1.1920928955078125e-07f
#define RMAX 3.4028234663852886e+38f
#include "lm_oracle_impl.h"

/* dtype: 0 = f32, 1 = f64; model: 0 = exp_decay, 1 = gauss2; linsolve: 0 = QR (reference), 1 = Gram+Cholesky,
 * 2 = kernel arithmetic (Gram+Cholesky and the predicted reduction through the Gram identity, rounded like the reference's),
 * 3 = as 2 with the exactly evaluated d.g + 0.5 d.G d (the round-1 kernels; documents the iteration-count bias) */
int oracle_lm_solve(int model, int dtype, int64_t batch, int n, int m, const void* y0, const void* x, const void* d,
                    int64_t d_stride, double rtol, double atol, int max_steps, int linsolve, void* y_out, int32_t* steps,
                    int32_t* result, int nthreads) {
  if (dtype == 1)
    return lm_batch_f64(model, batch, n, m, (const double*)y0, (const double*)x, (const double*)d, d_stride, rtol, atol,
                        max_steps, linsolve, (double*)y_out, steps, result, nthreads);
  return lm_batch_f32(model, batch, n, m, (const float*)y0, (const float*)x, (const float*)d, d_stride, rtol, atol,
                      max_steps, linsolve, (float*)y_out, steps, result, nthreads);
}
int oracle_max_threads(void) { return omp_get_max_threads(); }
