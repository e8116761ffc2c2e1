// Residual-model plugin registry, device side: the compiled-in model families.
//
// The reference has no "model family": its nearest concept is the user callable fn(y, args) whose
// Jacobian comes from jax.linearize (optimistix/_solver/gauss_newton.py:158-181).  Here a family is a
// struct of __device__ functions selected at launch by model_id; the analytic Jacobian below is what
// jax.linearize evaluates to for the same fn.  tests/ check each family against oracle/models.py.
//
// Least-squares "curve" families share one calling convention:
//   shared_args  = abscissa grid x[m]           (shared by every problem)
//   problem_args = data d[m] for this problem
//   Row          = (x_i, d_i), kept in registers by the solver for its rows
//   prepare(y)   = per-trial-point precomputation (e.g. 1/sigma), done once per evaluation
//   residual(y, pre, row, aux) -> r_i   and caches per-row intermediates (the exponentials) in aux[NAUX]
//   jac_row(y, pre, row, aux, j[N])     -> d r_i / d y
#pragma once
#include "common.cuh"

namespace b200opt {

template <typename T> struct CurveRow {
  T x;
  T d;
};

// id 0: r_i = a*exp(-b*t_i) + c - d_i,  y = [a, b, c]
template <typename T> struct ExpDecay {
  static constexpr int N = 3;
  static constexpr int NAUX = 1;
  struct Pre {};
  __device__ static __forceinline__ Pre prepare(const T (&)[N]) { return Pre{}; }
  __device__ static __forceinline__ T residual(const T (&y)[N], const Pre&, const CurveRow<T>& row, T (&aux)[NAUX]) {
    const T e = t_exp(-y[1] * row.x);
    aux[0] = e;
    return y[0] * e + y[2] - row.d;
  }
  __device__ static __forceinline__ void jac_row(const T (&y)[N], const Pre&, const CurveRow<T>& row,
                                                 const T (&aux)[NAUX], T (&j)[N]) {
    j[0] = aux[0];
    j[1] = -y[0] * row.x * aux[0];
    j[2] = T(1);
  }
};

// id 1: two Gaussians + linear baseline, y = [A1, mu1, s1, A2, mu2, s2, b0, b1]
template <typename T> struct Gauss2 {
  static constexpr int N = 8;
  static constexpr int NAUX = 2;  // e1, e2 (z is cheaper to recompute than to keep)
  struct Pre {
    T i1, i2;
  };
  __device__ static __forceinline__ Pre prepare(const T (&y)[N]) { return Pre{T(1) / y[2], T(1) / y[5]}; }
  __device__ static __forceinline__ T residual(const T (&y)[N], const Pre& p, const CurveRow<T>& row, T (&aux)[NAUX]) {
    const T z1 = (row.x - y[1]) * p.i1;
    const T z2 = (row.x - y[4]) * p.i2;
    const T e1 = t_exp(T(-0.5) * z1 * z1);
    const T e2 = t_exp(T(-0.5) * z2 * z2);
    aux[0] = e1;
    aux[1] = e2;
    return y[0] * e1 + y[3] * e2 + y[6] + y[7] * row.x - row.d;
  }
  __device__ static __forceinline__ void jac_row(const T (&y)[N], const Pre& p, const CurveRow<T>& row,
                                                 const T (&aux)[NAUX], T (&j)[N]) {
    const T z1 = (row.x - y[1]) * p.i1;
    const T z2 = (row.x - y[4]) * p.i2;
    const T e1 = aux[0], e2 = aux[1];
    j[0] = e1;
    j[1] = y[0] * e1 * z1 * p.i1;
    j[2] = y[0] * e1 * z1 * z1 * p.i1;
    j[3] = e2;
    j[4] = y[3] * e2 * z2 * p.i2;
    j[5] = y[3] * e2 * z2 * z2 * p.i2;
    j[6] = T(1);
    j[7] = row.x;
  }
};

// model ids (keep in sync with the table in capi.cu and optimistix_b200/_models.py)
enum ModelId : int {
  MODEL_EXP_DECAY = 0,
  MODEL_GAUSS2 = 1,
  MODEL_IMPLICIT_EULER = 2,
  MODEL_ROSENBROCK = 3,
  MODEL_LINEAR_ROWS = 4,
  MODEL_COUNT = 5,
};

}  // namespace b200opt
