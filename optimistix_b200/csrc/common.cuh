// Shared device helpers for the B200 (sm_100a) Optimistix hot path.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#endif

#include "../../include/b200opt.h"

namespace b200opt {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

enum Result : int {
  RES_OK = B200OPT_RESULT_SUCCESSFUL,
  RES_SINGULAR = B200OPT_RESULT_SINGULAR,
  RES_NL_MAX_STEPS = B200OPT_RESULT_NONLINEAR_MAX_STEPS_REACHED,
  RES_DIVERGED = B200OPT_RESULT_NONLINEAR_DIVERGENCE,
  RES_NONFINITE = B200OPT_RESULT_NONFINITE,
};

// ---- numeric traits (jnp.finfo) ---------------------------------------------------------------------------
template <typename T> struct Num;
template <> struct Num<float> {
  static __host__ __device__ constexpr float eps() { return 1.1920928955078125e-07f; }
  static __host__ __device__ constexpr float max() { return 3.4028234663852886e+38f; }
#ifdef __CUDACC_RTC__
  static __device__ float inf() { return __int_as_float(0x7f800000); }
#else
  static __host__ __device__ constexpr float inf() { return __builtin_huge_valf(); }
#endif
  static constexpr int precision = 6;  // jnp.finfo(float32).precision
  using Vec = float4;
  static constexpr int kVec = 4;
};
template <> struct Num<double> {
  static __host__ __device__ constexpr double eps() { return 2.220446049250313e-16; }
  static __host__ __device__ constexpr double max() { return 1.7976931348623157e+308; }
#ifdef __CUDACC_RTC__
  static __device__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
#else
  static __host__ __device__ constexpr double inf() { return __builtin_huge_val(); }
#endif
  static constexpr int precision = 15;
  using Vec = double2;
  static constexpr int kVec = 2;
};

__device__ __forceinline__ float t_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float t_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double t_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float t_exp(float x) { return expf(x); }
__device__ __forceinline__ double t_exp(double x) { return exp(x); }
__device__ __forceinline__ float t_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double t_abs(double x) { return fabs(x); }
__device__ __forceinline__ float t_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double t_tanh(double x) { return tanh(x); }
template <typename T> __device__ __forceinline__ bool t_finite(T x) { return isfinite(x); }

// jnp.max semantics: NaN propagates (fmax would drop it and wrongly report convergence).
template <typename T> __device__ __forceinline__ T nan_max(T a, T b) {
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}

// ---- warp collectives (deterministic butterfly; every lane ends with the same value) -------------------------
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
template <typename T> __device__ __forceinline__ T warp_nanmax(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// NaN-propagating max of NON-NEGATIVE values (x >= +0, +inf or NaN - what the scaled norms below feed it): such floats
// order like their bit patterns read as unsigned integers, with every NaN above +inf, so the warp maximum is one integer
// REDUX (two for a double: high words, then the low words of the lanes that hold the maximal high word) instead of five
// shuffle rounds of compare-and-select.  Exact, like any max.
__device__ __forceinline__ float warp_nanmax_nonneg(float v) {
  return __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(v)));
}
__device__ __forceinline__ double warp_nanmax_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(kFull, hi);
  const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}

// ---- norms over a small per-lane-replicated vector and over warp-distributed rows -----------------------------
// acc_* implement max_norm / two_norm / rms_norm (lineax.internal, re-exported at optimistix/_misc.py:41-45)
template <typename T> struct NormAcc {
  T v;
  int kind;
  __device__ __forceinline__ explicit NormAcc(int k) : v(T(0)), kind(k) {}
  __device__ __forceinline__ void add(T x) {  // x >= 0 or NaN
    if (kind == B200OPT_NORM_MAX) v = nan_max(v, x);
    else v += x * x;
  }
  // combine across the warp (rows distributed over lanes)
  __device__ __forceinline__ void warp_combine() {
    if (kind == B200OPT_NORM_MAX) v = warp_nanmax_nonneg(v);
    else v = warp_sum(v);
  }
  __device__ __forceinline__ T finish(int count) const {
    if (kind == B200OPT_NORM_MAX) return v;
    if (kind == B200OPT_NORM_TWO) return t_sqrt(v);
    return t_sqrt(v / T(count));
  }
};

// ---- dense SPD solve in registers: (G + lam I) p = g, G packed lower-triangular, fully unrolled ---------------
template <int N> struct Tri {
  static constexpr int size = N * (N + 1) / 2;
  static __host__ __device__ constexpr int at(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
};

template <typename T, int N>
__device__ __forceinline__ void spd_solve(const T (&G)[Tri<N>::size], T lam, const T (&g)[N], T (&p)[N]) {
  T L[Tri<N>::size];  // strictly-lower entries of the Cholesky factor; diagonal slot holds 1/L_jj
#pragma unroll
  for (int j = 0; j < N; ++j) {
    T s = G[Tri<N>::at(j, j)] + lam;
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[Tri<N>::at(j, k)] * L[Tri<N>::at(j, k)];
    const T inv = t_rsqrt(s);  // s < 0 or NaN -> NaN, s == 0 -> inf: reported as `singular` by the caller
    L[Tri<N>::at(j, j)] = inv;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      T t = G[Tri<N>::at(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[Tri<N>::at(i, k)] * L[Tri<N>::at(j, k)];
      L[Tri<N>::at(i, j)] = t * inv;
    }
  }
  T z[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T s = g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[Tri<N>::at(i, k)] * z[k];
    z[i] = s * L[Tri<N>::at(i, i)];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    T s = z[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s -= L[Tri<N>::at(k, i)] * p[k];
    p[i] = s * L[Tri<N>::at(i, i)];
  }
}

#ifndef __CUDACC_RTC__
// thread-local last-error string (host side)
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
#endif

}  // namespace b200opt
