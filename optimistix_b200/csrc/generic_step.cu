// Generic per-iteration steps for an ARBITRARY user function, root finding and minimisation (the least-squares
// counterpart is lsq_step.cu).  The caller (XLA, torch, ...) evaluates the function and its derivative at the current
// point; these kernels do everything else of one solver step plus the driver bookkeeping of _iterate.py:234-241 for
// every problem that is still running.
//
//   root_step   one _AbstractNewtonChord.step + terminate (optimistix/_solver/newton_chord.py:103-196): in-warp LU with
//               partial pivoting of the caller's Jacobian (Newton: J(y); Chord: J(y0), passed again on every call),
//               y <- y - J^-1 f, optional clip to [lower, upper], scaled step size, Cauchy or Hairer-Wanner
//               termination, result codes.  One warp per problem, n <= 32.
//   bfgs_step   one AbstractQuasiNewton.step (quasi_newton.py:204-300) with BacktrackingArmijo (backtracking.py:79-99),
//               the inverse BFGS update (quasi_newton.py:342-415, skipped when q.s <= eps) and NewtonDescent
//               (gauss_newton.py:60-62,126-127).  One CTA per problem; H^-1 (n x n) is carried in the caller-owned state
//               buffer in HBM between calls (a per-iteration kernel cannot keep it on chip), read twice and written
//               once per ACCEPTED step, untouched by a rejected one.
#include "newton_kernel.cuh"

namespace b200opt {

// =============================================================================================== root step
// state per problem (elements of T): [diffsize | nstep | terminate | pad]
constexpr int kRootStateElems = 4;

template <typename T> struct RootStepParams {
  const T* f_eval;  // [batch][n]      fn(y)
  const T* J_eval;  // [batch][n][n]   d fn / d y  (row-major)
  const T* lower;   // optional [n]
  const T* upper;   // optional [n]
  T* y;             // [batch][n] in/out
  T* state;
  int32_t* num_steps;
  int32_t* result;
  int32_t* active_count;
  int64_t batch;
  int n;
  T rtol, atol, kappa;
  int norm, max_steps, cauchy;
};

constexpr int kRootStepWarps = 4;

template <typename T, int N> __global__ void __launch_bounds__(kRootStepWarps * 32) root_step_kernel(const RootStepParams<T> p) {
  __shared__ __align__(16) T rowbuf_all[kRootStepWarps][N + 4];
  __shared__ unsigned char piv_all[kRootStepWarps][N];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t prob = (int64_t)blockIdx.x * kRootStepWarps + wib;
  if (prob >= p.batch) return;
  const int n = p.n;
  T* st = p.state + prob * kRootStateElems;
  int result = p.result[prob];
  int steps = p.num_steps[prob];
  bool terminate = st[2] != T(0);
  if (terminate || result != RES_OK || steps >= p.max_steps) return;  // this lane of the vmapped loop is frozen
  const bool act = lane < n;
  const T rtol = p.rtol, atol = p.atol;
  T small_tol = T(1);  // 10^(2 - precision), newton_chord.py:25-29
  for (int i = 0; i < Num<T>::precision - 2; ++i) small_tol *= T(0.1);

  const T y = act ? p.y[prob * n + lane] : T(0);
  const T fx = act ? p.f_eval[prob * n + lane] : T(0);
  const bool y_old_finite = __all_sync(kFull, !act || t_finite(y));
  // row `lane` of the Jacobian, padded with the identity to N x N
  T a[N];
#pragma unroll
  for (int j = 0; j < N; ++j) a[j] = (act && j < n) ? p.J_eval[(prob * n + lane) * (int64_t)n + j] : ((j == lane) ? T(1) : T(0));
  int mycol;
  T rinv;
  warp_lu_factor<T, N>(a, mycol, rinv, rowbuf_all[wib], piv_all[wib], lane);
  const T sol = warp_lu_solve<T, N>(a, mycol, rinv, fx, piv_all[wib], lane);
  const int lin_result = __all_sync(kFull, !act || t_finite(sol)) ? RES_OK : RES_SINGULAR;  // lineax: non-finite => singular
  const T lo = (p.lower && act) ? p.lower[lane] : -Num<T>::inf();
  const T hi = (p.upper && act) ? p.upper[lane] : Num<T>::inf();
  T new_y = y - sol;  // newton_chord.py:129
  T d = sol;
  if (p.lower != nullptr || p.upper != nullptr) {
    new_y = new_y < lo ? lo : new_y;  // :130-133
    new_y = new_y > hi ? hi : new_y;
    d = y - new_y;  // :134-135
  }
  NormAcc<T> ds(p.norm);
  if (act) ds.add(t_abs(d / (atol + rtol * t_abs(new_y))));  // :136-138
  ds.warp_combine();
  const T diffsize_prev = st[0];
  const T diffsize = ds.finish(n);
  const int nstep = (int)st[1] + 1;
  // ---- terminate, evaluated with the PRE-step y (:154-196, _iterate.py:238)
  int term_result = RES_OK;
  if (p.cauchy) {
    NormAcc<T> yn(p.norm), fn(p.norm);
    if (act) {
      yn.add(t_abs(d) / (atol + rtol * t_abs(y)));
      fn.add(t_abs(fx) / atol);  // f_scale = atol + rtol * |0|
    }
    yn.warp_combine();
    fn.warp_combine();
    terminate = (yn.finish(n) < T(1)) && (fn.finish(n) < T(1));
  } else {
    const bool at_least_two = nstep >= 2;
    const T rate = diffsize / diffsize_prev;
    const T factor = diffsize * rate / (T(1) - rate);
    const bool small = diffsize < small_tol;
    const bool diverged = !t_finite(rate) || rate > T(2);
    const bool converged = factor > T(0) && factor < p.kappa;
    terminate = at_least_two && (small || diverged || converged);
    if (at_least_two && !small && diverged) term_result = RES_DIVERGED;
  }
  const bool fail = lin_result != RES_OK;
  result = fail ? lin_result : term_result;
  terminate = terminate || fail;
  if (!y_old_finite) result = RES_NONFINITE;
  ++steps;
  if (act) p.y[prob * n + lane] = new_y;
  if (lane == 0) {
    st[0] = diffsize;
    st[1] = (T)nstep;
    st[2] = terminate ? T(1) : T(0);
    p.num_steps[prob] = steps;
    p.result[prob] = result;
    if (p.active_count && !terminate && result == RES_OK && steps < p.max_steps) atomicAdd(p.active_count, 1);
  }
}

// ---- n > 32: one CTA per problem, the Jacobian in shared memory -----------------------------------------------------------
// LU with partial pivoting (first maximal entry, multiplication by the pivot's reciprocal: the arithmetic of the in-warp
// kernels and of getf2) with the forward substitution fused into the elimination, then a column-oriented back substitution.
// Three block barriers per column: the generic path is host-latency bound anyway (torch evaluates f and J every step).
constexpr int kBigRootThreads = 256;
__host__ __device__ inline int big_root_ld(int n) { return n | 1; }  // odd pitch: conflict-free column access
template <typename T> __host__ __device__ inline size_t big_root_smem(int n) {
  return sizeof(T) * ((size_t)n * big_root_ld(n) + 2 * (size_t)n + 2 * (kBigRootThreads / 32) + 8);
}
// block-wide NaN-propagating max / sum of one value per thread (all threads get the result)
template <typename T> __device__ __forceinline__ T block_combine(T v, bool is_max, T* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = is_max ? warp_nanmax(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  T r = red[0];
  for (int i = 1; i < kBigRootThreads / 32; ++i) r = is_max ? nan_max(r, red[i]) : r + red[i];
  return r;
}
// A x = b in shared memory (A [n][ld] overwritten by its factors, b overwritten by the solution)
template <typename T> __device__ void cta_lu_solve(T* A, int ld, T* b, int n, T* red, int* ired) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int k = 0; k < n; ++k) {
    // pivot: first maximal |A[i][k]|, i >= k
    T best = T(-1);
    int bi = n;
    for (int i = k + tid; i < n; i += kBigRootThreads) {
      const T v = t_abs(A[i * ld + k]);
      if (v > best || (v != v && best == best)) best = v, bi = i;  // strided scan: indices increase, so > keeps the first; NaN wins
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(kFull, best, o);
      const int oi = __shfl_xor_sync(kFull, bi, o);
      const bool take = (ov != ov && best == best) || (!(best != best) && (ov > best || (ov == best && oi < bi)));
      if (take) best = ov, bi = oi;
    }
    if (lane == 0) red[w] = best, ired[w] = bi;
    __syncthreads();
    best = red[0];
    bi = ired[0];
    for (int i = 1; i < kBigRootThreads / 32; ++i) {
      const T ov = red[i];
      const int oi = ired[i];
      const bool take = (ov != ov && best == best) || (!(best != best) && (ov > best || (ov == best && oi < bi)));
      if (take) best = ov, bi = oi;
    }
    const int pr = bi < n ? bi : k;
    __syncthreads();
    if (pr != k) {  // swap rows k and pr (and the right-hand side)
      for (int j = tid; j < n; j += kBigRootThreads) {
        const T t = A[k * ld + j];
        A[k * ld + j] = A[pr * ld + j];
        A[pr * ld + j] = t;
      }
      if (tid == 0) {
        const T t = b[k];
        b[k] = b[pr];
        b[pr] = t;
      }
      __syncthreads();
    }
    const T rinv = T(1) / A[k * ld + k];
    const T bk = b[k];
    // multipliers and trailing update, row by row: thread t takes rows k + 1 + t, ...
    for (int i = k + 1 + w; i < n; i += kBigRootThreads / 32) {
      const T l = A[i * ld + k] * rinv;
      for (int j = k + 1 + lane; j < n; j += 32) A[i * ld + j] -= l * A[k * ld + j];
      if (lane == 0) {
        A[i * ld + k] = l;
        b[i] -= l * bk;
      }
    }
    __syncthreads();
  }
  for (int k = n - 1; k >= 0; --k) {
    if (tid == 0) b[k] = b[k] / A[k * ld + k];
    __syncthreads();
    const T xk = b[k];
    for (int i = tid; i < k; i += kBigRootThreads) b[i] -= A[i * ld + k] * xk;
    __syncthreads();
  }
}

template <typename T> __global__ void __launch_bounds__(kBigRootThreads) root_step_big_kernel(const RootStepParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, ld = big_root_ld(n), tid = threadIdx.x;
  T* A = reinterpret_cast<T*>(smem_raw);
  T* b = A + (size_t)n * ld;
  T* yv = b + n;
  T* red = yv + n;
  int* ired = reinterpret_cast<int*>(red + kBigRootThreads / 32);
  const int64_t prob = blockIdx.x;
  T* st = p.state + prob * kRootStateElems;
  int result = p.result[prob];
  int steps = p.num_steps[prob];
  bool terminate = st[2] != T(0);
  if (terminate || result != RES_OK || steps >= p.max_steps) return;  // frozen (block-uniform)
  const T rtol = p.rtol, atol = p.atol;
  T small_tol = T(1);
  for (int i = 0; i < Num<T>::precision - 2; ++i) small_tol *= T(0.1);
  bool fin = true;
  for (int i = tid; i < n; i += kBigRootThreads) {
    yv[i] = p.y[prob * n + i];
    b[i] = p.f_eval[prob * n + i];
    fin = fin && t_finite(yv[i]);
  }
  for (int e = tid; e < n * n; e += kBigRootThreads) A[(e / n) * ld + e % n] = p.J_eval[prob * (int64_t)n * n + e];
  const bool y_old_finite = __syncthreads_and(fin ? 1 : 0) != 0;
  // the pre-step f is needed by the Cauchy test after b has become the solution
  T fnacc = T(0);
  if (p.cauchy)
    for (int i = tid; i < n; i += kBigRootThreads) {
      const T q = t_abs(b[i]) / atol;
      fnacc = (p.norm == B200OPT_NORM_MAX) ? nan_max(fnacc, q) : fnacc + q * q;
    }
  cta_lu_solve<T>(A, ld, b, n, red, ired);
  bool solfin = true;
  T dsacc = T(0), ynacc = T(0);
  for (int i = tid; i < n; i += kBigRootThreads) {
    const T sol = b[i], y = yv[i];
    solfin = solfin && t_finite(sol);
    T new_y = y - sol, d = sol;
    if (p.lower != nullptr || p.upper != nullptr) {
      const T lo = p.lower ? p.lower[i] : -Num<T>::inf(), hi = p.upper ? p.upper[i] : Num<T>::inf();
      new_y = new_y < lo ? lo : new_y;
      new_y = new_y > hi ? hi : new_y;
      d = y - new_y;
    }
    const T q = t_abs(d / (atol + rtol * t_abs(new_y)));
    dsacc = (p.norm == B200OPT_NORM_MAX) ? nan_max(dsacc, q) : dsacc + q * q;
    const T qy = t_abs(d) / (atol + rtol * t_abs(y));
    ynacc = (p.norm == B200OPT_NORM_MAX) ? nan_max(ynacc, qy) : ynacc + qy * qy;
    p.y[prob * n + i] = new_y;
  }
  const int lin_result = __syncthreads_and(solfin ? 1 : 0) ? RES_OK : RES_SINGULAR;
  const bool is_max = p.norm == B200OPT_NORM_MAX;
  auto finish = [&](T v) { return is_max ? v : (p.norm == B200OPT_NORM_TWO ? t_sqrt(v) : t_sqrt(v / T(n))); };
  const T diffsize = finish(block_combine(dsacc, is_max, red));
  const T diffsize_prev = st[0];
  const int nstep = (int)st[1] + 1;
  int term_result = RES_OK;
  if (p.cauchy) {
    const T yn = finish(block_combine(ynacc, is_max, red));
    const T fn = finish(block_combine(fnacc, is_max, red));
    terminate = (yn < T(1)) && (fn < T(1));
  } else {
    const bool at_least_two = nstep >= 2;
    const T rate = diffsize / diffsize_prev;
    const T factor = diffsize * rate / (T(1) - rate);
    const bool small = diffsize < small_tol;
    const bool diverged = !t_finite(rate) || rate > T(2);
    const bool converged = factor > T(0) && factor < p.kappa;
    terminate = at_least_two && (small || diverged || converged);
    if (at_least_two && !small && diverged) term_result = RES_DIVERGED;
  }
  const bool fail = lin_result != RES_OK;
  result = fail ? lin_result : term_result;
  terminate = terminate || fail;
  if (!y_old_finite) result = RES_NONFINITE;
  ++steps;
  __syncthreads();  // every thread has read the state before it is rewritten
  if (tid == 0) {
    st[0] = diffsize;
    st[1] = (T)nstep;
    st[2] = terminate ? T(1) : T(0);
    p.num_steps[prob] = steps;
    p.result[prob] = result;
    if (p.active_count && !terminate && result == RES_OK && steps < p.max_steps) atomicAdd(p.active_count, 1);
  }
}

template <typename T> __global__ void __launch_bounds__(kBigRootThreads) lu_solve_big_kernel(const T* Ain, const T* bin, T* x, int32_t* status, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ld = big_root_ld(n), tid = threadIdx.x;
  T* A = reinterpret_cast<T*>(smem_raw);
  T* b = A + (size_t)n * ld;
  T* red = b + 2 * n;
  int* ired = reinterpret_cast<int*>(red + kBigRootThreads / 32);
  const int64_t prob = blockIdx.x;
  for (int i = tid; i < n; i += kBigRootThreads) b[i] = bin[prob * n + i];
  for (int e = tid; e < n * n; e += kBigRootThreads) A[(e / n) * ld + e % n] = Ain[prob * (int64_t)n * n + e];
  __syncthreads();
  cta_lu_solve<T>(A, ld, b, n, red, ired);
  bool fin = true;
  for (int i = tid; i < n; i += kBigRootThreads) {
    x[prob * n + i] = b[i];
    fin = fin && t_finite(b[i]);
  }
  const int ok = __syncthreads_and(fin ? 1 : 0);
  if (tid == 0 && status) status[prob] = ok ? RES_OK : RES_SINGULAR;
}

template <typename T> static int big_root_nmax() { return sizeof(T) == 8 ? 160 : 224; }  // n (n | 1) + 2 n elements within 227 KB

template <typename T> __global__ void root_step_init_kernel(int64_t batch, int n, const T* y0, T* y, T* state, int32_t* num_steps,
                                                            int32_t* result) {
  const int64_t prob = blockIdx.x;
  if (prob >= batch) return;
  bool fin = true;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const T v = y0[prob * n + j];
    y[prob * n + j] = v;
    fin = fin && t_finite(v);
  }
  const int allfin = __syncthreads_and(fin ? 1 : 0);
  if (threadIdx.x == 0) {
    T* st = state + prob * kRootStateElems;
    st[0] = Num<T>::inf();  // diffsize (:93), so that diffsize_prev of the second step is inf as in the reference carry
    st[1] = T(0);           // step
    st[2] = T(0);           // terminate
    st[3] = T(0);
    num_steps[prob] = 0;
    result[prob] = allfin ? RES_OK : RES_NONFINITE;  // _iterate.py:216-219
  }
}

// =============================================================================================== batched LU solve
// A x = b for `batch` independent n x n systems (n <= 32): the linear solve of ImplicitAdjoint (optimistix/_ad.py:83-118,
// lx.AutoLinearSolver(well_posed=None) on a square operator = LU with partial pivoting).  One warp per system, the same
// row-per-lane factorisation as the Newton kernels; status = singular when the solution is not finite.
template <typename T, int N>
__global__ void __launch_bounds__(kRootStepWarps * 32) lu_solve_kernel(const T* A, const T* b, T* x, int32_t* status, int64_t batch, int n) {
  __shared__ __align__(16) T rowbuf_all[kRootStepWarps][N + 4];
  __shared__ unsigned char piv_all[kRootStepWarps][N];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t prob = (int64_t)blockIdx.x * kRootStepWarps + wib;
  if (prob >= batch) return;
  const bool act = lane < n;
  T a[N];
#pragma unroll
  for (int j = 0; j < N; ++j) a[j] = (act && j < n) ? A[(prob * n + lane) * (int64_t)n + j] : ((j == lane) ? T(1) : T(0));
  const T rhs = act ? b[prob * n + lane] : T(0);
  int mycol;
  T rinv;
  warp_lu_factor<T, N>(a, mycol, rinv, rowbuf_all[wib], piv_all[wib], lane);
  const T sol = warp_lu_solve<T, N>(a, mycol, rinv, rhs, piv_all[wib], lane);
  const bool ok = __all_sync(kFull, !act || t_finite(sol));
  if (act) x[prob * n + lane] = sol;
  if (lane == 0 && status) status[prob] = ok ? RES_OK : RES_SINGULAR;
}

// =============================================================================================== BFGS step
// state per problem (elements of T):
//   [M (n*n: H^-1, or H when use_inverse=False) | g (n) | newton (n) | f | alpha | first | terminate | num_accepted | descent result | pad..]
// Update variants (quasi_newton.py:342-415 BFGS, :487-573 DFP), with s = y_eval - y, q = grad_new - grad, inner = q.s:
//   form A:  M += ((inner + v.Mv) / inner^2) u u^T - (Mv u^T + u Mv^T) / inner     BFGS inverse (u=s, v=q), DFP direct (u=q, v=s)
//   form B:  M += u u^T / inner - Mv Mv^T / (v.Mv)                                 DFP inverse (u=s, v=q), BFGS direct (u=q, v=s)
// Direct variants (use_inverse=False) get the direction from a Cholesky solve M newton = grad (NewtonDescent with
// lx.Cholesky(), quasi_newton.py:455; n <= 64, factorised in shared memory); a non-finite solution is `singular`.
constexpr int B200OPT_QN_FORM_B = 1, B200OPT_QN_SWAP = 2, B200OPT_QN_DIRECT = 4;
constexpr int kBfgsDirectNMax = 64;
__host__ __device__ inline int64_t bfgs_state_stride(int n) { return ((int64_t)n * n + 2 * n + 6 + 3) / 4 * 4; }

template <typename T> struct BfgsStepParams {
  const T* f_eval;       // [batch]     fn(y_eval)
  const T* g_eval;       // [batch][n]  grad fn(y_eval)
  const T* root_norm;    // optional [batch]: norm of the root function at the pre-step y (root_find through a minimiser)
  T* y;                  // [batch][n] in/out: last accepted point
  T* y_eval;             // [batch][n] in/out: next trial point
  T* state;
  int64_t state_stride;
  int32_t* num_steps;
  int32_t* result;
  int32_t* active_count;
  int64_t batch;
  int n;
  T rtol, atol;
  int norm, max_steps, mode;
};

constexpr int kBfgsStepThreads = 128;

// deterministic CTA-wide sum: butterfly inside each warp, then the warp totals in a fixed order; every thread gets the total
template <typename T> __device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  __syncthreads();  // `red` may still be read from a previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T t = T(0);
#pragma unroll
  for (int w = 0; w < kBfgsStepThreads / 32; ++w) t += red[w];
  return t;
}
template <typename T> __device__ __forceinline__ T block_nanmax(T v, T* red) {
  v = warp_nanmax(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T t = T(0);
#pragma unroll
  for (int w = 0; w < kBfgsStepThreads / 32; ++w) t = nan_max(t, red[w]);
  return t;
}

template <typename T> __global__ void __launch_bounds__(kBfgsStepThreads) bfgs_step_kernel(const BfgsStepParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ T red[kBfgsStepThreads / 32];
  const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kBfgsStepThreads / 32;
  T* vd = reinterpret_cast<T*>(smem_raw);  // s = y_eval - y
  T* vq = vd + n;                          // q = grad_new - grad
  T* vg = vq + n;                          // grad_new
  T* vhq = vg + n;                         // M v
  T* chol = vhq + n;                       // direct variants: n x (n + 1) copy of the Hessian approximation
  const bool form_b = (p.mode & B200OPT_QN_FORM_B) != 0, direct = (p.mode & B200OPT_QN_DIRECT) != 0;
  const T* vu = (p.mode & B200OPT_QN_SWAP) ? vq : vd;
  const T* vv = (p.mode & B200OPT_QN_SWAP) ? vd : vq;
  const int64_t prob = blockIdx.x;
  T* st = p.state + prob * p.state_stride;
  T* H = st;
  T* gst = st + (int64_t)n * n;
  T* nst = gst + n;
  T* sc = nst + n;  // f, alpha, first, terminate, num_accepted
  int result = p.result[prob];
  int steps = p.num_steps[prob];
  bool terminate = sc[3] != T(0);
  if (terminate || result != RES_OK || steps >= p.max_steps) return;  // block-uniform
  int descent_result = (int)sc[5];
  const T rtol = p.rtol, atol = p.atol;
  const T f = sc[0];
  T alpha = sc[1];
  const bool first = sc[2] != T(0);
  const T fe = p.f_eval[prob];

  // ---- vectors; BacktrackingArmijo.step (backtracking.py:79-99)
  T part_pred = T(0), part_inner = T(0);
  bool fin = true;
  for (int j = tid; j < n; j += kBfgsStepThreads) {
    const T yj = p.y[prob * n + j], yej = p.y_eval[prob * n + j], gj = gst[j], gej = p.g_eval[prob * n + j];
    const T dj = yej - yj, qj = gej - gj;
    vd[j] = dj;
    vq[j] = qj;
    vg[j] = gej;
    part_pred += gj * dj;
    part_inner += qj * dj;
    fin = fin && t_finite(yj);
  }
  const bool y_old_finite = __syncthreads_and(fin ? 1 : 0) != 0;
  const T pred = block_sum(part_pred, red);
  const T df = fe - f;
  const bool accept = first || ((df <= T(0.1) * pred) && (pred <= T(0)));
  alpha = accept ? T(1) : T(0.5) * alpha;
  if (accept) {
    // ---- update_hessian (quasi_newton.py:342-415, inverse branch :371-385)
    const T inner = block_sum(part_inner, red);
    const bool update = inner > Num<T>::eps();  // :360
    T c1 = T(0), c2 = T(0);
    if (update) {
      // pass 1: mvp = M v (one warp per row, coalesced)
      for (int i = warp; i < n; i += NW) {
        T acc = T(0);
        for (int j = lane; j < n; j += 32) acc += H[(int64_t)i * n + j] * vv[j];
        acc = warp_sum(acc);
        if (lane == 0) vhq[i] = acc;
      }
      __syncthreads();
      T part = T(0);
      for (int j = tid; j < n; j += kBfgsStepThreads) part += vv[j] * vhq[j];
      const T mvp_inner = block_sum(part, red);
      if (form_b) {
        c1 = T(1) / inner;      // u u^T / inner
        c2 = T(1) / mvp_inner;  // Mv Mv^T / (v.Mv)
      } else {
        c1 = (inner + mvp_inner) / (inner * inner);  // term1 coefficient (:379)
        c2 = T(1) / inner;                           // term2 coefficient (:380)
      }
    }
    // pass 2: the rank-2 update, fused (inverse variants) with newton = M' grad_new (gauss_newton.py:60-61)
    for (int i = warp; i < n; i += NW) {
      const T u_i = vu[i], mv_i = update ? vhq[i] : T(0);
      T acc = T(0);
      for (int j = lane; j < n; j += 32) {
        T h = H[(int64_t)i * n + j];
        if (update) {
          h = form_b ? h + c1 * (u_i * vu[j]) - c2 * (mv_i * vhq[j]) : h + c1 * (u_i * vu[j]) - c2 * (mv_i * vu[j] + u_i * vhq[j]);
          H[(int64_t)i * n + j] = h;
        }
        if (direct) chol[i * (n + 1) + j] = h;
        acc += h * vg[j];
      }
      acc = warp_sum(acc);
      if (lane == 0 && !direct) nst[i] = acc;
    }
    if (direct) {
      // newton = M'^-1 grad_new by Cholesky (lower factor in place; a non-positive pivot gives NaN -> `singular`)
      __syncthreads();
      for (int k = 0; k < n; ++k) {
        const T dk = t_sqrt(chol[k * (n + 1) + k]);
        __syncthreads();
        for (int i = k + tid; i < n; i += kBfgsStepThreads) chol[i * (n + 1) + k] = (i == k) ? dk : chol[i * (n + 1) + k] / dk;
        __syncthreads();
        const int rem = n - k - 1;
        for (int e = tid; e < rem * rem; e += kBfgsStepThreads) {
          const int i = k + 1 + e / rem, j = k + 1 + e % rem;
          if (j <= i) chol[i * (n + 1) + j] -= chol[i * (n + 1) + k] * chol[j * (n + 1) + k];
        }
        __syncthreads();
      }
      if (warp == 0) {  // substitutions: one warp, lane-strided rows, serial in k
        for (int i = lane; i < n; i += 32) vhq[i] = vg[i];
        __syncwarp();
        for (int k = 0; k < n; ++k) {
          const T zk = vhq[k] / chol[k * (n + 1) + k];
          __syncwarp();
          if (lane == 0) vhq[k] = zk;
          for (int i = k + 1 + lane; i < n; i += 32) vhq[i] -= chol[i * (n + 1) + k] * zk;
          __syncwarp();
        }
        for (int k = n - 1; k >= 0; --k) {
          const T xk = vhq[k] / chol[k * (n + 1) + k];
          __syncwarp();
          if (lane == 0) vhq[k] = xk;
          for (int i = lane; i < k; i += 32) vhq[i] -= chol[k * (n + 1) + i] * xk;
          __syncwarp();
        }
        bool fin_sol = true;
        for (int i = lane; i < n; i += 32) {
          nst[i] = vhq[i];
          fin_sol = fin_sol && t_finite(vhq[i]);
        }
        fin_sol = __all_sync(kFull, fin_sol);
        if (lane == 0) red[0] = fin_sol ? T(0) : T(1);
      }
      __syncthreads();
      descent_result = (red[0] != T(0)) ? RES_SINGULAR : RES_OK;  // newton_step: RESULTS.promote(out.result)
      __syncthreads();
    }
    // ---- termination (quasi_newton.py:244-249): cauchy_termination on y and on the scalar f, never on the first step
    NormAcc<T> yn(p.norm);
    for (int j = tid; j < n; j += kBfgsStepThreads) yn.add(t_abs(vd[j]) / (atol + rtol * t_abs(p.y_eval[prob * n + j])));
    const T ynorm_acc = (p.norm == B200OPT_NORM_MAX) ? block_nanmax(yn.v, red) : block_sum(yn.v, red);
    yn.v = ynorm_acc;
    NormAcc<T> fnrm(p.norm);
    fnrm.add(t_abs(df) / (atol + rtol * t_abs(fe)));
    terminate = (yn.finish(n) < T(1)) && (fnrm.finish(1) < T(1)) && !first;
    __syncthreads();  // nst complete, every read of y_eval above done
    for (int j = tid; j < n; j += kBfgsStepThreads) {
      const T yj = p.y_eval[prob * n + j];  // y <- y_eval
      p.y[prob * n + j] = yj;
      gst[j] = vg[j];
      p.y_eval[prob * n + j] = yj - alpha * nst[j];  // gauss_newton.py:126-127
    }
  } else {
    terminate = false;
    for (int j = tid; j < n; j += kBfgsStepThreads) p.y_eval[prob * n + j] = p.y[prob * n + j] - alpha * nst[j];
  }
  if (p.root_norm) terminate = terminate && (p.root_norm[prob] < atol);  // _ToRoot.terminate (_root_find.py:64-73)
  result = descent_result;  // quasi_newton.py:280-283: the descent's result (the search always succeeds)
  if (!y_old_finite) result = RES_NONFINITE;  // _iterate.py:216
  ++steps;
  if (tid == 0) {
    if (accept) {
      sc[0] = fe;
      sc[4] = sc[4] + T(1);
      sc[5] = (T)descent_result;
    }
    sc[1] = alpha;
    sc[2] = T(0);
    sc[3] = terminate ? T(1) : T(0);
    p.num_steps[prob] = steps;
    p.result[prob] = result;
    if (p.active_count && !terminate && result == RES_OK && steps < p.max_steps) atomicAdd(p.active_count, 1);
  }
}

template <typename T> __global__ void bfgs_step_init_kernel(int64_t batch, int n, int64_t stride, const T* y0, T* y, T* y_eval, T* state,
                                                            int32_t* num_steps, int32_t* result) {
  const int64_t prob = blockIdx.x;
  if (prob >= batch) return;
  T* st = state + prob * stride;
  for (int64_t e = threadIdx.x; e < stride; e += blockDim.x) {
    const int64_t i = e / n, j = e - i * n;
    st[e] = (e < (int64_t)n * n && i == j) ? T(1) : T(0);  // H^-1 = I (quasi_newton.py:41-75), f = 0, g = 0, newton = 0
  }
  __syncthreads();
  bool fin = true;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const T v = y0[prob * n + j];
    y[prob * n + j] = v;
    y_eval[prob * n + j] = v;
    fin = fin && t_finite(v);
  }
  const int allfin = __syncthreads_and(fin ? 1 : 0);
  if (threadIdx.x == 0) {
    T* sc = st + (int64_t)n * n + 2 * n;
    sc[1] = T(1);  // alpha = step_init (backtracking.py:51-53)
    sc[2] = T(1);  // first_step
    num_steps[prob] = 0;
    result[prob] = allfin ? RES_OK : RES_NONFINITE;
  }
}

template <typename T> __global__ void step_finish_kernel(int64_t batch, int64_t stride, int64_t term_off, const T* state, int32_t* result) {
  const int64_t prob = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (prob >= batch) return;
  const bool terminate = state[prob * stride + term_off] != T(0);
  if (result[prob] == RES_OK && !terminate) result[prob] = RES_NL_MAX_STEPS;  // _iterate.py:253-257
}

template <typename T>
static int root_step_typed(int64_t batch, int n, const void* f, const void* J, const void* lower, const void* upper, void* y, void* state,
                           const b200opt_options_t* o, int32_t* num_steps, int32_t* result, int32_t* active, cudaStream_t st) {
  RootStepParams<T> p;
  p.f_eval = static_cast<const T*>(f);
  p.J_eval = static_cast<const T*>(J);
  p.lower = static_cast<const T*>(lower);
  p.upper = static_cast<const T*>(upper);
  p.y = static_cast<T*>(y);
  p.state = static_cast<T*>(state);
  p.num_steps = num_steps;
  p.result = result;
  p.active_count = active;
  p.batch = batch;
  p.n = n;
  p.rtol = (T)o->rtol;
  p.atol = (T)o->atol;
  p.kappa = (T)o->kappa;
  p.norm = o->norm;
  p.max_steps = o->max_steps;
  p.cauchy = o->cauchy_termination;
  if (active && check_cuda(cudaMemsetAsync(active, 0, sizeof(int32_t), st), "memset active_count")) return B200OPT_ECUDA;
  if (n > 32) {  // one CTA per problem, Jacobian in shared memory
    const size_t smem = big_root_smem<T>(n);
    static LaunchCache cache;
    if (cache.configure(root_step_big_kernel<T>, kBigRootThreads, smem) < 0) return B200OPT_ECUDA;
    root_step_big_kernel<T><<<(unsigned)batch, kBigRootThreads, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "root_step_big_kernel launch");
  }
  const unsigned grid = (unsigned)((batch + kRootStepWarps - 1) / kRootStepWarps);
  if (n <= 8) root_step_kernel<T, 8><<<grid, kRootStepWarps * 32, 0, st>>>(p);
  else if (n <= 16) root_step_kernel<T, 16><<<grid, kRootStepWarps * 32, 0, st>>>(p);
  else root_step_kernel<T, 32><<<grid, kRootStepWarps * 32, 0, st>>>(p);
  return check_cuda(cudaGetLastError(), "root_step_kernel launch");
}

template <typename T>
static int bfgs_step_typed(int mode, int64_t batch, int n, const void* f, const void* g, const void* root_norm, void* y, void* y_eval,
                           void* state, const b200opt_options_t* o, int32_t* num_steps, int32_t* result, int32_t* active, cudaStream_t st) {
  BfgsStepParams<T> p;
  p.f_eval = static_cast<const T*>(f);
  p.g_eval = static_cast<const T*>(g);
  p.root_norm = static_cast<const T*>(root_norm);
  p.y = static_cast<T*>(y);
  p.y_eval = static_cast<T*>(y_eval);
  p.state = static_cast<T*>(state);
  p.state_stride = bfgs_state_stride(n);
  p.num_steps = num_steps;
  p.result = result;
  p.active_count = active;
  p.batch = batch;
  p.n = n;
  p.rtol = (T)o->rtol;
  p.atol = (T)o->atol;
  p.norm = o->norm;
  p.max_steps = o->max_steps;
  p.mode = mode;
  if (active && check_cuda(cudaMemsetAsync(active, 0, sizeof(int32_t), st), "memset active_count")) return B200OPT_ECUDA;
  const size_t smem = (4 * (size_t)n + ((mode & B200OPT_QN_DIRECT) ? (size_t)n * (n + 1) : 0)) * sizeof(T);
  bfgs_step_kernel<T><<<(unsigned)batch, kBfgsStepThreads, smem, st>>>(p);
  return check_cuda(cudaGetLastError(), "bfgs_step_kernel launch");
}

}  // namespace b200opt

using namespace b200opt;

static int gstep_check(const char* who, int dtype, int64_t batch, int n, int n_max) {
  if (dtype != B200OPT_F32 && dtype != B200OPT_F64) {
    set_error("%s: unknown dtype %d", who, dtype);
    return B200OPT_EINVAL;
  }
  if (batch < 0 || batch > 2147483647LL || n < 1 || n > n_max) {
    set_error("%s: need 0 <= batch < 2^31 and 1 <= n <= %d (n=%d)", who, n_max, n);
    return B200OPT_EINVAL;
  }
  return 0;
}
static int opts_check(const char* who, const b200opt_options_t* o) {
  if (!o || o->norm < 0 || o->norm > 2 || o->max_steps < 0) {
    set_error("%s: bad options", who);
    return B200OPT_EINVAL;
  }
  return 0;
}

constexpr int kBfgsStepNMax = 1024;  // 4 n elements of shared memory per CTA, n^2 elements of state per problem

extern "C" {

int64_t b200opt_root_step_state_size(int32_t dtype, int32_t n) {
  if (gstep_check("root_step_state_size", dtype, 0, n, dtype == B200OPT_F64 ? 160 : 224)) return -1;
  return (int64_t)kRootStateElems * (dtype == B200OPT_F32 ? 4 : 8);
}

int32_t b200opt_root_step_init(int32_t dtype, int64_t batch, int32_t n, const void* y0, void* y, void* state, int32_t* num_steps,
                               int32_t* result, void* stream) {
  if (int e = gstep_check("root_step_init", dtype, batch, n, dtype == B200OPT_F64 ? 160 : 224)) return e;
  if (batch == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200OPT_F64)
    root_step_init_kernel<double><<<(unsigned)batch, 32, 0, st>>>(batch, n, static_cast<const double*>(y0), static_cast<double*>(y),
                                                                  static_cast<double*>(state), num_steps, result);
  else
    root_step_init_kernel<float><<<(unsigned)batch, 32, 0, st>>>(batch, n, static_cast<const float*>(y0), static_cast<float*>(y),
                                                                 static_cast<float*>(state), num_steps, result);
  return check_cuda(cudaGetLastError(), "root_step_init launch");
}

int32_t b200opt_root_step(int32_t dtype, int64_t batch, int32_t n, const void* f_eval, const void* J_eval, const void* lower,
                          const void* upper, void* y, void* state, const b200opt_options_t* opts, int32_t* num_steps,
                          int32_t* result, int32_t* active_count, void* stream) {
  if (int e = gstep_check("root_step", dtype, batch, n, dtype == B200OPT_F64 ? 160 : 224)) return e;
  if (int e = opts_check("root_step", opts)) return e;
  if (batch == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200OPT_F64)
    return root_step_typed<double>(batch, n, f_eval, J_eval, lower, upper, y, state, opts, num_steps, result, active_count, st);
  return root_step_typed<float>(batch, n, f_eval, J_eval, lower, upper, y, state, opts, num_steps, result, active_count, st);
}

int32_t b200opt_root_step_finish(int32_t dtype, int64_t batch, int32_t n, const void* state, int32_t* result, void* stream) {
  if (int e = gstep_check("root_step_finish", dtype, batch, n, dtype == B200OPT_F64 ? 160 : 224)) return e;
  if (batch == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((batch + 127) / 128);
  if (dtype == B200OPT_F64)
    step_finish_kernel<double><<<grid, 128, 0, st>>>(batch, kRootStateElems, 2, static_cast<const double*>(state), result);
  else
    step_finish_kernel<float><<<grid, 128, 0, st>>>(batch, kRootStateElems, 2, static_cast<const float*>(state), result);
  return check_cuda(cudaGetLastError(), "root_step_finish launch");
}

int32_t b200opt_lu_solve(int32_t dtype, int64_t batch, int32_t n, const void* A, const void* b, void* x, int32_t* status, void* stream) {
  if (int e = gstep_check("lu_solve", dtype, batch, n, dtype == B200OPT_F64 ? 160 : 224)) return e;
  if (batch == 0) return 0;
  if (!A || !b || !x) {
    set_error("lu_solve: NULL buffer");
    return B200OPT_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n > 32) {  // one CTA per system, matrix in shared memory
    if (dtype == B200OPT_F64) {
      static LaunchCache cache;
      const size_t smem = big_root_smem<double>(n);
      if (cache.configure(lu_solve_big_kernel<double>, kBigRootThreads, smem) < 0) return B200OPT_ECUDA;
      lu_solve_big_kernel<double><<<(unsigned)batch, kBigRootThreads, smem, st>>>(static_cast<const double*>(A), static_cast<const double*>(b),
                                                                                  static_cast<double*>(x), status, n);
    } else {
      static LaunchCache cache;
      const size_t smem = big_root_smem<float>(n);
      if (cache.configure(lu_solve_big_kernel<float>, kBigRootThreads, smem) < 0) return B200OPT_ECUDA;
      lu_solve_big_kernel<float><<<(unsigned)batch, kBigRootThreads, smem, st>>>(static_cast<const float*>(A), static_cast<const float*>(b),
                                                                                 static_cast<float*>(x), status, n);
    }
    return check_cuda(cudaGetLastError(), "lu_solve_big_kernel launch");
  }
  const unsigned grid = (unsigned)((batch + kRootStepWarps - 1) / kRootStepWarps);
  const int threads = kRootStepWarps * 32;
#define B200OPT_LU_LAUNCH(T, N) lu_solve_kernel<T, N><<<grid, threads, 0, st>>>(static_cast<const T*>(A), static_cast<const T*>(b), static_cast<T*>(x), status, batch, n)
  if (dtype == B200OPT_F64) {
    if (n <= 8) B200OPT_LU_LAUNCH(double, 8);
    else if (n <= 16) B200OPT_LU_LAUNCH(double, 16);
    else B200OPT_LU_LAUNCH(double, 32);
  } else {
    if (n <= 8) B200OPT_LU_LAUNCH(float, 8);
    else if (n <= 16) B200OPT_LU_LAUNCH(float, 16);
    else B200OPT_LU_LAUNCH(float, 32);
  }
#undef B200OPT_LU_LAUNCH
  return check_cuda(cudaGetLastError(), "lu_solve_kernel launch");
}

int64_t b200opt_bfgs_step_state_size(int32_t dtype, int32_t n) {
  if (gstep_check("bfgs_step_state_size", dtype, 0, n, kBfgsStepNMax)) return -1;
  return bfgs_state_stride(n) * (dtype == B200OPT_F32 ? 4 : 8);
}

int32_t b200opt_bfgs_step_init(int32_t dtype, int64_t batch, int32_t n, const void* y0, void* y, void* y_eval, void* state,
                               int32_t* num_steps, int32_t* result, void* stream) {
  if (int e = gstep_check("bfgs_step_init", dtype, batch, n, kBfgsStepNMax)) return e;
  if (batch == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t stride = bfgs_state_stride(n);
  if (dtype == B200OPT_F64)
    bfgs_step_init_kernel<double><<<(unsigned)batch, 128, 0, st>>>(batch, n, stride, static_cast<const double*>(y0), static_cast<double*>(y),
                                                                   static_cast<double*>(y_eval), static_cast<double*>(state), num_steps, result);
  else
    bfgs_step_init_kernel<float><<<(unsigned)batch, 128, 0, st>>>(batch, n, stride, static_cast<const float*>(y0), static_cast<float*>(y),
                                                                  static_cast<float*>(y_eval), static_cast<float*>(state), num_steps, result);
  return check_cuda(cudaGetLastError(), "bfgs_step_init launch");
}

int32_t b200opt_bfgs_step(int32_t update, int32_t use_inverse, int32_t dtype, int64_t batch, int32_t n, const void* f_eval,
                          const void* g_eval, const void* root_norm, void* y, void* y_eval, void* state,
                          const b200opt_options_t* opts, int32_t* num_steps, int32_t* result, int32_t* active_count, void* stream) {
  if (int e = gstep_check("bfgs_step", dtype, batch, n, use_inverse ? kBfgsStepNMax : kBfgsDirectNMax)) return e;
  if (int e = opts_check("bfgs_step", opts)) return e;
  if (update != B200OPT_UPDATE_BFGS && update != B200OPT_UPDATE_DFP) {
    set_error("bfgs_step: unknown quasi-Newton update %d", update);
    return B200OPT_EINVAL;
  }
  if (batch == 0) return 0;
  // BFGS inverse: form A (u=s, v=q); DFP inverse: form B (u=s, v=q); BFGS direct: form B swapped; DFP direct: form A swapped
  int mode = 0;
  if ((update == B200OPT_UPDATE_DFP) == (use_inverse != 0)) mode |= B200OPT_QN_FORM_B;
  if (!use_inverse) mode |= B200OPT_QN_SWAP | B200OPT_QN_DIRECT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200OPT_F64)
    return bfgs_step_typed<double>(mode, batch, n, f_eval, g_eval, root_norm, y, y_eval, state, opts, num_steps, result, active_count, st);
  return bfgs_step_typed<float>(mode, batch, n, f_eval, g_eval, root_norm, y, y_eval, state, opts, num_steps, result, active_count, st);
}

int32_t b200opt_bfgs_step_finish(int32_t dtype, int64_t batch, int32_t n, const void* state, int32_t* result, void* stream) {
  if (int e = gstep_check("bfgs_step_finish", dtype, batch, n, kBfgsStepNMax)) return e;
  if (batch == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((batch + 127) / 128);
  const int64_t stride = bfgs_state_stride(n), term_off = (int64_t)n * n + 2 * n + 3;
  if (dtype == B200OPT_F64)
    step_finish_kernel<double><<<grid, 128, 0, st>>>(batch, stride, term_off, static_cast<const double*>(state), result);
  else
    step_finish_kernel<float><<<grid, 128, 0, st>>>(batch, stride, term_off, static_cast<const float*>(state), result);
  return check_cuda(cudaGetLastError(), "bfgs_step_finish launch");
}

}  // extern "C"
