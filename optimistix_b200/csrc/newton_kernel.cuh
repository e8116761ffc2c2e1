// Batched Newton / Chord root finding: one warp per problem, lane i owns row i of the Jacobian.
//
// Restates, per problem:
//   driver      optimistix/_iterate.py:194-268
//   init        optimistix/_solver/newton_chord.py:69-101
//   step        optimistix/_solver/newton_chord.py:103-152   (diff = J^-1 f, y <- y - diff, optional clip)
//   terminate   optimistix/_solver/newton_chord.py:154-196   (Cauchy test against f=0, or Hairer-Wanner)
//   linear solve: lx.AutoLinearSolver(well_posed=None) on a square operator = LU with partial pivoting
//               (lineax, not vendored); a non-finite solution is reported as `singular`.
//
// B200 design: the n x n Jacobian never exists in memory.  Lane i evaluates row i into registers, the
// warp factorises it in place (row-per-lane LU, partial pivoting by a 2-instruction REDUX arg-max, no
// physical row swaps: the pivot ORDER is tracked instead), the pivot row of each elimination step is
// broadcast through a 264-byte shared-memory line with 128-bit accesses (half the LSU cost of 64-bit
// shuffles), and forward/back substitution run on shuffles.  Chord keeps the factors of J(y0) in the
// same registers for the whole solve.  HBM traffic per problem: y_prev, dt in; y, num_steps, result out.
#pragma once
#include "common.cuh"
#include "models.cuh"

// Tuning switches (A/B-measured on B200, 256 Ki implicit-Euler root finds; defaults are the fastest):
//   ROOT_ARGMAX_BRANCH  second REDUX only when several candidates share the high word     6.65 ms vs 6.55 ms (off)
//   ROOT_EARLY_RCP      start every candidate's reciprocal before the pivot is known       6.73 ms vs 6.55 ms (off)
#ifndef ROOT_ARGMAX_BRANCH
#define ROOT_ARGMAX_BRANCH 0
#endif
#ifndef ROOT_EARLY_RCP
#define ROOT_EARLY_RCP 0
#endif

namespace b200opt {

template <typename T> struct RootParams {
  const T* y0;
  const T* shared;
  const T* prob;
  int64_t prob_stride;
  const T* lower;  // optional [n]
  const T* upper;  // optional [n]
  T* y_out;
  int32_t* num_steps;
  int32_t* result;
  int64_t batch;
  int n;
  T rtol, atol, kappa;
  int norm;
  int max_steps;
  int cauchy;
  int chord;
};

constexpr int kRootWarpsPerBlock = 4;

// ---- arg-max of |v| over the lanes in `valid`, first lane wins ties (LAPACK i?amax) -------------------------------
__device__ __forceinline__ int warp_argmax_abs(double v, bool valid) {
  const double a = fabs(v);
  const unsigned hi = valid ? (unsigned)__double2hiint(a) + 1u : 0u;  // non-negative doubles order like integers
  const unsigned m1 = __reduce_max_sync(kFull, hi);
  const bool c1 = valid && hi == m1;
#if ROOT_ARGMAX_BRANCH
  unsigned b = __ballot_sync(kFull, c1);
  if (b & (b - 1)) {  // several candidates share the high word (rare): the low words decide (warp-uniform branch)
    const unsigned lo = c1 ? (unsigned)__double2loint(a) : 0u;
    const unsigned m2 = __reduce_max_sync(kFull, lo);
    b = __ballot_sync(kFull, c1 && lo == m2);
  }
  return __ffs(b) - 1;
#else
  const unsigned lo = c1 ? (unsigned)__double2loint(a) : 0u;
  const unsigned m2 = __reduce_max_sync(kFull, lo);
  const unsigned b = __ballot_sync(kFull, c1 && lo == m2);
  return __ffs(b) - 1;
#endif
}
__device__ __forceinline__ int warp_argmax_abs(float v, bool valid) {
  const unsigned key = valid ? __float_as_uint(fabsf(v)) + 1u : 0u;
  const unsigned m1 = __reduce_max_sync(kFull, key);
  const unsigned b = __ballot_sync(kFull, valid && key == m1);
  return __ffs(b) - 1;
}

// In-place LU of the matrix whose row `lane` is a[0..N).  On exit:
//   mycol      = the elimination step at which this lane's row was the pivot row (its U row starts there)
//   a[k]       = multiplier l_ik for k < mycol, U entries for k >= mycol;  rinv = 1 / U[mycol][mycol]
//   piv[k]     (shared, per warp) = lane that was the pivot of step k
// `rowbuf` is a per-warp shared line of N + 4 elements, 16-byte aligned (slot N carries 1 / pivot).
#ifndef ROOT_FAST_RCP
#define ROOT_FAST_RCP 1
#endif
#ifndef ROOT_STS64
#define ROOT_STS64 1
#endif
__device__ __forceinline__ float t_rcp(float x) { return __frcp_rn(x); }
#if ROOT_FAST_RCP
// 1 / pivot: hardware seed (rcp.approx.ftz.f64, relative error 2^-23) + two Newton steps -> ~1 ulp in 5 instructions
// (the correctly rounded __drcp_rn costs 13 and sits on the critical path of every elimination step).  A zero pivot
// gives NaN instead of inf; both end as a non-finite solution = `singular`.
__device__ __forceinline__ double t_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
#else
__device__ __forceinline__ double t_rcp(double x) { return __drcp_rn(x); }
#endif

#ifndef ROOT_LOOKAHEAD
#define ROOT_LOOKAHEAD 1
#endif
#if ROOT_LOOKAHEAD
// Look-ahead form: step k first brings column k + 1 up to date (one FMA per lane), starts the arg-max of step k + 1 on it
// (two dependent REDUX + a ballot, ~70 cycles of latency) and only then applies the rest of step k's row update, which no
// longer sits behind that latency.  The update is predicated, not branched, so that the whole step is one basic block the
// assembler can schedule.  Same arithmetic, same pivots, same results as the plain form below.
template <typename T, int N>
__device__ __forceinline__ void warp_lu_factor(T (&a)[N], int& mycol, T& rinv, T* rowbuf, unsigned char* piv, int lane) {
  using Vec = typename Num<T>::Vec;
  constexpr int V = Num<T>::kVec;
  static_assert(N % V == 0, "N must be a multiple of the 128-bit vector width");
  bool done = lane >= N;
  mycol = N;
  rinv = T(0);
  int p = warp_argmax_abs(a[0], !done);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int k0 = (k / V) * V;
    if (lane == p) {
      if (sizeof(T) == 8 && ROOT_STS64) {
#pragma unroll
        for (int j = k; j < N; ++j) rowbuf[j] = a[j];
      } else {
#pragma unroll
        for (int j = k0; j < N; j += V) {
          Vec v;
          T* vv = reinterpret_cast<T*>(&v);
#pragma unroll
          for (int q = 0; q < V; ++q) vv[q] = a[j + q];
          *reinterpret_cast<Vec*>(rowbuf + j) = v;
        }
      }
      piv[k] = (unsigned char)p;
    }
    __syncwarp();
    T row[N];  // the part of the pivot row still in use (entries below k0 are never read)
#pragma unroll
    for (int j = k0; j < N; j += V) {
      const Vec v = *reinterpret_cast<const Vec*>(rowbuf + j);
      const T* vv = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int q = 0; q < V; ++q) row[j + q] = vv[q];
    }
    const T r = t_rcp(row[k]);
    const bool elim = !done && lane != p;
    if (lane == p) {
      done = true;
      mycol = k;
      rinv = r;
    }
    const T l = a[k] * r;
    if (elim) a[k] = l;
    int pn = 0;
    if (k + 1 < N) {
      if (elim) a[k + 1] -= l * row[k + 1];
      pn = warp_argmax_abs(a[k + 1], !done);
    }
#pragma unroll
    for (int j = k + 2; j < N; ++j)
      if (elim) a[j] -= l * row[j];
    p = pn;
    __syncwarp();
  }
}
#else
template <typename T, int N>
__device__ __forceinline__ void warp_lu_factor(T (&a)[N], int& mycol, T& rinv, T* rowbuf, unsigned char* piv, int lane) {
  using Vec = typename Num<T>::Vec;
  constexpr int V = Num<T>::kVec;
  static_assert(N % V == 0, "N must be a multiple of the 128-bit vector width");
  bool done = lane >= N;
  mycol = N;
  rinv = T(0);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int k0 = (k / V) * V;  // 128-bit aligned start of the part of the pivot row still in use
    // every candidate's reciprocal is started before the pivot is known (getf2 scales the column by 1/pivot): the
    // 5-instruction dependent chain overlaps the arg-max reduction instead of following the pivot-row broadcast
#if ROOT_EARLY_RCP
    const T rc = t_rcp(a[k]);
#endif
    const int p = warp_argmax_abs(a[k], !done);
    if (lane == p) {
#if ROOT_EARLY_RCP
      rowbuf[N] = rc;
#endif
#if ROOT_STS64
      // element-wise stores: every double already sits in an aligned register pair, a 128-bit store would first have
      // to move two of them into an aligned quad
      if (sizeof(T) == 8) {
#pragma unroll
        for (int j = k; j < N; ++j) rowbuf[j] = a[j];
      } else
#endif
      {
#pragma unroll
        for (int j = k0; j < N; j += V) {
          Vec v;
          T* vv = reinterpret_cast<T*>(&v);
#pragma unroll
          for (int q = 0; q < V; ++q) vv[q] = a[j + q];
          *reinterpret_cast<Vec*>(rowbuf + j) = v;
        }
      }
      piv[k] = (unsigned char)p;
    }
    __syncwarp();
    // pivot row: broadcast reads (every lane reads the same 16 bytes: one wavefront per 128-bit load)
#if ROOT_EARLY_RCP
    const T r = rowbuf[N];
#else
    const T r = t_rcp(rowbuf[k]);
#endif
    const bool elim = !done && lane != p;
    if (lane == p) {
      done = true;
      mycol = k;
      rinv = r;
    }
    if (elim) {  // a real (divergent) branch: rows already used as pivots idle instead of computing-and-discarding
      const T l = a[k] * r;
      a[k] = l;
#pragma unroll
      for (int j = k0; j < N; j += V) {
        const Vec v = *reinterpret_cast<const Vec*>(rowbuf + j);
        const T* vv = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int q = 0; q < V; ++q)
          if (j + q > k) a[j + q] -= l * vv[q];  // compile-time condition after unrolling
      }
    }
    __syncwarp();
  }
}

#endif  // ROOT_LOOKAHEAD

// Solve with the factors above.  In: rhs (this lane's row entry).  Out: x_lane (solution component `lane`).
template <typename T, int N>
__device__ __forceinline__ T warp_lu_solve(const T (&a)[N], int mycol, T rinv, T rhs, const unsigned char* piv, int lane) {
  // forward: step k subtracts l_ik * rhs_pivot(k) from every row eliminated after k
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const T bk = __shfl_sync(kFull, rhs, piv[k]);
    if (mycol > k) rhs -= a[k] * bk;
  }
  // backward: x_k = (rhs_pk - sum_{j>k} U_kj x_j) / U_kk, broadcast as soon as it is known
  T x = T(0);
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    const T xk = __shfl_sync(kFull, rhs * rinv, piv[k]);
    if (mycol < k) rhs -= a[k] * xk;
    if (lane == k) x = xk;
  }
  return x;
}

// Newton (a fresh Jacobian every step, factors never reused): Gauss-Jordan elimination with partial pivoting on the
// augmented rows [J_i | f_i].  In the row-per-lane layout the elimination of the rows ABOVE the pivot costs no extra
// instruction (the same predicated FMAs simply run on more lanes), and it removes both substitution sweeps - 64
// dependent shuffle-broadcast steps, ~650 of the ~4100 instructions of a Newton iteration.  Pivots, multipliers and the
// rows of U are those of the LU above bit for bit; only the order of the operations that would be the back-substitution
// differs (rounding-level differences in the solution).  Chord keeps the LU: it reuses the factors.
// The arg-max of step k + 1 is started as soon as column k + 1 is up to date (look-ahead, see warp_lu_factor).
#ifndef ROOT_GJ
#define ROOT_GJ 1
#endif
template <typename T, int N>
__device__ __forceinline__ T warp_gj_solve(T (&a)[N], T rhs, T* rowbuf, int lane) {
  using Vec = typename Num<T>::Vec;
  constexpr int V = Num<T>::kVec;
  static_assert(N % V == 0, "N must be a multiple of the 128-bit vector width");
  const bool act = lane < N;
  bool done = !act;
  int mycol = 0;
  T rinv = T(0);
  int p = warp_argmax_abs(a[0], !done);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int k0 = (k / V) * V;
    if (lane == p) {
#pragma unroll
      for (int j = k0; j < N; j += V) {
        Vec v;
        T* vv = reinterpret_cast<T*>(&v);
#pragma unroll
        for (int q = 0; q < V; ++q) vv[q] = a[j + q];
        *reinterpret_cast<Vec*>(rowbuf + j) = v;
      }
      rowbuf[N] = rhs;
    }
    __syncwarp();
    T row[N];
#pragma unroll
    for (int j = k0; j < N; j += V) {
      const Vec v = *reinterpret_cast<const Vec*>(rowbuf + j);
      const T* vv = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int q = 0; q < V; ++q) row[j + q] = vv[q];
    }
    const T rb = rowbuf[N];
    const T r = t_rcp(row[k]);
    const bool other = act && lane != p;  // every other row, already used as a pivot or not
    if (lane == p) {
      done = true;
      mycol = k;
      rinv = r;
    }
    // the multiplier is zeroed for the pivot row itself (and for lanes that own no row) so that the updates below are plain
    // FMAs.  0 * inf = NaN can then reach a row that a branch would have left alone - only when the pivot row holds a
    // non-finite entry, in which case the solution is non-finite (`singular`) either way.
    const T l = other ? a[k] * r : T(0);
    int pn = 0;
    if (k + 1 < N) {
      a[k + 1] -= l * row[k + 1];
      pn = warp_argmax_abs(a[k + 1], !done);
    }
#pragma unroll
    for (int j = k + 2; j < N; ++j) a[j] -= l * row[j];
    rhs -= l * rb;
    p = pn;
    __syncwarp();
  }
  // row `lane` is now e_mycol * U[mycol][mycol]: unknown number mycol sits in this lane; route it to lane mycol
  if (act) rowbuf[mycol] = rhs * rinv;
  __syncwarp();
  const T x = act ? rowbuf[lane] : T(0);
  __syncwarp();
  return x;
}

// ---- model family id 2: implicit Euler step of y' = A y + y*(B y) ----------------------------------------------
//   F(y) = y - y_prev - dt (A y + y*(B y));   J = I - dt (A + diag(B y) + diag(y) B)
//   shared_args = A[n][n] then B[n][n] (row-major); problem_args = y_prev[n], dt
template <typename T> struct ImplicitEuler {
  static constexpr int N = 32;
  static constexpr int kSharedElems = 2 * N * (N + 1);  // A and B staged in shared memory with pitch N+1
  struct Ctx {
    T y_prev, dt;
  };
  __device__ static __forceinline__ void stage_shared(const T* g, T* s, int tid, int nthreads) {
    for (int e = tid; e < 2 * N * N; e += nthreads) {
      const int mat = e / (N * N), i = (e / N) % N, j = e % N;
      s[mat * N * (N + 1) + i * (N + 1) + j] = g[e];
    }
  }
  __device__ static __forceinline__ Ctx load(const T* prob, int lane) { return Ctx{prob[lane], prob[N]}; }
  // row `lane` of F and J at y (ys = the whole iterate in shared memory)
  __device__ static __forceinline__ void eval_row(const T* sh, const T* ys, const Ctx& c, int lane, T& f, T (&row)[N], bool want_jac) {
    const T* Ar = sh + lane * (N + 1);
    const T* Br = sh + N * (N + 1) + lane * (N + 1);
    const T yi = ys[lane];
    T Ay = T(0), By = T(0);
    if (want_jac) {
      // one pass over the row of A and B: matvec partials and the y-dependent part of the Jacobian row
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const T yj = ys[j], aj = Ar[j], bj = Br[j];
        Ay += aj * yj;
        By += bj * yj;
        row[j] = aj + yi * bj;
      }
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const bool diag = (j == lane);
        row[j] = (diag ? T(1) : T(0)) - c.dt * (diag ? row[j] + By : row[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const T yj = ys[j];
        Ay += Ar[j] * yj;
        By += Br[j] * yj;
      }
    }
    f = yi - c.y_prev - c.dt * (Ay + yi * By);
  }
};

template <class Model, typename T> __host__ __device__ constexpr size_t root_smem_bytes() {
  // shared model data + per warp: y[N] + rowbuf[N + 4] (16B aligned) + piv[N]
  return sizeof(T) * Model::kSharedElems + kRootWarpsPerBlock * (sizeof(T) * (2 * Model::N + 4) + 32);
}

#ifndef ROOT_MINBLOCKS
#define ROOT_MINBLOCKS 4
#endif
template <class Model, typename T>
__global__ void __launch_bounds__(kRootWarpsPerBlock * 32, ROOT_MINBLOCKS) root_solve_kernel(const RootParams<T> p) {
  constexpr int N = Model::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sh = reinterpret_cast<T*>(smem_raw);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + sizeof(T) * Model::kSharedElems + wib * (sizeof(T) * (2 * N + 4) + 32);
  T* ys = reinterpret_cast<T*>(wbase);
  T* rowbuf = ys + N;
  unsigned char* piv = reinterpret_cast<unsigned char*>(rowbuf + N + 4);
  Model::stage_shared(p.shared, sh, threadIdx.x, blockDim.x);
  __syncthreads();

  const int64_t warp0 = (int64_t)blockIdx.x * kRootWarpsPerBlock + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kRootWarpsPerBlock;
  const T rtol = p.rtol, atol = p.atol;
  const bool act = lane < N;
  const T lo = (p.lower && act) ? p.lower[lane] : -Num<T>::inf();
  const T hi = (p.upper && act) ? p.upper[lane] : Num<T>::inf();
  const bool clip = p.lower != nullptr || p.upper != nullptr;
  T small_tol = T(1);  // 10^(2 - precision), newton_chord.py:25-29
  for (int i = 0; i < Num<T>::precision - 2; ++i) small_tol *= T(0.1);

  for (int64_t prob = warp0; prob < p.batch; prob += nwarps) {
    const typename Model::Ctx ctx = Model::load(p.prob + prob * p.prob_stride, act ? lane : 0);
    T y = act ? p.y0[prob * N + lane] : T(0);
    // ---- init (newton_chord.py:88-101)
    T a[N];
    int mycol = N;
    T rinv = T(0);
    T fval = Num<T>::inf(), diff = Num<T>::inf();
    T diffsize = Num<T>::inf(), diffsize_prev = T(1);
    int lin_result = RES_OK, nstep = 0, steps = 0;
    bool terminate = false;  // init terminate is false in both modes (diff = inf / step = 0)
    int result = __all_sync(kFull, !act || t_finite(y)) ? RES_OK : RES_NONFINITE;
    if (p.chord && result == RES_OK) {  // :80-87 Jacobian linearised and factorised once, at y0
      if (act) ys[lane] = y;
      __syncwarp();
      T f = T(0);
      if (act) Model::eval_row(sh, ys, ctx, lane, f, a, true);  // lanes >= N own no row (families with N < 32)
      __syncwarp();
      warp_lu_factor<T, N>(a, mycol, rinv, rowbuf, piv, lane);
    }

    while (!terminate && result == RES_OK && steps < p.max_steps) {
      const bool y_old_finite = __all_sync(kFull, !act || t_finite(y));
      // ---- step (:103-152)
      if (act) ys[lane] = y;
      __syncwarp();
      T fx = T(0);
      if (act) Model::eval_row(sh, ys, ctx, lane, fx, a, !p.chord);
      __syncwarp();
      T sol;
      if (ROOT_GJ && !p.chord) {
        sol = warp_gj_solve<T, N>(a, fx, rowbuf, lane);
      } else {
        if (!p.chord) warp_lu_factor<T, N>(a, mycol, rinv, rowbuf, piv, lane);
        sol = warp_lu_solve<T, N>(a, mycol, rinv, fx, piv, lane);
      }
      lin_result = __all_sync(kFull, !act || t_finite(sol)) ? RES_OK : RES_SINGULAR;
      T new_y = y - sol;
      T d = sol;
      if (clip) {
        new_y = new_y < lo ? lo : new_y;  // jnp.clip(min=), (max=)  :130-133
        new_y = new_y > hi ? hi : new_y;
        d = y - new_y;  // :134-135
      }
      NormAcc<T> ds(p.norm);
      if (act) ds.add(t_abs(d / (atol + rtol * t_abs(new_y))));  // :136-138
      ds.warp_combine();
      diffsize_prev = diffsize;
      diffsize = ds.finish(N);
      fval = fx;
      diff = d;
      ++nstep;
      // ---- terminate, evaluated with the PRE-step y (:154-196, _iterate.py:238)
      int term_result = RES_OK;
      if (p.cauchy) {
        NormAcc<T> yn(p.norm), fn(p.norm);
        if (act) {
          yn.add(t_abs(diff) / (atol + rtol * t_abs(y)));
          fn.add(t_abs(fval) / atol);  // f_scale = atol + rtol*|0|
        }
        yn.warp_combine();
        fn.warp_combine();
        terminate = (yn.finish(N) < T(1)) && (fn.finish(N) < T(1));
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
      y = new_y;
      ++steps;
    }
    if (result == RES_OK && !terminate) result = RES_NL_MAX_STEPS;
    if (act) p.y_out[prob * N + lane] = y;
    if (lane == 0) {
      p.num_steps[prob] = steps;
      p.result[prob] = result;
    }
  }
}

}  // namespace b200opt
