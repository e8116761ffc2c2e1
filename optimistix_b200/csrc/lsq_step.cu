// Generic per-iteration LM / GN step for an ARBITRARY user function: the caller (XLA, torch, ...) evaluates the
// residual r and the Jacobian J at the trial point, this kernel does everything else of one
// AbstractGaussNewton.step (optimistix/_solver/gauss_newton.py:243-334) plus the driver bookkeeping of
// _iterate.py:234-241 for every problem that is still running:
//   search (trust_region.py:77-106 / learning_rate.py:24-35), accept/reject select, Gram J^T J and J^T r of the
//   accepted point, Cauchy termination (_misc.py:291-310), damped Cholesky solve (levenberg_marquardt.py:34-76 in
//   normal-equation form), next trial point, step counter, result codes.
// One warp per problem; the per-problem carry lives in a caller-owned opaque state buffer between calls:
//   [ G packed lower (NP(NP+1)/2) | g (NP) | rr | delta | first | terminate | num_accepted | r_accepted (m) ]
// with NP = n padded to 4 / 8 / 16 (padded unknowns get a unit diagonal and stay exactly zero).
#include "lsq_kernel.cuh"

namespace b200opt {
int device_sm_count();

template <int NP> __host__ __device__ constexpr int step_state_head() { return Tri<NP>::size + NP + 5; }
__host__ __device__ inline int64_t step_state_stride(int np_head, int m) { return ((int64_t)np_head + m + 3) / 4 * 4; }

template <typename T> struct StepParams {
  const T* r_eval;
  const T* J_eval;
  T* y;
  T* y_eval;
  T* state;
  int64_t state_stride;
  int32_t* num_steps;
  int32_t* result;
  int32_t* active_count;
  int64_t batch;
  int n, m;
  T rtol, atol;
  int norm, max_steps, root_near_zero, trust;
};

constexpr int kStepWarps = 4;

template <typename T, int NP> __global__ void __launch_bounds__(kStepWarps * 32) lsq_step_kernel(const StepParams<T> p) {
  constexpr int NG = Tri<NP>::size;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  T* scratch = reinterpret_cast<T*>(smem_raw) + (size_t)wib * WarpReduceLayout<T, NG + NP>::elems;
  const int64_t prob = (int64_t)blockIdx.x * kStepWarps + wib;
  if (prob >= p.batch) return;
  const int n = p.n, m = p.m;
  T* st = p.state + prob * p.state_stride;
  T* st_r = st + step_state_head<NP>();
  int result = p.result[prob];
  int steps = p.num_steps[prob];
  bool terminate = st[NG + NP + 3] != T(0);
  if (terminate || result != RES_OK || steps >= p.max_steps) return;  // this lane of the vmapped loop is frozen

  T y[NP], ye[NP], G[NG], g[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    y[j] = j < n ? p.y[prob * n + j] : T(0);
    ye[j] = j < n ? p.y_eval[prob * n + j] : T(0);
    g[j] = st[NG + j];
  }
#pragma unroll
  for (int a = 0; a < NG; ++a) G[a] = st[a];
  T rr = st[NG + NP + 0], delta = st[NG + NP + 1];
  const bool first = st[NG + NP + 2] != T(0);
  int naccept = (int)st[NG + NP + 4];
  const T rtol = p.rtol, atol = p.atol;

  bool y_old_finite = true;
#pragma unroll
  for (int j = 0; j < NP; ++j) y_old_finite = y_old_finite && t_finite(y[j]);

  const T* re = p.r_eval + prob * m;
  const T* Je = p.J_eval + prob * (int64_t)m * n;
  T part = T(0);
  for (int i = lane; i < m; i += 32) {
    const T r = re[i];
    part += r * r;
  }
  const T rre = warp_sum(part);
  T d[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) d[j] = ye[j] - y[j];

  bool accept = true;
  if (p.trust) {  // trust_region.py:86-104
    T lin = T(0), quad = T(0);
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      T gd = T(0);
#pragma unroll
      for (int j = 0; j < NP; ++j) gd += G[Tri<NP>::at(i, j)] * d[j];
      quad += d[i] * gd;
      lin += d[i] * g[i];
    }
    const T pred = predicted_reduction(lin, quad, rr);
    const T df = T(0.5) * rre - T(0.5) * rr;
    accept = (df <= T(0.01) * pred) || first;
    const bool good = (df < T(0.99) * pred) && (pred < T(0)) && !first;
    delta *= accept ? (good ? T(3.5) : T(1)) : T(0.25);
  } else {
    delta = T(1);
  }

  if (accept) {  // gauss_newton.py:276-291
    NormAcc<T> yn(p.norm), fn(p.norm), zn(p.norm);
#pragma unroll
    for (int j = 0; j < NP; ++j)
      if (j < n) yn.add(t_abs(d[j]) / (atol + rtol * t_abs(ye[j])));
    T acc[NG + NP];
#pragma unroll
    for (int a = 0; a < NG + NP; ++a) acc[a] = T(0);
    for (int i = lane; i < m; i += 32) {
      T j[NP];
#pragma unroll
      for (int a = 0; a < NP; ++a) j[a] = a < n ? Je[(int64_t)i * n + a] : T(0);
      const T r = re[i], ro = st_r[i];
#pragma unroll
      for (int a = 0; a < NP; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) acc[Tri<NP>::at(a, b)] += j[a] * j[b];
        acc[NG + a] += j[a] * r;
      }
      fn.add(t_abs(r - ro) / (atol + rtol * t_abs(r)));
      if (p.root_near_zero) zn.add(t_abs(first ? r : ro));
      st_r[i] = r;
    }
    warp_allreduce_array<T, NG + NP>(acc, scratch, lane);
#pragma unroll
    for (int a = 0; a < NG; ++a) G[a] = acc[a];
#pragma unroll
    for (int a = 0; a < NP; ++a) {
      g[a] = acc[NG + a];
      if (a >= n) G[Tri<NP>::at(a, a)] = T(1);  // padded unknowns: unit pivot, zero right-hand side
    }
    // carry [G | g] back to the state buffer straight from the reduction totals (same layout), then the unit pivots
    const T* tot = scratch + 32 * WarpReduceLayout<T, NG + NP>::KP;
    for (int a = lane; a < NG + NP; a += 32) st[a] = tot[a];
    __syncwarp();
    if (lane >= n && lane < NP) st[Tri<NP>::at(lane, lane)] = T(1);
    fn.warp_combine();
    terminate = (yn.finish(n) < T(1)) && (fn.finish(m) < T(1));
    if (p.root_near_zero) {
      zn.warp_combine();
      terminate = terminate && (zn.finish(m) < atol);
    }
    ++naccept;
#pragma unroll
    for (int j = 0; j < NP; ++j) y[j] = ye[j];
    rr = rre;
  } else {
    terminate = false;
  }

  T lam = T(0);
  if (p.trust) lam = (delta > Num<T>::eps()) ? T(1) / delta : Num<T>::max();
  T lamv = lam;
  T pstep[NP];
  spd_solve<T, NP>(G, lamv, g, pstep);
  bool pfin = true;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    pfin = pfin && t_finite(pstep[j]);
    ye[j] = y[j] - pstep[j];
  }
  result = pfin ? RES_OK : RES_SINGULAR;
  if (!y_old_finite) result = RES_NONFINITE;
  ++steps;

  // write the rest of the carry back
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    if (lane == j && j < n) {
      p.y[prob * n + j] = y[j];
      p.y_eval[prob * n + j] = ye[j];
    }
  }
  if (lane == 0) {
    st[NG + NP + 0] = rr;
    st[NG + NP + 1] = delta;
    st[NG + NP + 2] = T(0);
    st[NG + NP + 3] = terminate ? T(1) : T(0);
    st[NG + NP + 4] = (T)naccept;
    p.num_steps[prob] = steps;
    p.result[prob] = result;
    if (p.active_count && !terminate && result == RES_OK && steps < p.max_steps) atomicAdd(p.active_count, 1);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 16 < n <= 64: the Gram matrix no longer fits one thread's registers.  One CTA (128 threads) per problem, G in shared
// memory / in the state buffer as a full n x n matrix; same state machine, same carry semantics as lsq_step_kernel.
//   state: [ G (n*n) | g (n) | rr | delta | first | terminate | num_accepted | r_accepted (m) ]
constexpr int kBigThreads = 128;
constexpr int kBigNMax = 64;
__host__ __device__ inline int big_state_head(int n) { return n * n + n + 5; }

template <typename T> __device__ __forceinline__ T big_block_sum(T v, T* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}
template <typename T> __device__ __forceinline__ T big_block_nanmax(T v, T* red) {
  v = warp_nanmax(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return nan_max(nan_max(red[0], red[1]), nan_max(red[2], red[3]));
}
// norm of a non-negative quantity distributed over the threads (max / two / rms, NaN-propagating max)
template <typename T> __device__ __forceinline__ T big_norm(T local, int kind, int count, T* red) {
  if (kind == B200OPT_NORM_MAX) return big_block_nanmax(local, red);
  const T s = big_block_sum(local, red);
  return kind == B200OPT_NORM_TWO ? t_sqrt(s) : t_sqrt(s / T(count));
}

template <typename T> __global__ void __launch_bounds__(kBigThreads) lsq_step_big_kernel(const StepParams<T> p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ T red[4];
  const int n = p.n, m = p.m, tid = threadIdx.x, ld = n + 1;
  T* Gs = reinterpret_cast<T*>(smem_raw);  // [n][n+1]: G, later the Cholesky factor
  T* gs = Gs + n * ld;                     // g
  T* dv = gs + n;                          // d = y_eval - y
  T* pv = dv + n;                          // solve scratch / step
  const int64_t prob = blockIdx.x;
  T* st = p.state + prob * p.state_stride;
  T* stG = st;
  T* stg = st + n * n;
  T* sc = stg + n;  // rr, delta, first, terminate, num_accepted
  T* st_r = sc + 5;
  int result = p.result[prob];
  int steps = p.num_steps[prob];
  bool terminate = sc[3] != T(0);
  if (terminate || result != RES_OK || steps >= p.max_steps) return;  // block-uniform
  T rr = sc[0], delta = sc[1];
  const bool first = sc[2] != T(0);
  const T rtol = p.rtol, atol = p.atol;
  const T* re = p.r_eval + prob * m;
  const T* Je = p.J_eval + prob * (int64_t)m * n;

  bool fin = true;
  for (int j = tid; j < n; j += kBigThreads) {
    const T yj = p.y[prob * n + j];
    dv[j] = p.y_eval[prob * n + j] - yj;
    gs[j] = stg[j];
    fin = fin && t_finite(yj);
  }
  for (int e = tid; e < n * n; e += kBigThreads) Gs[(e / n) * ld + e % n] = stG[e];
  const bool y_old_finite = __syncthreads_and(fin ? 1 : 0) != 0;
  T part = T(0);
  for (int i = tid; i < m; i += kBigThreads) part += re[i] * re[i];
  const T rre = big_block_sum(part, red);

  bool accept = true;
  if (p.trust) {  // trust_region.py:86-104, predicted reduction from d.g and d.G d (common.cuh predicted_reduction)
    T lq = T(0);
    for (int i = tid; i < n; i += kBigThreads) {
      T gd = T(0);
      for (int j = 0; j < n; ++j) gd += Gs[i * ld + j] * dv[j];
      lq += dv[i] * (T(2) * gs[i] + gd);
    }
    const T pred = predicted_reduction(T(0), big_block_sum(lq, red), rr);
    const T df = T(0.5) * rre - T(0.5) * rr;
    accept = (df <= T(0.01) * pred) || first;
    const bool good = (df < T(0.99) * pred) && (pred < T(0)) && !first;
    delta *= accept ? (good ? T(3.5) : T(1)) : T(0.25);
  } else {
    delta = T(1);
  }

  if (accept) {  // gauss_newton.py:276-291
    T yl = T(0);
    for (int j = tid; j < n; j += kBigThreads) {
      const T q = t_abs(dv[j]) / (atol + rtol * t_abs(p.y_eval[prob * n + j]));
      yl = (p.norm == B200OPT_NORM_MAX) ? nan_max(yl, q) : yl + q * q;
    }
    const T ynorm = big_norm(yl, p.norm, n, red);
    // Gram matrix and gradient of the accepted point: entry e = a * n + b by thread e % 128 (coalesced over b)
    __syncthreads();
    for (int e = tid; e < n * n; e += kBigThreads) {
      const int a = e / n, b = e % n;
      T acc = T(0);
      for (int i = 0; i < m; ++i) acc += Je[(int64_t)i * n + a] * Je[(int64_t)i * n + b];
      Gs[a * ld + b] = acc;
      stG[e] = acc;
    }
    for (int a = tid; a < n; a += kBigThreads) {
      T acc = T(0);
      for (int i = 0; i < m; ++i) acc += Je[(int64_t)i * n + a] * re[i];
      gs[a] = acc;
      stg[a] = acc;
    }
    T fl = T(0), zl = T(0);
    for (int i = tid; i < m; i += kBigThreads) {
      const T r = re[i], ro = st_r[i];
      const T q = t_abs(r - ro) / (atol + rtol * t_abs(r));
      const T z = t_abs(first ? r : ro);
      if (p.norm == B200OPT_NORM_MAX) {
        fl = nan_max(fl, q);
        zl = nan_max(zl, z);
      } else {
        fl += q * q;
        zl += z * z;
      }
      st_r[i] = r;
    }
    const T fnorm = big_norm(fl, p.norm, m, red);
    terminate = (ynorm < T(1)) && (fnorm < T(1));
    if (p.root_near_zero) terminate = terminate && (big_norm(zl, p.norm, m, red) < atol);  // _root_find.py:64-73
    rr = rre;
  } else {
    terminate = false;
  }
  __syncthreads();

  // ---- descent: (G + lam I) p = g by Cholesky in shared memory (levenberg_marquardt.py:55-75 in normal-equation form)
  T lam = T(0);
  if (p.trust) lam = (delta > Num<T>::eps()) ? T(1) / delta : Num<T>::max();
  for (int j = tid; j < n; j += kBigThreads) Gs[j * ld + j] += lam;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const T dk = t_sqrt(Gs[k * ld + k]);  // non-positive pivot -> NaN -> non-finite step -> `singular`
    __syncthreads();
    for (int i = k + tid; i < n; i += kBigThreads) Gs[i * ld + k] = (i == k) ? dk : Gs[i * ld + k] / dk;
    __syncthreads();
    const int rem = n - k - 1;
    for (int e = tid; e < rem * rem; e += kBigThreads) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j <= i) Gs[i * ld + j] -= Gs[i * ld + k] * Gs[j * ld + k];
    }
    __syncthreads();
  }
  if (tid < 32) {  // substitutions: one warp, serial in k
    const int lane = tid;
    for (int i = lane; i < n; i += 32) pv[i] = gs[i];
    __syncwarp();
    for (int k = 0; k < n; ++k) {
      const T zk = pv[k] / Gs[k * ld + k];
      __syncwarp();
      if (lane == 0) pv[k] = zk;
      for (int i = k + 1 + lane; i < n; i += 32) pv[i] -= Gs[i * ld + k] * zk;
      __syncwarp();
    }
    for (int k = n - 1; k >= 0; --k) {
      const T xk = pv[k] / Gs[k * ld + k];
      __syncwarp();
      if (lane == 0) pv[k] = xk;
      for (int i = lane; i < k; i += 32) pv[i] -= Gs[k * ld + i] * xk;
      __syncwarp();
    }
  }
  __syncthreads();
  bool pfin = true;
  for (int j = tid; j < n; j += kBigThreads) {
    const T yj = accept ? p.y_eval[prob * n + j] : p.y[prob * n + j];
    pfin = pfin && t_finite(pv[j]);
    p.y[prob * n + j] = yj;
    p.y_eval[prob * n + j] = yj - pv[j];
  }
  const bool all_fin = __syncthreads_and(pfin ? 1 : 0) != 0;
  result = all_fin ? RES_OK : RES_SINGULAR;
  if (!y_old_finite) result = RES_NONFINITE;
  ++steps;
  if (tid == 0) {
    sc[0] = rr;
    sc[1] = delta;
    sc[2] = T(0);
    sc[3] = terminate ? T(1) : T(0);
    if (accept) sc[4] = sc[4] + T(1);
    p.num_steps[prob] = steps;
    p.result[prob] = result;
    if (p.active_count && !terminate && result == RES_OK && steps < p.max_steps) atomicAdd(p.active_count, 1);
  }
}

template <typename T> __global__ void lsq_step_init_kernel(int64_t batch, int n, int head, int64_t stride, int ng_np, const T* y0,
                                                           T* y, T* y_eval, T* state, int32_t* num_steps, int32_t* result) {
  const int64_t prob = blockIdx.x;
  if (prob >= batch) return;
  T* st = state + prob * stride;
  for (int64_t i = threadIdx.x; i < stride; i += blockDim.x) st[i] = T(0);
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
    st[ng_np + 1] = T(1);  // delta = 1 (trust_region.py:75)
    st[ng_np + 2] = T(1);  // first_step = True (gauss_newton.py:231)
    num_steps[prob] = 0;
    result[prob] = allfin ? RES_OK : RES_NONFINITE;  // _iterate.py:216-219
  }
  (void)head;
}

template <typename T> __global__ void lsq_step_finish_kernel(int64_t batch, int64_t stride, int term_off, const T* state, int32_t* result) {
  const int64_t prob = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (prob >= batch) return;
  const bool terminate = state[prob * stride + term_off] != T(0);
  if (result[prob] == RES_OK && !terminate) result[prob] = RES_NL_MAX_STEPS;  // _iterate.py:253-257
}

// register kernel: n padded to 4 / 8 / 16; 0 = CTA kernel.  fp64 with 8 < n <= 16 takes the CTA kernel too: 136 + 16 packed Gram
// doubles per thread do not fit the register file (the <double, 16> instantiation spilled 20 KB per thread).
static int pad_n(int n, bool f64) { return n <= 4 ? 4 : (n <= 8 ? 8 : ((n <= 16 && !f64) ? 16 : 0)); }
static int head_for(int np) { return np == 4 ? step_state_head<4>() : (np == 8 ? step_state_head<8>() : step_state_head<16>()); }
// state head (elements before the accepted residuals) for any supported n: register kernel up to 16, CTA kernel up to 64
static int head_any(int n, bool f64) { return pad_n(n, f64) ? head_for(pad_n(n, f64)) : big_state_head(n); }

template <typename T, int NP> static int step_launch(const StepParams<T>& p, cudaStream_t stream) {
  auto kern = lsq_step_kernel<T, NP>;
  const size_t smem = sizeof(T) * kStepWarps * WarpReduceLayout<T, Tri<NP>::size + NP>::elems;
  static LaunchCache cache;
  if (cache.configure(kern, kStepWarps * 32, smem) < 0) return B200OPT_ECUDA;
  const int grid = (int)((p.batch + kStepWarps - 1) / kStepWarps);
  kern<<<grid, kStepWarps * 32, smem, stream>>>(p);
  return check_cuda(cudaGetLastError(), "lsq_step_kernel launch");
}

template <typename T>
static int step_typed(int solver, int64_t batch, int n, int m, const void* r_eval, const void* J_eval, void* y, void* y_eval,
                      void* state, const b200opt_options_t* o, int32_t* num_steps, int32_t* result, int32_t* active, cudaStream_t st) {
  const int np = pad_n(n, sizeof(T) == 8);
  StepParams<T> p;
  const int head_sel = head_any(n, sizeof(T) == 8);
  p.r_eval = static_cast<const T*>(r_eval);
  p.J_eval = static_cast<const T*>(J_eval);
  p.y = static_cast<T*>(y);
  p.y_eval = static_cast<T*>(y_eval);
  p.state = static_cast<T*>(state);
  p.state_stride = step_state_stride(head_sel, m);
  p.num_steps = num_steps;
  p.result = result;
  p.active_count = active;
  p.batch = batch;
  p.n = n;
  p.m = m;
  p.rtol = (T)o->rtol;
  p.atol = (T)o->atol;
  p.norm = o->norm;
  p.max_steps = o->max_steps;
  p.root_near_zero = o->root_near_zero;
  p.trust = solver == B200OPT_SOLVER_LEVENBERG_MARQUARDT;
  if (active && check_cuda(cudaMemsetAsync(active, 0, sizeof(int32_t), st), "memset active_count")) return B200OPT_ECUDA;
  if (np == 0) {  // 16 < n <= 64 (fp64: 8 < n <= 64)
    const size_t smem = sizeof(T) * ((size_t)n * (n + 1) + 3 * (size_t)n);
    lsq_step_big_kernel<T><<<(unsigned)batch, kBigThreads, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "lsq_step_big_kernel launch");
  }
  if (np == 4) return step_launch<T, 4>(p, st);
  if (np == 8) return step_launch<T, 8>(p, st);
  if constexpr (sizeof(T) == 4) return step_launch<T, 16>(p, st);
  set_error("lsq_step: internal dispatch error (n=%d)", n);
  return B200OPT_EINVAL;
}
}  // namespace b200opt

using namespace b200opt;

static int step_check(const char* who, int dtype, int64_t batch, int n, int m) {
  if (dtype != B200OPT_F32 && dtype != B200OPT_F64) {
    set_error("%s: unknown dtype %d", who, dtype);
    return B200OPT_EINVAL;
  }
  if (batch < 0 || batch > 2147483647LL || m < 1 || n < 1 || n > kBigNMax) {
    set_error("%s: need 0 <= batch < 2^31, m >= 1 and 1 <= n <= 64 (n=%d, m=%d)", who, n, m);
    return B200OPT_EINVAL;
  }
  return 0;
}

extern "C" {

int64_t b200opt_lsq_step_state_size(int32_t dtype, int32_t n, int32_t m) {
  if (step_check("lsq_step_state_size", dtype, 0, n, m)) return -1;
  return step_state_stride(head_any(n, dtype == B200OPT_F64), m) * (dtype == B200OPT_F32 ? 4 : 8);
}

int32_t b200opt_lsq_step_init(int32_t dtype, int64_t batch, int32_t n, int32_t m, const void* y0, void* y, void* y_eval,
                              void* state, int32_t* num_steps, int32_t* result, void* stream) {
  if (int e = step_check("lsq_step_init", dtype, batch, n, m)) return e;
  if (batch == 0) return 0;
  const int head = head_any(n, dtype == B200OPT_F64);
  const int64_t stride = step_state_stride(head, m);
  const int ng_np = head - 5;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200OPT_F64)
    lsq_step_init_kernel<double><<<(unsigned)batch, 64, 0, st>>>(batch, n, head, stride, ng_np, static_cast<const double*>(y0),
                                                                 static_cast<double*>(y), static_cast<double*>(y_eval),
                                                                 static_cast<double*>(state), num_steps, result);
  else
    lsq_step_init_kernel<float><<<(unsigned)batch, 64, 0, st>>>(batch, n, head, stride, ng_np, static_cast<const float*>(y0),
                                                                static_cast<float*>(y), static_cast<float*>(y_eval),
                                                                static_cast<float*>(state), num_steps, result);
  return check_cuda(cudaGetLastError(), "lsq_step_init launch");
}

int32_t b200opt_lsq_step(int32_t solver, int32_t dtype, int64_t batch, int32_t n, int32_t m, const void* r_eval,
                         const void* J_eval, void* y, void* y_eval, void* state, const b200opt_options_t* opts,
                         int32_t* num_steps, int32_t* result, int32_t* active_count, void* stream) {
  if (int e = step_check("lsq_step", dtype, batch, n, m)) return e;
  if (!opts || opts->norm < 0 || opts->norm > 2 || opts->max_steps < 0) {
    set_error("lsq_step: bad options");
    return B200OPT_EINVAL;
  }
  if (opts->linear_solver != B200OPT_LINSOLVE_NORMAL_CHOLESKY) {
    set_error("lsq_step: the per-iteration step kernel carries (G, g) and solves the normal equations; linear_solver = QR is available "
              "in the persistent kernels of the registered families (b200opt_least_squares)");
    return B200OPT_EINVAL;
  }
  if (solver != B200OPT_SOLVER_LEVENBERG_MARQUARDT && solver != B200OPT_SOLVER_GAUSS_NEWTON) {
    set_error("lsq_step: solver %d is not a least-squares solver", solver);
    return B200OPT_EINVAL;
  }
  if (batch == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200OPT_F64)
    return step_typed<double>(solver, batch, n, m, r_eval, J_eval, y, y_eval, state, opts, num_steps, result, active_count, st);
  return step_typed<float>(solver, batch, n, m, r_eval, J_eval, y, y_eval, state, opts, num_steps, result, active_count, st);
}

int32_t b200opt_lsq_step_finish(int32_t dtype, int64_t batch, int32_t n, int32_t m, const void* state, int32_t* result,
                                void* stream) {
  if (int e = step_check("lsq_step_finish", dtype, batch, n, m)) return e;
  if (batch == 0) return 0;
  const int head = head_any(n, dtype == B200OPT_F64);
  const int64_t stride = step_state_stride(head, m);
  const int term_off = head - 5 + 3;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((batch + 127) / 128);
  if (dtype == B200OPT_F64)
    lsq_step_finish_kernel<double><<<grid, 128, 0, st>>>(batch, stride, term_off, static_cast<const double*>(state), result);
  else
    lsq_step_finish_kernel<float><<<grid, 128, 0, st>>>(batch, stride, term_off, static_cast<const float*>(state), result);
  return check_cuda(cudaGetLastError(), "lsq_step_finish launch");
}

}  // extern "C"
