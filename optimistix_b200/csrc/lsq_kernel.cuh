// Batched Levenberg-Marquardt / Gauss-Newton: one warp per problem, persistent grid-stride kernel.
//
// Restates, per problem, the flat state machine of SURVEY.md section 3.1:
//   driver        optimistix/_iterate.py:194-268
//   step          optimistix/_solver/gauss_newton.py:243-334
//   search        optimistix/_solver/trust_region.py:77-106,130-193 (LM) / learning_rate.py:24-35 (GN)
//   descent       optimistix/_solver/levenberg_marquardt.py:34-76,126-133 (LM) / gauss_newton.py:64-127 (GN)
//   termination   optimistix/_misc.py:291-310
//
// B200 design (not a translation):
//   * the whole solve of a problem stays on chip.  Registers (replicated in every lane): y, y_eval, the packed
//     Gram matrix G = J^T J, g = J^T r, the radius.  Shared memory (per warp, lane-interleaved so every access is
//     conflict free): this problem's data row, the accepted and the trial residuals (ping-pong), the model's
//     cached per-row intermediates.  HBM is touched twice per problem: one coalesced (128-bit where aligned)
//     read of y0/data, one write of y/num_steps/result.
//   * J is never materialised.  The accepted (r, J) of the reference is carried as (r, G, g, r^T r): the
//     predicted reduction 0.5(|J d + r|^2 - |r|^2) of trust_region.py:182-186 is evaluated as the
//     algebraically identical  d.g + 0.5 d.G d,  and the augmented QR solve of
//     levenberg_marquardt.py:66-75 as the Cholesky solve of (G + lambda I) p = g  (BASELINE.json north_star;
//     D = I as in the reference).
//   * the Jacobian/Gram pass only runs for ACCEPTED trial points (the reference linearises every step and
//     discards rejected ones; values are identical, work is not).
//   * warp-uniform control flow: one problem per warp means accept/reject and the iteration count never
//     diverge inside a warp.  Row loops are rolled (rows live in shared memory, not in registers), which keeps
//     the kernel at ~100 registers (16 resident warps/SM) and inside the instruction cache for any m.
#pragma once
#include "common.cuh"
#include "models.cuh"

namespace b200opt {

template <typename T> struct LsqParams {
  const T* y0;
  const T* shared;        // abscissa grid [m]
  const T* prob;          // data [batch][prob_stride]
  int64_t prob_stride;
  T* y_out;
  int32_t* num_steps;
  int32_t* result;
  T* step_size_out;        // optional
  int32_t* num_accepted;   // optional
  int64_t batch;
  int m;
  int m_pad;               // m rounded up to a multiple of 32 (row pitch of the per-warp shared arrays)
  T rtol, atol;
  int norm;
  int max_steps;
  int root_near_zero;
  int vec_ok;              // data rows 16-byte aligned and m a multiple of the vector width
  unsigned long long* queue;  // ticket counter of the dynamic problem queue (common.cuh), or NULL: static grid-stride
  int data_kind;           // B200OPT_DATA_*: 0 = `prob` holds T; otherwise 16-bit values (prob_stride in 16-bit elements)
  int data_vec;            // 16-bit rows: 16-byte aligned and m % 8 == 0
};

// QR mode (B200OPT_LINSOLVE_QR): the reference's own operations instead of the normal equations.  The Jacobian of the
// accepted point is kept in shared memory (column-major, one row per lane and slot: conflict free), the predicted reduction
// is 0.5 (|J d + r|^2 - |r|^2) from the rows themselves (trust_region.py:182-186), and the descent is the Householder QR of the
// augmented system [J; sqrt(lambda) I] p = [r; 0] (levenberg_marquardt.py:66-75; lx.QR = geqrf + Q^T b + triangular solve) on
// a work copy: rows 0..m-1 are J, rows m_pad..m_pad+N-1 (an extra group of 32 rows, lane a holds row a) are sqrt(lambda) I.
// kappa(J) enters once (not squared as in G = J^T J); the price is 2 N (m_pad + 32) more shared-memory elements per warp and a
// few times the arithmetic per step.
__host__ __device__ inline int lsq_qr_rows(int m_pad) { return m_pad + 32; }

constexpr int kLsqWarpsPerBlock = 4;

// ---- warp all-reduce of K per-lane partial sums through shared memory, 128-bit accesses ---------------------------
// Layout: row `lane` of the scratch holds that lane's K partials (pitch KP chosen so 128-bit row accesses of the
// 8 lanes of a quarter-warp fall in distinct bank groups).  Column sums: lane (q, h) sums quad-column q over the
// rows of half h in {0,1} (rows h*16 .. h*16+15), halves are combined with one shuffle, totals are written once
// and read back by every lane as broadcasts.  Order of additions is fixed (deterministic results).
template <typename T, int K> struct WarpReduceLayout {
  static constexpr int V = 16 / sizeof(T);                       // elements per 128-bit access
  static constexpr int KQ = (K + V - 1) / V;                     // quads per row
  static constexpr int KPAD = KQ * V;
  static constexpr int KP = KPAD + V * ((KQ % 2 == 0) ? 1 : 0);  // odd number of quads per row -> conflict free
  static constexpr int elems = 32 * KP + KPAD;
};

template <typename T, int K> __device__ __forceinline__ void warp_allreduce_array(T (&v)[K], T* sm, int lane) {
  using L = WarpReduceLayout<T, K>;
  using Vec = typename Num<T>::Vec;
  constexpr int V = L::V;
  T* row = sm + lane * L::KP;
#pragma unroll
  for (int q = 0; q < L::KQ; ++q) {
    Vec w;
    T* ww = reinterpret_cast<T*>(&w);
#pragma unroll
    for (int e = 0; e < V; ++e) ww[e] = (q * V + e < K) ? v[q * V + e] : T(0);
    *reinterpret_cast<Vec*>(row + q * V) = w;
  }
  __syncwarp();
  T* tot = sm + 32 * L::KP;
  // lanes 0-15 sum rows 0-15 of quad-column q, lanes 16-31 rows 16-31 of the same quad (a quarter-warp reads 8
  // consecutive quads of one row: conflict free); the halves meet through one xor-16 shuffle per element.
  // (no divergence: lanes beyond the last quad column recompute that column - broadcast loads, nothing extra in shared memory -
  //  and only the store is predicated; a branch around the block cost a reconvergence point per reduction)
  for (int qb = 0; qb < L::KQ; qb += 16) {
    const int q0 = qb + (lane & 15), h = lane >> 4;
    const bool on = q0 < L::KQ;
    const int q = on ? q0 : L::KQ - 1;
    T a[V];
#pragma unroll
    for (int e = 0; e < V; ++e) a[e] = T(0);
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      const Vec w = *reinterpret_cast<const Vec*>(sm + (h * 16 + l) * L::KP + q * V);
      const T* ww = reinterpret_cast<const T*>(&w);
#pragma unroll
      for (int e = 0; e < V; ++e) a[e] += ww[e];
    }
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const T o = __shfl_xor_sync(kFull, a[e], 16);
      a[e] = (h == 0) ? a[e] + o : o + a[e];  // rows 0-15 first, then 16-31: same order in both halves
    }
    if (h == 0 && on) {
      Vec w;
      T* ww = reinterpret_cast<T*>(&w);
#pragma unroll
      for (int e = 0; e < V; ++e) ww[e] = a[e];
      *reinterpret_cast<Vec*>(tot + q * V) = w;
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < L::KQ; ++q) {
    const Vec w = *reinterpret_cast<const Vec*>(tot + q * V);
    const T* ww = reinterpret_cast<const T*>(&w);
#pragma unroll
    for (int e = 0; e < V; ++e)
      if (q * V + e < K) v[q * V + e] = ww[e];
  }
  __syncwarp();
}

// per-warp shared memory (in elements of T): reduction scratch | data d | residual A | residual B | aux[NAUX]
template <class Model, typename T> __host__ __device__ constexpr int lsq_scratch_elems() {
  return WarpReduceLayout<T, Tri<Model::N>::size + Model::N>::elems;
}
template <class Model, typename T> __host__ __device__ inline size_t lsq_smem_bytes(int m_pad, bool qr = false) {
  const size_t per_warp = lsq_scratch_elems<Model, T>() + (size_t)(3 + Model::NAUX) * m_pad + (qr ? (size_t)2 * Model::N * lsq_qr_rows(m_pad) : 0);
  return sizeof(T) * ((size_t)m_pad /* shared abscissa grid */ + kLsqWarpsPerBlock * per_warp);
}

#ifndef LSQ_MINBLOCKS_F64
#define LSQ_MINBLOCKS_F64 2
#endif
template <class Model, typename T, bool TRUST, bool QRMODE = false>
__global__ void __launch_bounds__(kLsqWarpsPerBlock * 32, (QRMODE ? 1 : (sizeof(T) == 4 ? 4 : LSQ_MINBLOCKS_F64))) lsq_solve_kernel(const LsqParams<T> p) {
  constexpr int N = Model::N;
  constexpr int NG = Tri<N>::size;
  constexpr int NAUX = Model::NAUX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int m = p.m, mp = p.m_pad;
  T* xs = reinterpret_cast<T*>(smem_raw);
  const int ma = lsq_qr_rows(mp);  // QR mode: rows of the augmented work matrix
  T* wbase = xs + mp + (size_t)wib * (lsq_scratch_elems<Model, T>() + (size_t)(3 + NAUX) * mp + (QRMODE ? (size_t)2 * N * ma : 0));
  T* scratch = wbase;
  T* ds = wbase + lsq_scratch_elems<Model, T>();
  T* racc = ds + mp;       // residuals at the last accepted point
  T* rtry = racc + mp;     // residuals at the trial point
  T* auxs = rtry + mp;     // NAUX arrays of m_pad
  T* Jacc = auxs + (size_t)NAUX * mp;  // QR mode: Jacobian at the accepted point, [N][ma] column-major
  T* Wq = Jacc + (size_t)N * ma;       // QR mode: work copy of the augmented matrix
  const int64_t warp0 = (int64_t)blockIdx.x * kLsqWarpsPerBlock + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kLsqWarpsPerBlock;
  const T rtol = p.rtol, atol = p.atol;
  const bool maxnorm = p.norm == B200OPT_NORM_MAX;

  for (int i = threadIdx.x; i < mp; i += blockDim.x) xs[i] = p.shared[i < m ? i : m - 1];
  __syncthreads();

  for (int64_t prob = warp0; prob < p.batch; prob = next_problem_warp(p.queue, prob, nwarps, lane)) {
    // ---------------- load: y0 (broadcast) and the data row (coalesced; 128-bit when aligned) into shared memory
    T y[N], ye[N];
#pragma unroll
    for (int j = 0; j < N; ++j) y[j] = ye[j] = __ldg(p.y0 + prob * N + j);
    const T* drow = p.prob + prob * p.prob_stride;
    if (p.data_kind != B200OPT_DATA_AS_COMPUTE) {
      load_row16<T>(p.prob, prob, p.prob_stride, m, p.data_kind, p.data_vec != 0, ds, lane);
    } else if (p.vec_ok) {
      using Vec = typename Num<T>::Vec;
      constexpr int V = Num<T>::kVec;
      for (int i = lane; i < m / V; i += 32)
        reinterpret_cast<Vec*>(ds)[i] = __ldg(reinterpret_cast<const Vec*>(drow) + i);
    } else {
      for (int i = lane; i < m; i += 32) ds[i] = __ldg(drow + i);
    }
    for (int i = lane; i < mp; i += 32) racc[i] = T(0);
    if (QRMODE)
      for (int i = lane; i < N * ma; i += 32) Jacc[i] = T(0);  // gauss_newton.py:229-230: zero-filled f_info
    __syncwarp();

    // ---------------- init (gauss_newton.py:229-241, trust_region.py:75, _iterate.py:214-219)
    T G[NG], g[N];
#pragma unroll
    for (int a = 0; a < NG; ++a) G[a] = T(0);
#pragma unroll
    for (int j = 0; j < N; ++j) g[j] = T(0);
    T rr = T(0);
    T delta = T(1);
    bool first = true, terminate = false;
    bool yfin = true;
#pragma unroll
    for (int j = 0; j < N; ++j) yfin = yfin && t_finite(y[j]);
    int result = yfin ? RES_OK : RES_NONFINITE;
    int steps = 0, naccept = 0;

    while (!terminate && result == RES_OK && steps < p.max_steps) {
      // _iterate.py:216,238: finiteness of the carry's y from BEFORE this step
      bool y_old_finite = true;
#pragma unroll
      for (int j = 0; j < N; ++j) y_old_finite = y_old_finite && t_finite(y[j]);

      // ---- residual at the trial point (gauss_newton.py:253)
      const typename Model::Pre pre = Model::prepare(ye);
      T part = T(0);
#pragma unroll 2
      for (int i = lane; i < m; i += 32) {
        T aux[NAUX];
        const T r = Model::residual(ye, pre, CurveRow<T>{xs[i], ds[i]}, aux);
        rtry[i] = r;
#pragma unroll
        for (int q = 0; q < NAUX; ++q) auxs[q * mp + i] = aux[q];
        part += r * r;
      }
      const T rre = warp_sum(part);

      T d[N];
#pragma unroll
      for (int j = 0; j < N; ++j) d[j] = ye[j] - y[j];

      // ---- search (trust_region.py:86-104 with constants :125-128 / learning_rate.py:24-35)
      bool accept = true;
      if (TRUST) {
        T pred;
        if (QRMODE) {
          T sj = T(0);
          for (int i = lane; i < m; i += 32) {
            T s = T(0);
#pragma unroll
            for (int a = 0; a < N; ++a) s += Jacc[a * ma + i] * d[a];
            s += racc[i];
            sj += s * s;
          }
          pred = T(0.5) * (warp_sum(sj) - rr);  // the reference's expression, term by term
        } else {
          T lin = T(0), quad = T(0);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            T gd = T(0);
#pragma unroll
            for (int j = 0; j < N; ++j) gd += G[Tri<N>::at(i, j)] * d[j];
            quad += d[i] * gd;
            lin += d[i] * g[i];
          }
          pred = predicted_reduction(lin, quad, rr);  // 0.5(|J d + r|^2 - |r|^2), trust_region.py:182-186
        }
        const T df = T(0.5) * rre - T(0.5) * rr;
        accept = (df <= T(0.01) * pred) || first;
        const bool good = (df < T(0.99) * pred) && (pred < T(0)) && !first;
        delta *= accept ? (good ? T(3.5) : T(1)) : T(0.25);
      } else {
        delta = T(1);
      }

      // ---- accepted / rejected (gauss_newton.py:276-296)
      if (accept) {
        // y-space half of cauchy_termination (_misc.py:306-309). For max_norm, max_i(|d_i|/s_i) < 1 is evaluated
        // as all_i(|d_i| < s_i): same truth value except when a quotient rounds to exactly 1; no divisions.
        bool yconv;
        if (maxnorm) {
          yconv = true;
#pragma unroll
          for (int j = 0; j < N; ++j) yconv = yconv && (t_abs(d[j]) < atol + rtol * t_abs(ye[j]));
        } else {
          NormAcc<T> yn(p.norm);
#pragma unroll
          for (int j = 0; j < N; ++j) yn.add(t_abs(d[j]) / (atol + rtol * t_abs(ye[j])));
          yconv = yn.finish(N) < T(1);
        }
        ++naccept;
#pragma unroll
        for (int j = 0; j < N; ++j) y[j] = ye[j];
        rr = rre;
        // one fused pass over the rows at the newly accepted point: Jacobian row -> Gram/gradient partials,
        // f-space half of the termination test, and (root_find through LM) the near-zero test.
        T acc[NG + N];
#pragma unroll
        for (int a = 0; a < NG + N; ++a) acc[a] = T(0);
        NormAcc<T> fn(p.norm), zn(p.norm);
        bool fconv = true;
#pragma unroll 2
        for (int i = lane; i < m; i += 32) {
          T aux[NAUX], j[N];
#pragma unroll
          for (int q = 0; q < NAUX; ++q) aux[q] = auxs[q * mp + i];
          const T re = rtry[i], ro = racc[i];
          Model::jac_row(y, pre, CurveRow<T>{xs[i], T(0)}, aux, j);
          if (QRMODE) {
#pragma unroll
            for (int a = 0; a < N; ++a) Jacc[a * ma + i] = j[a];
          } else {
#pragma unroll
            for (int a = 0; a < N; ++a) {
#pragma unroll
              for (int b = 0; b <= a; ++b) acc[Tri<N>::at(a, b)] += j[a] * j[b];
              acc[NG + a] += j[a] * re;
            }
          }
          const T fs = atol + rtol * t_abs(re);
          if (maxnorm) fconv = fconv && (t_abs(re - ro) < fs);
          else fn.add(t_abs(re - ro) / fs);
          if (p.root_near_zero) zn.add(t_abs(first ? re : ro));
        }
        if (!QRMODE) {
          warp_allreduce_array<T, NG + N>(acc, scratch, lane);
#pragma unroll
          for (int a = 0; a < NG; ++a) G[a] = acc[a];
#pragma unroll
          for (int a = 0; a < N; ++a) g[a] = acc[NG + a];
        }
        if (maxnorm) {
          fconv = __all_sync(kFull, fconv);
        } else {
          fn.warp_combine();
          fconv = fn.finish(m) < T(1);
        }
        terminate = yconv && fconv;  // _misc.py:310
        if (p.root_near_zero) {
          // _LstsqToRoot (_root_find.py:64-73): norm(residual at the PRE-step accepted y) < atol
          zn.warp_combine();
          terminate = terminate && (zn.finish(m) < atol);
        }
        T* t = racc;  // the trial residuals become the accepted ones
        racc = rtry;
        rtry = t;
      } else {
        terminate = false;
      }

      // ---- descent (levenberg_marquardt.py:55-57,66-75,132 / gauss_newton.py:64-78,126-127)
      T lam = T(0);
      if (TRUST) lam = (delta > Num<T>::eps()) ? T(1) / delta : Num<T>::max();
      T pstep[N];
      if (QRMODE) {
        __syncwarp();
        T* bq = rtry;  // free until the next trial evaluation: right-hand side [r; 0] of the augmented system
        const T sq = t_sqrt(lam);
#pragma unroll
        for (int a = 0; a < N; ++a)
          for (int i = lane; i < ma; i += 32) Wq[a * ma + i] = (i < m) ? Jacc[a * ma + i] : ((i - mp == a) ? sq : T(0));
        for (int i = lane; i < mp; i += 32) bq[i] = (i < m) ? racc[i] : T(0);
        __syncwarp();
        // Householder reflections, column by column; the rows at or below the diagonal are all rows except 0..k-1.  The
        // right-hand side of the 32 augmented rows is zero before the first reflection and is kept in a register (baug).
        T baug = T(0);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          T sig = T(0);
          for (int i = lane; i < ma; i += 32)
            if (i >= k) sig += Wq[k * ma + i] * Wq[k * ma + i];
          sig = warp_sum(sig);
          const T xk = Wq[k * ma + k];
          const T alpha = (xk > T(0)) ? -t_sqrt(sig) : t_sqrt(sig);
          const T inv = T(1) / (sig - alpha * xk);  // 2 / v^T v; a zero column gives 0/0 -> non-finite step -> `singular`
          T dots[N + 1];
#pragma unroll
          for (int j = 0; j <= N; ++j) dots[j] = T(0);
          for (int i = lane; i < ma; i += 32) {
            if (i >= k) {
              const T vi = (i == k) ? xk - alpha : Wq[k * ma + i];
#pragma unroll
              for (int j = k + 1; j < N; ++j) dots[j] += vi * Wq[j * ma + i];
              dots[N] += vi * (i < mp ? bq[i] : baug);
            }
          }
#pragma unroll
          for (int j = k + 1; j <= N; ++j) dots[j] = warp_sum(dots[j]) * inv;
          for (int i = lane; i < ma; i += 32) {
            if (i >= k) {
              const T vi = (i == k) ? xk - alpha : Wq[k * ma + i];
#pragma unroll
              for (int j = k + 1; j < N; ++j) Wq[j * ma + i] -= vi * dots[j];
              if (i < mp) bq[i] -= vi * dots[N];
              else baug -= vi * dots[N];
            }
          }
          if (lane == k) Wq[k * ma + k] = alpha;  // R_kk (row k is slot 0 of lane k)
          __syncwarp();
        }
        // back substitution R p = (Q^T b)[0..N): R sits in rows 0..N-1
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
          T s = bq[i];
#pragma unroll
          for (int j = i + 1; j < N; ++j) s -= Wq[j * ma + i] * pstep[j];
          pstep[i] = s / Wq[i * ma + i];
        }
        __syncwarp();
      } else {
        spd_solve<T, N>(G, lam, g, pstep);
      }
      bool pfin = true;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        pfin = pfin && t_finite(pstep[j]);
        ye[j] = y[j] - pstep[j];
      }
      result = pfin ? RES_OK : RES_SINGULAR;              // lineax: non-finite solution => singular
      if (!y_old_finite) result = RES_NONFINITE;          // _iterate.py:216
      first = false;
      ++steps;
    }
    if (result == RES_OK && !terminate) result = RES_NL_MAX_STEPS;  // _iterate.py:253-257

    {  // lane j stores y[j]: a select chain and ONE coalesced store (`if (lane == j) y_out[j] = y[j]` compiles to a jump table
       // whose N divergent paths cost 3.6 % of the kernel's stall samples, profiles/lsq_r2z)
      T yv = y[0];
#pragma unroll
      for (int j = 1; j < N; ++j) yv = (lane == j) ? y[j] : yv;
      if (lane < N) p.y_out[prob * N + lane] = yv;
    }
    if (lane == 0) {
      p.num_steps[prob] = steps;
      p.result[prob] = result;
      if (p.step_size_out) p.step_size_out[prob] = delta;
      if (p.num_accepted) p.num_accepted[prob] = naccept;
    }
    __syncwarp();
  }
}

}  // namespace b200opt
