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
//   * the whole solve of a problem stays in one warp's registers: y, y_eval, the accepted residuals of
//     this lane's rows, the packed Gram matrix G = J^T J and g = J^T r.  HBM is touched twice per problem:
//     one coalesced (128-bit where aligned) read of y0/data, one write of y/num_steps/result.
//   * J is never materialised.  The accepted (r, J) of the reference is carried as (r, G, g, r^T r): the
//     predicted reduction 0.5(|J d + r|^2 - |r|^2) of trust_region.py:182-186 is evaluated as the
//     algebraically identical  d.g + 0.5 d.G d,  and the augmented QR solve of
//     levenberg_marquardt.py:66-75 as the Cholesky solve of (G + lambda I) p = g  (BASELINE.json north_star;
//     D = I as in the reference).
//   * the Jacobian/Gram pass only runs for ACCEPTED trial points (the reference linearises every step and
//     discards rejected ones; values are identical, work is not).
//   * warp-uniform control flow: one problem per warp means accept/reject and the iteration count never
//     diverge inside a warp.
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
  T rtol, atol;
  int norm;
  int max_steps;
  int root_near_zero;
  int vec_ok;              // data rows 16-byte aligned and m a multiple of the vector width
};

constexpr int kLsqWarpsPerBlock = 4;

template <class Model, typename T, int RPL> constexpr int lsq_smem_elems_per_warp() {
  return warp_allreduce_scratch<T, Tri<Model::N>::size + Model::N>();
}

// Row owned by (lane, k): vectorised layout when V > 1 so that a lane's V consecutive rows are one 128-bit load.
template <int V> __device__ __forceinline__ int lsq_row_index(int k, int lane) {
  return (k / V) * (32 * V) + lane * V + (k % V);
}

template <class Model, typename T, int RPL, bool TRUST>
__global__ void __launch_bounds__(kLsqWarpsPerBlock * 32) lsq_solve_kernel(const LsqParams<T> p) {
  constexpr int N = Model::N;
  constexpr int NG = Tri<N>::size;
  constexpr int NAUX = Model::NAUX;
  constexpr int V = (RPL >= Num<T>::kVec) ? Num<T>::kVec : 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  T* scratch = reinterpret_cast<T*>(smem_raw) + warp_in_block * lsq_smem_elems_per_warp<Model, T, RPL>();
  const int64_t warp0 = (int64_t)blockIdx.x * kLsqWarpsPerBlock + warp_in_block;
  const int64_t nwarps = (int64_t)gridDim.x * kLsqWarpsPerBlock;
  const int m = p.m;
  const T rtol = p.rtol, atol = p.atol;

  // abscissa grid for this lane's rows: identical for every problem, loaded once per warp
  T xs[RPL];
  T wgt[RPL];  // 1 for rows < m, 0 for the ragged tail (their clamped loads are multiplied away)
#pragma unroll
  for (int k = 0; k < RPL; ++k) {
    const int i = lsq_row_index<V>(k, lane);
    wgt[k] = (i < m) ? T(1) : T(0);
    xs[k] = p.shared[i < m ? i : m - 1];
  }

  for (int64_t prob = warp0; prob < p.batch; prob += nwarps) {
    // ---------------- load: y0 (broadcast) and this lane's data rows (coalesced; 128-bit when aligned)
    T y[N], ye[N];
#pragma unroll
    for (int j = 0; j < N; ++j) y[j] = ye[j] = p.y0[prob * N + j];
    CurveRow<T> rows[RPL];
    const T* drow = p.prob + prob * p.prob_stride;
    if (V > 1 && p.vec_ok) {
      using Vec = typename Num<T>::Vec;
#pragma unroll
      for (int k = 0; k < RPL; k += V) {
        const int i = lsq_row_index<V>(k, lane);
        if (i < m) {
          const Vec v = __ldg(reinterpret_cast<const Vec*>(drow + i));
          const T* vv = reinterpret_cast<const T*>(&v);
#pragma unroll
          for (int q = 0; q < V; ++q) rows[k + q].d = vv[q];
        } else {
#pragma unroll
          for (int q = 0; q < V; ++q) rows[k + q].d = T(0);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const int i = lsq_row_index<V>(k, lane);
        rows[k].d = (i < m) ? __ldg(drow + i) : T(0);
      }
    }
#pragma unroll
    for (int k = 0; k < RPL; ++k) rows[k].x = xs[k];

    // ---------------- init (gauss_newton.py:229-241, trust_region.py:75, _iterate.py:214-219)
    T G[NG], g[N], r[RPL];
#pragma unroll
    for (int a = 0; a < NG; ++a) G[a] = T(0);
#pragma unroll
    for (int j = 0; j < N; ++j) g[j] = T(0);
#pragma unroll
    for (int k = 0; k < RPL; ++k) r[k] = T(0);
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
      T re[RPL], aux[RPL][NAUX];
      T part = T(0);
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        re[k] = wgt[k] * Model::residual(ye, pre, rows[k], aux[k]);
        part += re[k] * re[k];
      }
      const T rre = warp_sum(part);

      T d[N];
#pragma unroll
      for (int j = 0; j < N; ++j) d[j] = ye[j] - y[j];

      // ---- search (trust_region.py:86-104 with constants :125-128 / learning_rate.py:24-35)
      bool accept = true;
      if (TRUST) {
        T lin = T(0), quad = T(0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
          T gd = T(0);
#pragma unroll
          for (int j = 0; j < N; ++j) gd += G[Tri<N>::at(i, j)] * d[j];
          quad += d[i] * gd;
          lin += d[i] * g[i];
        }
        const T pred = lin + T(0.5) * quad;  // == 0.5(|J d + r|^2 - |r|^2), trust_region.py:182-186
        const T df = T(0.5) * rre - T(0.5) * rr;
        accept = (df <= T(0.01) * pred) || first;
        const bool good = (df < T(0.99) * pred) && (pred < T(0)) && !first;
        delta *= accept ? (good ? T(3.5) : T(1)) : T(0.25);
      } else {
        delta = T(1);
      }

      // ---- accepted / rejected (gauss_newton.py:276-296)
      if (accept) {
        NormAcc<T> yn(p.norm), fn(p.norm);
#pragma unroll
        for (int j = 0; j < N; ++j) yn.add(t_abs(d[j]) / (atol + rtol * t_abs(ye[j])));
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          const T q = t_abs(re[k] - r[k]) / (atol + rtol * t_abs(re[k]));
          fn.add(wgt[k] != T(0) ? q : T(0));
        }
        fn.warp_combine();
        terminate = (yn.finish(N) < T(1)) && (fn.finish(m) < T(1));  // _misc.py:306-310
        if (p.root_near_zero) {
          // _LstsqToRoot (_root_find.py:64-73): norm(residual at the PRE-step accepted y) < atol
          NormAcc<T> zn(p.norm);
#pragma unroll
          for (int k = 0; k < RPL; ++k) zn.add(t_abs(first ? re[k] : r[k]));
          zn.warp_combine();
          terminate = terminate && (zn.finish(m) < atol);
        }
        ++naccept;
#pragma unroll
        for (int j = 0; j < N; ++j) y[j] = ye[j];
#pragma unroll
        for (int k = 0; k < RPL; ++k) r[k] = re[k];
        rr = rre;
        // Gram pass at the newly accepted point: acc = [G (packed lower) | g]
        T acc[NG + N];
#pragma unroll
        for (int a = 0; a < NG + N; ++a) acc[a] = T(0);
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          T j[N];
          Model::jac_row(y, pre, rows[k], aux[k], j);
#pragma unroll
          for (int a = 0; a < N; ++a) j[a] *= wgt[k];
#pragma unroll
          for (int a = 0; a < N; ++a) {
#pragma unroll
            for (int b = 0; b <= a; ++b) acc[Tri<N>::at(a, b)] += j[a] * j[b];
            acc[NG + a] += j[a] * re[k];
          }
        }
        warp_allreduce_array<T, NG + N>(acc, scratch, lane);
#pragma unroll
        for (int a = 0; a < NG; ++a) G[a] = acc[a];
#pragma unroll
        for (int a = 0; a < N; ++a) g[a] = acc[NG + a];
      } else {
        terminate = false;
      }

      // ---- descent (levenberg_marquardt.py:55-57,66-75,132 / gauss_newton.py:64-78,126-127)
      T lam = T(0);
      if (TRUST) lam = (delta > Num<T>::eps()) ? T(1) / delta : Num<T>::max();
      T pstep[N];
      spd_solve<T, N>(G, lam, g, pstep);
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

#pragma unroll
    for (int j = 0; j < N; ++j)
      if (lane == j) p.y_out[prob * N + j] = y[j];
    if (lane == 0) {
      p.num_steps[prob] = steps;
      p.result[prob] = result;
      if (p.step_size_out) p.step_size_out[prob] = delta;
      if (p.num_accepted) p.num_accepted[prob] = naccept;
    }
  }
}

}  // namespace b200opt
