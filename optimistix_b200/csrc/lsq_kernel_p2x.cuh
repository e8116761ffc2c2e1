// Packed-fp32 LM / GN kernel, TWO problems per warp ("slots").
//
// lsq_solve_kernel_p2 (one problem per warp) spends about half of its ~900 instructions per problem-iteration in the 8 x 8
// algebra (predicted reduction, Cholesky, substitutions), replicated in all 32 lanes because the row sweeps that follow need
// the result everywhere.  Earlier attempts to share that work lost: splitting the ROWS of a problem over
// half-warps idles lanes when the two problems diverge, moving the algebra of several problems into one warp costs block
// barriers.  This kernel keeps what worked and changes only the replication:
//   * a warp owns two problems.  The row sweeps (residual; Jacobian -> Gram partials) of a slot use all 32 lanes exactly as
//     before - two rows per instruction, conflict-free 64-bit shared-memory accesses - one slot after the other;
//   * the algebra of the two slots runs AT THE SAME TIME in the two half-warps: lanes 0-15 carry slot 0's (y, y_eval, radius,
//     flags), lanes 16-31 slot 1's, one instruction stream, no divergence (accept / reject only decides, warp-uniformly per slot,
//     whether that slot's Gram sweep runs).  The per-problem algebra cost halves;
//   * G = J^T J and g = J^T r of a slot live in shared memory between iterations (the warp all-reduce leaves its totals there and
//     the algebra reads them back with broadcast loads), so a lane holds no Gram registers while it sweeps the other slot's rows;
//   * a slot that finishes writes its result and takes the next problem from the queue on its own: the two slots of a warp are
//     independent state machines, and a warp keeps 2 x (data row, two residual rows, cached exponentials) in shared memory:
//     3 CTAs = 24 resident problems per SM instead of 16.
// The arithmetic of a problem - every sum, its order, every rounding - is that of lsq_solve_kernel_p2: results are bit-identical
// (tools/lm_ab_check.py).
#pragma once
#include "lsq_kernel_p2.cuh"

namespace b200opt {

// warp_allreduce_array (lsq_kernel.cuh) that leaves the K totals in `tot` (shared memory) instead of reading them back
template <int K> __device__ __forceinline__ void warp_reduce_to_smem(const float (&v)[K], float* sm, float* tot, int lane) {
  using L = WarpReduceLayout<float, K>;
  float* row = sm + lane * L::KP;
#pragma unroll
  for (int q = 0; q < L::KQ; ++q) {
    float4 w;
    float* ww = reinterpret_cast<float*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) ww[e] = (q * 4 + e < K) ? v[q * 4 + e] : 0.f;
    *reinterpret_cast<float4*>(row + q * 4) = w;
  }
  __syncwarp();
  for (int qb = 0; qb < L::KQ; qb += 16) {
    const int q = qb + (lane & 15), h = lane >> 4;
    const bool on = q < L::KQ;
    const unsigned mask = __ballot_sync(kFull, on);
    if (on) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int l = 0; l < 16; ++l) {
        const float4 w = *reinterpret_cast<const float4*>(sm + (h * 16 + l) * L::KP + q * 4);
        a[0] += w.x, a[1] += w.y, a[2] += w.z, a[3] += w.w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float o = __shfl_xor_sync(mask, a[e], 16);
        a[e] = (h == 0) ? a[e] + o : o + a[e];  // rows 0-15 first, then 16-31: same order in both halves
      }
      if (h == 0) *reinterpret_cast<float4*>(tot + q * 4) = make_float4(a[0], a[1], a[2], a[3]);
    }
  }
  __syncwarp();
}

template <class PModel> __host__ __device__ inline size_t lsq_p2x_smem_bytes(int m_pad) {
  using L = WarpReduceLayout<float, Tri<PModel::N>::size + PModel::N>;
  const size_t per_warp = (size_t)32 * L::KP + 2 * L::KPAD + 2 * (size_t)(3 + PModel::NAUX) * m_pad;
  return sizeof(float) * ((size_t)m_pad + kLsqWarpsPerBlock * per_warp);
}

template <class PModel, bool TRUST, int NPAIRS>
__global__ void __launch_bounds__(kLsqWarpsPerBlock * 32, 3) lsq_solve_kernel_p2x(const LsqParams<float> p) {
  using T = float;
  constexpr int N = PModel::N;
  constexpr int NG = Tri<N>::size;
  constexpr int NAUX = PModel::NAUX;
  constexpr int K = NG + N;
  using RL = WarpReduceLayout<T, K>;
  constexpr int KPAD = RL::KPAD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int half = lane >> 4;       // the slot this lane carries the state of
  const int src0 = 0, src1 = 16;    // first lane of each half: where the other lanes read a slot's state from
  const int m = p.m, mp = p.m_pad;  // m % 64 == 0, m_pad == m
  const size_t slot_elems = (size_t)(3 + NAUX) * mp;
  T* xs = reinterpret_cast<T*>(smem_raw);
  T* wbase = xs + mp + (size_t)wib * ((size_t)32 * RL::KP + 2 * KPAD + 2 * slot_elems);
  T* scratch = wbase;                     // 32 rows of the reduction
  T* totals = wbase + 32 * RL::KP;        // [2][KPAD]: G (packed lower) and g of the two slots
  T* slots = totals + 2 * KPAD;           // [2][d | residual 0 | residual 1 | aux[NAUX]]
  const int64_t warp0 = (int64_t)blockIdx.x * kLsqWarpsPerBlock + wib;
  const int64_t nworkers = (int64_t)gridDim.x * kLsqWarpsPerBlock * 2;  // slots in the grid
  const T rtol = p.rtol, atol = p.atol;
  const int npairs = NPAIRS > 0 ? NPAIRS : (m >> 6);

  for (int i = threadIdx.x; i < mp; i += blockDim.x) xs[i] = p.shared[i];
  __syncthreads();
  const float2* xs2 = reinterpret_cast<const float2*>(xs);
  const T* mytot = totals + half * KPAD;

  // ---- state of this lane's slot
  T y[N], ye[N];
  T rr = T(0), delta = T(1);
  bool first = true, terminate = false, active = false;
  int result = RES_OK, steps = 0, naccept = 0, flip = 0;  // flip: which residual row holds the accepted point
  int64_t prob = warp0 * 2 + half;
  bool fresh = true;  // the slot has no problem loaded yet

  for (;;) {
    // ---------------- slots that are done write their result and take the next problem
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int src = s ? src1 : src0;
      const bool finished = active && (terminate || result != RES_OK || steps >= p.max_steps);
      const bool need = __shfl_sync(kFull, (finished || fresh) ? 1 : 0, src) != 0;  // warp-uniform
      if (!need) continue;
      if (half == s && finished) {
        if (result == RES_OK && !terminate) result = RES_NL_MAX_STEPS;  // _iterate.py:253-257
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (lane == src + j) p.y_out[prob * N + j] = y[j];
        if (lane == src) {
          p.num_steps[prob] = steps;
          p.result[prob] = result;
          if (p.step_size_out) p.step_size_out[prob] = delta;
          if (p.num_accepted) p.num_accepted[prob] = naccept;
        }
      }
      // next problem of this slot: its static first one, then tickets from the queue (or the static stride without a queue)
      const bool was_fresh = __shfl_sync(kFull, fresh ? 1 : 0, src) != 0;
      int64_t next;
      if (was_fresh) {
        next = warp0 * 2 + s;
      } else if (p.queue != nullptr) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(p.queue, 1ull);
        t = __shfl_sync(kFull, t, 0);
        next = nworkers + (int64_t)t;
      } else {
        next = __shfl_sync(kFull, prob, src) + nworkers;
      }
      const bool have = next < p.batch;
      T* ds = slots + s * slot_elems;
      if (have) {  // data row -> shared memory (all lanes), accepted residuals and the slot's (G, g) zeroed
        if (p.data_kind != B200OPT_DATA_AS_COMPUTE) {
          load_row16<T>(p.prob, next, p.prob_stride, m, p.data_kind, p.data_vec != 0, ds, lane);
        } else {
          const T* drow = p.prob + next * p.prob_stride;
          for (int i = lane; i < m / 4; i += 32) reinterpret_cast<float4*>(ds)[i] = __ldg(reinterpret_cast<const float4*>(drow) + i);
        }
        for (int i = lane; i < mp; i += 32) ds[mp + i] = T(0);  // residual row 0 = the accepted residuals of a fresh problem
        for (int i = lane; i < KPAD; i += 32) totals[s * KPAD + i] = T(0);
      }
      if (half == s) {
        prob = next;
        active = have;
        fresh = false;
        if (have) {
          bool yfin = true;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            y[j] = ye[j] = __ldg(p.y0 + next * N + j);
            yfin = yfin && t_finite(y[j]);
          }
          rr = T(0), delta = T(1);
          first = true, terminate = false;
          result = yfin ? RES_OK : RES_NONFINITE;  // _iterate.py:214-219
          steps = 0, naccept = 0, flip = 0;
        }
      }
      __syncwarp();
    }
    // a slot that starts non-finite (or with max_steps == 0) is finished before its first step: handled at the next loop top
    const bool running = active && !(terminate || result != RES_OK || steps >= p.max_steps);
    const unsigned run_mask = __ballot_sync(kFull, running);
    const unsigned act_mask = __ballot_sync(kFull, active);
    if (act_mask == 0u) break;      // both slots have run out of problems
    if (run_mask == 0u) continue;   // nothing to step: finished slots are retired at the loop top

    bool y_old_finite = true;
#pragma unroll
    for (int j = 0; j < N; ++j) y_old_finite = y_old_finite && t_finite(y[j]);

    // ---------------- residual at the trial point of each running slot, two rows per instruction (gauss_newton.py:253)
    T rre = T(0);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int src = s ? src1 : src0;
      if (!((run_mask >> src) & 1u)) continue;  // warp-uniform
      T yes[N];
#pragma unroll
      for (int j = 0; j < N; ++j) yes[j] = __shfl_sync(kFull, ye[j], src);
      const int fl = __shfl_sync(kFull, flip, src);
      const typename PModel::Consts pc = PModel::prepare(yes);
      T* ds = slots + s * slot_elems;
      const float2* ds2 = reinterpret_cast<const float2*>(ds);
      float2* rtry2 = reinterpret_cast<float2*>(ds + mp + (fl ? 0 : mp));  // the row that does NOT hold the accepted residuals
      float2* auxs2 = reinterpret_cast<float2*>(ds + 3 * mp);
      float2 part = make_float2(0.f, 0.f);
      auto res_pair = [&](int k) {
        const int idx = k * 32 + lane;
        float2 aux[NAUX];
        const float2 r = PModel::residual(pc, xs2[idx], ds2[idx], aux);
        rtry2[idx] = r;
#pragma unroll
        for (int q = 0; q < NAUX; ++q) auxs2[q * (mp >> 1) + idx] = aux[q];
        part = fma2(r, r, part);
      };
      if (NPAIRS > 0) {
#pragma unroll
        for (int k = 0; k < (NPAIRS > 0 ? NPAIRS : 1); ++k) res_pair(k);
      } else {
#pragma unroll kUnrollR
        for (int k = 0; k < npairs; ++k) res_pair(k);
      }
      const T rs = warp_sum(part.x + part.y);
      if (half == s) rre = rs;
    }

    // ---------------- search, both slots at once in the two half-warps (trust_region.py:86-104 / learning_rate.py:24-35)
    T d[N];
#pragma unroll
    for (int j = 0; j < N; ++j) d[j] = ye[j] - y[j];
    bool accept = true;
    if (TRUST) {
      T G[NG], g[N];
#pragma unroll
      for (int q = 0; q < RL::KQ; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(mytot + q * 4);
        const float* ww = reinterpret_cast<const float*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int a = q * 4 + e;
          if (a < NG) G[a] = ww[e];
          else if (a < K) g[a - NG] = ww[e];
        }
      }
      T lin = T(0), hq = T(0);  // 0.5 d.G d = sum_i d_i (0.5 G_ii d_i + sum_{j<i} G_ij d_j)
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T gd = T(0.5) * G[Tri<N>::at(i, i)] * d[i];
#pragma unroll
        for (int j = 0; j < i; ++j) gd += G[Tri<N>::at(i, j)] * d[j];
        hq += d[i] * gd;
        lin += d[i] * g[i];
      }
      const T quad = hq + hq;
      const T pred = predicted_reduction(lin, quad, rr);
      const T df = T(0.5) * rre - T(0.5) * rr;
      accept = (df <= T(0.01) * pred) || first;
      const bool good = (df < T(0.99) * pred) && (pred < T(0)) && !first;
      if (running) delta *= accept ? (good ? T(3.5) : T(1)) : T(0.25);
    } else {
      delta = T(1);
    }
    accept = accept && running;
    const unsigned acc_mask = __ballot_sync(kFull, accept);
    bool yconv = true;
#pragma unroll
    for (int j = 0; j < N; ++j) yconv = yconv && (t_abs(d[j]) < atol + rtol * t_abs(ye[j]));
    if (accept) {
      ++naccept;
#pragma unroll
      for (int j = 0; j < N; ++j) y[j] = ye[j];
      rr = rre;
    } else if (running) {
      terminate = false;
    }

    // ---------------- accepted slots: Jacobian rows -> Gram / gradient partials -> totals in shared memory (gauss_newton.py:276-296)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int src = s ? src1 : src0;
      if (!((acc_mask >> src) & 1u)) continue;  // warp-uniform
      T yes[N];
#pragma unroll
      for (int j = 0; j < N; ++j) yes[j] = __shfl_sync(kFull, y[j], src);  // == the trial point that was just accepted
      const int fl = __shfl_sync(kFull, flip, src);
      const typename PModel::Consts pc = PModel::prepare(yes);
      T* ds = slots + s * slot_elems;
      const float2* racc2 = reinterpret_cast<const float2*>(ds + mp + (fl ? mp : 0));
      const float2* rtry2 = reinterpret_cast<const float2*>(ds + mp + (fl ? 0 : mp));
      const float2* auxs2 = reinterpret_cast<const float2*>(ds + 3 * mp);
      T acc[K];
      bool fconv = true;
      constexpr int SPLIT = LSQ_P2_SPLIT;
#pragma unroll
      for (int hs = 0; hs < (SPLIT > 0 ? 2 : 1); ++hs) {
        const int a_lo = (hs == 0) ? 0 : SPLIT, a_hi = (hs == 0 && SPLIT > 0) ? SPLIT : N;
        float2 acc2[K];
        if (NPAIRS == 0) {
#pragma unroll
          for (int a = 0; a < K; ++a) acc2[a] = make_float2(0.f, 0.f);
        }
        auto jac_pair = [&](int k, bool init) {
          const int idx = k * 32 + lane;
          float2 aux[NAUX], j[N];
#pragma unroll
          for (int q = 0; q < NAUX; ++q) aux[q] = auxs2[q * (mp >> 1) + idx];
          const float2 re = rtry2[idx];
          PModel::jac(pc, xs2[idx], aux, j);
#pragma unroll
          for (int a = 0; a < N; ++a) {
            if (a >= a_lo && a < a_hi) {
#pragma unroll
              for (int b = 0; b <= a; ++b)
                acc2[Tri<N>::at(a, b)] = init ? mul2(j[a], j[b]) : fma2(j[a], j[b], acc2[Tri<N>::at(a, b)]);
              acc2[NG + a] = init ? mul2(j[a], re) : fma2(j[a], re, acc2[NG + a]);
            }
          }
          if (hs == 0) {
            const float2 ro = racc2[idx];
            fconv = fconv && (t_abs(re.x - ro.x) < atol + rtol * t_abs(re.x)) && (t_abs(re.y - ro.y) < atol + rtol * t_abs(re.y));
          }
        };
        if (NPAIRS > 0) {
#pragma unroll
          for (int k = 0; k < (NPAIRS > 0 ? NPAIRS : 1); ++k) jac_pair(k, k == 0);
        } else {
#pragma unroll kUnrollJ
          for (int k = 0; k < npairs; ++k) jac_pair(k, false);
        }
#pragma unroll
        for (int a = 0; a < N; ++a) {
          if (a >= a_lo && a < a_hi) {
#pragma unroll
            for (int b = 0; b <= a; ++b) acc[Tri<N>::at(a, b)] = acc2[Tri<N>::at(a, b)].x + acc2[Tri<N>::at(a, b)].y;
            acc[NG + a] = acc2[NG + a].x + acc2[NG + a].y;
          }
        }
      }
      warp_reduce_to_smem<K>(acc, scratch, totals + s * KPAD, lane);
      fconv = __all_sync(kFull, fconv);
      if (half == s) {
        terminate = yconv && fconv;  // _misc.py:310
        flip ^= 1;                   // the trial residuals become the accepted ones
      }
    }

    // ---------------- descent, both slots at once (levenberg_marquardt.py:55-57,66-75,132 / gauss_newton.py:64-78,126-127)
    {
      T G[NG], g[N];
#pragma unroll
      for (int q = 0; q < RL::KQ; ++q) {
        const float4 w = *reinterpret_cast<const float4*>(mytot + q * 4);
        const float* ww = reinterpret_cast<const float*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int a = q * 4 + e;
          if (a < NG) G[a] = ww[e];
          else if (a < K) g[a - NG] = ww[e];
        }
      }
      T lam = T(0);
      if (TRUST) lam = (delta > Num<T>::eps()) ? rcp_approx(delta) : Num<T>::max();
      T pstep[N], zz;
      spd_solve_fast<N>(G, lam, g, pstep, zz);
      if (running) {
        bool pfin = true;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          pfin = pfin && t_finite(pstep[j]);
          ye[j] = y[j] - pstep[j];
        }
        result = pfin ? RES_OK : RES_SINGULAR;          // lineax: non-finite solution => singular
        if (!y_old_finite) result = RES_NONFINITE;      // _iterate.py:216
        first = false;
        ++steps;
      }
    }
  }
}

}  // namespace b200opt
