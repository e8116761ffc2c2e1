// Cooperative-algebra variant of the packed-fp32 LM / GN kernel (lsq_kernel_p2.cuh) for sm_100a.
//
// In lsq_solve_kernel_p2 every lane of a problem's warp executes the 8 x 8 algebra of a step (trust-region update, damped
// Cholesky solve, predicted reduction) redundantly: ~450 of the 878 warp-instructions of a problem-iteration buy one
// lane's worth of work.  Here a CTA of W warps iterates W problems in lock step and the algebra of all W problems runs in
// W LANES of one warp:
//
//   phase A (every warp, its own problem)   residual sweep at the trial point, accept / reject, and for an accepted point
//                                           the Jacobian / Gram sweeps; the warp reduction leaves G = J^T J and g = J^T r
//                                           in the problem's shared-memory slot
//   barrier
//   phase S (warp 0, lane l <-> problem l)  Delta update, lambda = 1 / Delta, Cholesky solve of (G + lambda I) p = g,
//                                           y_eval = y - p, predicted reduction of the NEXT trial point; results -> slot
//   barrier
//
// The reference's step (gauss_newton.py:243-334) is exactly one A and one S per iteration - the predicted reduction that
// trust_region.py:86-104 evaluates at the start of the next step depends only on (G, g, y, y_eval), all known at the end
// of S - so the state machine, every floating-point expression and therefore every result is that of
// lsq_solve_kernel_p2, bit for bit (tests/test_lm_gpu.py::test_gauss2_cooperative_kernel_is_bit_identical).  Two or more
// CTAs per SM overlap one CTA's phase S (a latency chain on one warp) with the others' sweeps.  A warp whose problem
// ends takes the next one of its static sequence and rejoins the rounds; the CTA leaves when all sequences are exhausted.
//
// Used for fp32 gauss2-like packed models with m = 64 * NPAIRS; selected by B200OPT_LSQ_COOP=<W> (lsq_gauss2.cu).
#pragma once
#include "lsq_kernel_p2.cuh"

namespace b200opt {

// per-problem slot in shared memory (floats); the odd stride makes "lane l reads word a of slot l" conflict free
struct CoopSlot {
  static constexpr int G = 0;        // 36
  static constexpr int g = 36;       // 8
  static constexpr int y = 44;       // 8   last accepted point (posted by the problem's warp)
  static constexpr int ye = 52;      // 8   trial point (posted by the algebra lane)
  static constexpr int delta = 60;
  static constexpr int pred = 61;
  static constexpr int flags = 62;   // int: bit 0 active, bit 1 accepted, bit 2 good
  static constexpr int pfin = 63;    // int: the solve produced a finite step
  static constexpr int stride = 65;
};

template <int W> constexpr size_t lsq_coop_smem_bytes(int m_pad, int n, int naux) {
  return sizeof(float) * ((size_t)m_pad + (size_t)W * CoopSlot::stride +
                          (size_t)W * (WarpReduceLayout<float, 44>::elems + (size_t)(3 + naux) * m_pad));
}

template <class PModel, bool TRUST, int NPAIRS, int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB) lsq_solve_kernel_p2c(const LsqParams<float> p) {
  using T = float;
  constexpr int N = PModel::N;
  constexpr int NG = Tri<N>::size;
  constexpr int NAUX = PModel::NAUX;
  static_assert(N == 8 && NG + N == 44, "slot layout is written for n = 8");
  static_assert(W <= 32, "one algebra lane per problem");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int mp = p.m_pad;  // == m == 64 * NPAIRS
  constexpr int SCR = WarpReduceLayout<T, NG + N>::elems;
  T* xs = reinterpret_cast<T*>(smem_raw);
  T* slots = xs + mp;
  T* wbase = slots + W * CoopSlot::stride + (size_t)wib * (SCR + (size_t)(3 + NAUX) * mp);
  T* scratch = wbase;
  T* ds = wbase + SCR;
  T* racc = ds + mp;
  T* rtry = racc + mp;
  T* auxs = rtry + mp;
  T* slot = slots + wib * CoopSlot::stride;
  int* sloti = reinterpret_cast<int*>(slot);
  const T rtol = p.rtol, atol = p.atol;
  const int64_t stride_prob = (int64_t)gridDim.x * W;
  int64_t next_prob = (int64_t)blockIdx.x * W + wib;

  for (int i = threadIdx.x; i < mp; i += blockDim.x) xs[i] = p.shared[i];
  __syncthreads();
  const float2* xs2 = reinterpret_cast<const float2*>(xs);
  const float2* ds2 = reinterpret_cast<const float2*>(ds);

  // ---- per-problem state of this warp (replicated in its lanes)
  T y[N], ye[N];
  T rr = T(0), pred = T(0);
  bool first = true, terminate = false, have = false, y_old_finite = true;
  int result = RES_OK, steps = 0, naccept = 0;
  int64_t prob = 0;

  auto write_out = [&]() {
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (lane == j) p.y_out[prob * N + j] = y[j];
    if (lane == 0) {
      p.num_steps[prob] = steps;
      p.result[prob] = result;
      if (p.step_size_out) p.step_size_out[prob] = steps > 0 ? slot[CoopSlot::delta] : T(1);
      if (p.num_accepted) p.num_accepted[prob] = naccept;
    }
  };
  // take the next runnable problem of this warp's sequence (problems that cannot start are finished on the spot)
  auto load_next = [&]() {
    have = false;
    while (next_prob < p.batch) {
      prob = next_prob;
      next_prob += stride_prob;
#pragma unroll
      for (int j = 0; j < N; ++j) y[j] = ye[j] = __ldg(p.y0 + prob * N + j);
      bool yfin = true;
#pragma unroll
      for (int j = 0; j < N; ++j) yfin = yfin && t_finite(y[j]);
      result = yfin ? RES_OK : RES_NONFINITE;
      steps = 0, naccept = 0;
      rr = T(0), pred = T(0);
      first = true, terminate = false;
      if (result != RES_OK || p.max_steps <= 0) {  // the while loop of the sequential kernel would not run
        if (result == RES_OK) result = RES_NL_MAX_STEPS;
        write_out();
        continue;
      }
      const T* drow = p.prob + prob * p.prob_stride;
      for (int i = lane; i < mp / 4; i += 32) reinterpret_cast<float4*>(ds)[i] = __ldg(reinterpret_cast<const float4*>(drow) + i);
      for (int i = lane; i < mp; i += 32) racc[i] = T(0);
      if (lane == 0) slot[CoopSlot::delta] = T(1);
      __syncwarp();
      have = true;
      return;
    }
  };
  load_next();

  for (;;) {
    // =================================================================== phase A: sweeps of this warp's problem
    if (have) {
      y_old_finite = true;
#pragma unroll
      for (int j = 0; j < N; ++j) y_old_finite = y_old_finite && t_finite(y[j]);

      // ---- residual at the trial point, two rows per instruction (gauss_newton.py:253)
      const typename PModel::Consts pc = PModel::prepare(ye);
      float2 part = make_float2(0.f, 0.f);
      float2* rtry2 = reinterpret_cast<float2*>(rtry);
      float2* auxs2 = reinterpret_cast<float2*>(auxs);
#pragma unroll
      for (int k = 0; k < NPAIRS; ++k) {
        const int idx = k * 32 + lane;
        float2 aux[NAUX];
        const float2 r = PModel::residual(pc, xs2[idx], ds2[idx], aux);
        rtry2[idx] = r;
#pragma unroll
        for (int q = 0; q < NAUX; ++q) auxs2[q * (mp >> 1) + idx] = aux[q];
        part = fma2(r, r, part);
      }
      const T rre = warp_sum(part.x + part.y);

      T d[N];
#pragma unroll
      for (int j = 0; j < N; ++j) d[j] = ye[j] - y[j];

      // ---- search (trust_region.py:86-104 / learning_rate.py:24-35); `pred` was evaluated by the algebra lane
      bool accept = true, good = false;
      if (TRUST) {
        const T df = T(0.5) * rre - T(0.5) * rr;
        accept = (df <= T(0.01) * pred) || first;
        good = (df < T(0.99) * pred) && (pred < T(0)) && !first;
      }

      if (accept) {
        bool yconv = true;
#pragma unroll
        for (int j = 0; j < N; ++j) yconv = yconv && (t_abs(d[j]) < atol + rtol * t_abs(ye[j]));
        ++naccept;
#pragma unroll
        for (int j = 0; j < N; ++j) y[j] = ye[j];
        rr = rre;
        T acc[NG + N];
        bool fconv = true;
        const float2* racc2 = reinterpret_cast<const float2*>(racc);
        constexpr int SPLIT = LSQ_P2_SPLIT > 0 ? LSQ_P2_SPLIT : 5;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int a_lo = (half == 0) ? 0 : SPLIT, a_hi = (half == 0) ? SPLIT : N;
          float2 acc2[NG + N];
#pragma unroll
          for (int k = 0; k < NPAIRS; ++k) {
            const bool init = k == 0;
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
            if (half == 0) {
              const float2 ro = racc2[idx];
              fconv = fconv && (t_abs(re.x - ro.x) < atol + rtol * t_abs(re.x)) && (t_abs(re.y - ro.y) < atol + rtol * t_abs(re.y));
            }
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
        // same reduction (same order of additions) as the one-warp kernel; every lane ends with the 44 totals, lanes
        // 0..10 leave them in the slot for the algebra lane: [G (36) | g (8)]
        warp_allreduce_array<T, NG + N>(acc, scratch, lane);
#pragma unroll
        for (int q = 0; q < (NG + N) / 4; ++q) {
          if (lane == q) {
#pragma unroll
            for (int e = 0; e < 4; ++e) slot[4 * q + e] = acc[4 * q + e];
          }
        }
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (lane == 16 + j) slot[CoopSlot::y + j] = y[j];
        fconv = __all_sync(kFull, fconv);
        terminate = yconv && fconv;
        T* t = racc;
        racc = rtry;
        rtry = t;
      } else {
        terminate = false;
      }
      if (lane == 0) sloti[CoopSlot::flags] = 1 | (accept ? 2 : 0) | (good ? 4 : 0);
    } else {
      if (lane == 0) sloti[CoopSlot::flags] = 0;
    }
    if (!__syncthreads_or(have ? 1 : 0)) break;  // barrier 1: every sequence of the CTA is exhausted

    // =================================================================== phase S: algebra, lane l <-> problem l
    if (wib == 0 && lane < W) {
      T* s = slots + lane * CoopSlot::stride;
      int* si = reinterpret_cast<int*>(s);
      const int fl = si[CoopSlot::flags];
      if (fl & 1) {
        T G[NG], g[N], ya[N];
#pragma unroll
        for (int a = 0; a < NG; ++a) G[a] = s[CoopSlot::G + a];
#pragma unroll
        for (int j = 0; j < N; ++j) g[j] = s[CoopSlot::g + j], ya[j] = s[CoopSlot::y + j];
        T delta = s[CoopSlot::delta];
        // trust_region.py:95-104 (the first step: accept, not good -> factor 1) / learning_rate.py:24-35
        if (TRUST) delta *= (fl & 2) ? ((fl & 4) ? T(3.5) : T(1)) : T(0.25);
        else delta = T(1);
        s[CoopSlot::delta] = delta;
        // ---- descent (levenberg_marquardt.py:55-57,66-75,132 / gauss_newton.py:64-78,126-127)
        T lam = T(0);
        if (TRUST) lam = (delta > Num<T>::eps()) ? ((LSQ_P2_FAST & 1) ? rcp_approx(delta) : T(1) / delta) : Num<T>::max();
        T pstep[N];
        if (LSQ_P2_FAST & 1) {
          T zz;
          spd_solve_fast<N>(G, lam, g, pstep, zz);
        } else {
          spd_solve<T, N>(G, lam, g, pstep);
        }
        bool pfin = true;
        T yn[N], d[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
          pfin = pfin && t_finite(pstep[j]);
          yn[j] = ya[j] - pstep[j];
          s[CoopSlot::ye + j] = yn[j];
          d[j] = yn[j] - ya[j];
        }
        si[CoopSlot::pfin] = pfin ? 1 : 0;
        // ---- predicted reduction of the next trial point, d.g + 0.5 d.G d (trust_region.py:182-186), same expression
        //      as in lsq_solve_kernel_p2
        if (TRUST) {
          T lin = T(0), quad = T(0);
          if (LSQ_P2_FAST & 2) {
            T hq = T(0);
#pragma unroll
            for (int i = 0; i < N; ++i) {
              T gd = T(0.5) * G[Tri<N>::at(i, i)] * d[i];
#pragma unroll
              for (int j = 0; j < i; ++j) gd += G[Tri<N>::at(i, j)] * d[j];
              hq += d[i] * gd;
              lin += d[i] * g[i];
            }
            quad = hq + hq;
          } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
              T gd = T(0);
#pragma unroll
              for (int j = 0; j < N; ++j) gd += G[Tri<N>::at(i, j)] * d[j];
              quad += d[i] * gd;
              lin += d[i] * g[i];
            }
          }
          s[CoopSlot::pred] = lin + T(0.5) * quad;
        }
      }
    }
    __syncthreads();  // barrier 2: trial points are posted

    if (have) {
#pragma unroll
      for (int j = 0; j < N; ++j) ye[j] = slot[CoopSlot::ye + j];
      pred = slot[CoopSlot::pred];
      result = sloti[CoopSlot::pfin] ? RES_OK : RES_SINGULAR;
      if (!y_old_finite) result = RES_NONFINITE;
      first = false;
      ++steps;
      if (terminate || result != RES_OK || steps >= p.max_steps) {
        if (result == RES_OK && !terminate) result = RES_NL_MAX_STEPS;
        write_out();
        load_next();
      }
    }
  }
}

}  // namespace b200opt
