// Register-lean variant of the packed-fp32 LM / GN kernel (lsq_kernel_p2.cuh): same state machine, same floating-point
// expressions in the same order - results are bit-identical - arranged so that a thread needs <= 96 registers without
// spills and FIVE CTAs (20 problems) are resident per SM instead of four.
//
// The one-warp kernel is bound by (problems resident per SM) / (latency of one step), not by instruction issue
// (cooperative-algebra experiment), and its three phases each peak above 100 registers:
//   residual sweep   model constants + y, y_eval + the replicated G (36) and g (8) that merely wait for the next phase
//   Gram sweep       24 pair accumulators (48 registers) + the Jacobian row pairs + the totals of the previous half-sweep
//   algebra          G (36) and its Cholesky factor (36) side by side
// Here
//   * G and g live in SHARED MEMORY between reductions: the warp all-reduce leaves its 44 totals in the `tot` line of the
//     scratch (where it used to read them back from) and the algebra loads them when it needs them - once for the
//     predicted reduction, once for the solve, where the loaded values die into the Cholesky factor as they are consumed;
//   * the Gram sweep runs in three parts (parameter rows [0,4), [4,6), [6,8): 14 / 13 / 17 pair accumulators) and every
//     part stores its per-lane partial sums straight into the lane's row of the reduction scratch.
// Used for m = 64 * NPAIRS (compile time); selected by B200OPT_LSQ_LEAN=1 until it has been measured (lsq_gauss2.cu).
#pragma once
#include "lsq_kernel_p2.cuh"

namespace b200opt {

// second half of warp_allreduce_array (lsq_kernel.cuh): the lanes' rows are already in the scratch; column sums in the
// same fixed order; the totals stay in `tot` (sm + 32 * KP), nothing is read back.
template <int K> __device__ __forceinline__ void warp_reduce_rows_keep(float* sm, int lane) {
  using L = WarpReduceLayout<float, K>;
  static_assert(L::KQ <= 16, "one pass of 16 quad columns");
  __syncwarp();
  float* tot = sm + 32 * L::KP;
  const int q = lane & 15, h = lane >> 4;
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
      a[e] = (h == 0) ? a[e] + o : o + a[e];
    }
    if (h == 0) *reinterpret_cast<float4*>(tot + q * 4) = make_float4(a[0], a[1], a[2], a[3]);
  }
  __syncwarp();
}

// which of the three Gram parts (parameter rows [0,4), [4,6), [6,N)) produces total number idx of [G (packed lower
// triangle) | g], and which part completes quad q of a scratch row
template <int N> __host__ __device__ constexpr int lean_index_part(int idx) {
  int a = idx - Tri<N>::size;  // gradient entry a, or ...
  if (idx < Tri<N>::size) {    // ... row a of the packed triangle
    a = 0;
    while ((a + 1) * (a + 2) / 2 <= idx) ++a;
  }
  return a < 4 ? 0 : (a < 6 ? 1 : 2);
}
template <int N> __host__ __device__ constexpr int lean_quad_part(int q) {
  int m = 0;
  for (int e = 0; e < 4; ++e) {
    const int pi = lean_index_part<N>(4 * q + e);
    m = pi > m ? pi : m;
  }
  return m;
}

#ifndef LSQ_P2L_MINBLOCKS
#define LSQ_P2L_MINBLOCKS 5
#endif
template <class PModel, bool TRUST, int NPAIRS>
__global__ void __launch_bounds__(kLsqWarpsPerBlock * 32, LSQ_P2L_MINBLOCKS) lsq_solve_kernel_p2l(const LsqParams<float> p) {
  using T = float;
  constexpr int N = PModel::N;
  constexpr int NG = Tri<N>::size;
  constexpr int NAUX = PModel::NAUX;
  constexpr int K = NG + N;
  static_assert(NPAIRS > 0 && K % 4 == 0, "compile-time pair count; whole quads of totals");
  using RL = WarpReduceLayout<T, K>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int mp = p.m_pad;  // == m == 64 * NPAIRS
  constexpr int SCR = RL::elems;
  T* xs = reinterpret_cast<T*>(smem_raw);
  T* wbase = xs + mp + (size_t)wib * (SCR + (size_t)(3 + NAUX) * mp);
  T* scratch = wbase;
  T* tot = scratch + 32 * RL::KP;  // [G (36) | g (8)] of the last accepted point, resident between reductions
  T* myrow = scratch + lane * RL::KP;
  T* ds = wbase + SCR;
  T* racc = ds + mp;
  T* rtry = racc + mp;
  T* auxs = rtry + mp;
  const int64_t warp0 = (int64_t)blockIdx.x * kLsqWarpsPerBlock + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kLsqWarpsPerBlock;
  const T rtol = p.rtol, atol = p.atol;

  for (int i = threadIdx.x; i < mp; i += blockDim.x) xs[i] = p.shared[i];
  __syncthreads();
  const float2* xs2 = reinterpret_cast<const float2*>(xs);
  const float2* ds2 = reinterpret_cast<const float2*>(ds);

  // the 44 totals, as the replicated registers of the one-warp kernel held them (broadcast 128-bit loads)
  auto load_Gg = [&](T (&G)[NG], T (&g)[N]) {
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(tot + q * 4);
      const T ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int a = q * 4 + e;
        if (a < NG) G[a] = ww[e];
        else g[a - NG] = ww[e];
      }
    }
  };

  for (int64_t prob = warp0; prob < p.batch; prob += nwarps) {
    T y[N], ye[N];
#pragma unroll
    for (int j = 0; j < N; ++j) y[j] = ye[j] = __ldg(p.y0 + prob * N + j);
    const T* drow = p.prob + prob * p.prob_stride;
    for (int i = lane; i < mp / 4; i += 32) reinterpret_cast<float4*>(ds)[i] = __ldg(reinterpret_cast<const float4*>(drow) + i);
    for (int i = lane; i < mp; i += 32) racc[i] = T(0);
    for (int i = lane; i < RL::KPAD; i += 32) tot[i] = T(0);  // G = 0, g = 0
    __syncwarp();

    T rr = T(0);
    T delta = T(1);
    bool first = true, terminate = false;
    bool yfin = true;
#pragma unroll
    for (int j = 0; j < N; ++j) yfin = yfin && t_finite(y[j]);
    int result = yfin ? RES_OK : RES_NONFINITE;
    int steps = 0, naccept = 0;

    while (!terminate && result == RES_OK && steps < p.max_steps) {
      bool y_old_finite = true;
#pragma unroll
      for (int j = 0; j < N; ++j) y_old_finite = y_old_finite && t_finite(y[j]);

      // ---- residual at the trial point, two rows per instruction (gauss_newton.py:253)
      float2* rtry2 = reinterpret_cast<float2*>(rtry);
      float2* auxs2 = reinterpret_cast<float2*>(auxs);
      T rre;
      {
        const typename PModel::Consts pc = PModel::prepare(ye);
        float2 part = make_float2(0.f, 0.f);
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
        rre = warp_sum(part.x + part.y);
      }

      T d[N];
#pragma unroll
      for (int j = 0; j < N; ++j) d[j] = ye[j] - y[j];

      // ---- search (trust_region.py:86-104 / learning_rate.py:24-35)
      bool accept = true;
      if (TRUST) {
        T G[NG], g[N];
        load_Gg(G, g);
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
        const T pred = lin + T(0.5) * quad;
        const T df = T(0.5) * rre - T(0.5) * rr;
        accept = (df <= T(0.01) * pred) || first;
        const bool good = (df < T(0.99) * pred) && (pred < T(0)) && !first;
        delta *= accept ? (good ? T(3.5) : T(1)) : T(0.25);
      } else {
        delta = T(1);
      }

      if (accept) {
        bool yconv = true;
#pragma unroll
        for (int j = 0; j < N; ++j) yconv = yconv && (t_abs(d[j]) < atol + rtol * t_abs(ye[j]));
        ++naccept;
#pragma unroll
        for (int j = 0; j < N; ++j) y[j] = ye[j];
        rr = rre;
        bool fconv = true;
        const float2* racc2 = reinterpret_cast<const float2*>(racc);
        const typename PModel::Consts pc = PModel::prepare(y);  // same values as at the residual sweep (y == y_eval)
        T acc[K];
        // Gram / gradient partials in three parts; each part leaves its per-lane sums in this lane's scratch row
#pragma unroll
        for (int part = 0; part < 3; ++part) {
          const int a_lo = part == 0 ? 0 : (part == 1 ? 4 : 6), a_hi = part == 0 ? 4 : (part == 1 ? 6 : N);
          float2 acc2[K];
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
            if (part == 0) {
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
          // 128-bit stores of every quad of the row that this part completes (conflict-free layout of lsq_kernel.cuh);
          // the few sums of a quad that straddles two parts wait in registers
#pragma unroll
          for (int q = 0; q < K / 4; ++q) {
            if (lean_quad_part<N>(q) == part)
              *reinterpret_cast<float4*>(myrow + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
          }
        }
        warp_reduce_rows_keep<K>(scratch, lane);
        fconv = __all_sync(kFull, fconv);
        terminate = yconv && fconv;
        T* t = racc;
        racc = rtry;
        rtry = t;
      } else {
        terminate = false;
      }

      // ---- descent (levenberg_marquardt.py:55-57,66-75,132 / gauss_newton.py:64-78,126-127)
      T lam = T(0);
      if (TRUST) lam = (delta > Num<T>::eps()) ? ((LSQ_P2_FAST & 1) ? rcp_approx(delta) : T(1) / delta) : Num<T>::max();
      T pstep[N];
      {
        T G[NG], g[N];
        load_Gg(G, g);
        if (LSQ_P2_FAST & 1) {
          T zz;
          spd_solve_fast<N>(G, lam, g, pstep, zz);
        } else {
          spd_solve<T, N>(G, lam, g, pstep);
        }
      }
      bool pfin = true;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        pfin = pfin && t_finite(pstep[j]);
        ye[j] = y[j] - pstep[j];
      }
      result = pfin ? RES_OK : RES_SINGULAR;
      if (!y_old_finite) result = RES_NONFINITE;
      first = false;
      ++steps;
    }
    if (result == RES_OK && !terminate) result = RES_NL_MAX_STEPS;

#pragma unroll
    for (int j = 0; j < N; ++j)
      if (lane == j) p.y_out[prob * N + j] = y[j];
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
