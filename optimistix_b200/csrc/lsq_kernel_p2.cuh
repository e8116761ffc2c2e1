// Packed-fp32 variant of the batched LM / GN kernel (lsq_kernel.cuh) for sm_100a.
//
// Blackwell issues fp32 FMA / MUL / ADD on register PAIRS (FFMA2 / FMUL2 / FADD2: two lanes of arithmetic per issue slot,
// `__ffma2_rn` & co.).  The scalar kernel is bound by issue slots, not by the FMA pipe (profiles/README.md: 1505
// instructions per problem-iteration, pipe 45% busy), so this variant evaluates TWO ROWS of the problem per instruction:
// lane l owns the row pairs (64 k + 2 l, 64 k + 2 l + 1), which are adjacent in memory, so every shared-memory access
// is one conflict-free 64-bit access, and the residual, the Jacobian row and the 44 Gram / gradient partial sums are
// all carried as (even row, odd row) pairs.  The state machine, the replicated 8 x 8 algebra and the results are those
// of lsq_kernel.cuh; only the order of the floating-point additions inside the row sums differs.
//
// exp(-z^2/2) = 2^(t K), K = -log2(e)/2, is evaluated as ex2.approx(hi) * (1 + ln2 * lo) with hi + lo = t K split exactly
// by one extra FMA (error <= 2 ulp; the hardware ex2 alone would lose |hi| ulp to the rounding of its argument).
//
// Used when: fp32, max_norm, no root-finding extra condition, m a multiple of 64, 16-byte aligned rows.
#pragma once
#include "lsq_kernel.cuh"

namespace b200opt {

__device__ __forceinline__ float2 p2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Tuning switches of the replicated 8 x 8 algebra (bit mask; A/B builds: tools/lm_variants.sh):
//   1  hardware rsqrt / reciprocal (MUFU, no denormal fix-up sequence) in the Cholesky solve and for lambda = 1 / Delta;
//      back-substitution sums the newest unknown LAST (dependent chain of 1 instead of N - 1 - i FMAs per row)
//   2  predicted reduction through the symmetric form  sum_i d_i (G_ii d_i + 2 sum_{j<i} G_ij d_j)  (36 instead of 64 FMAs)
//   4  (off) packed adds (FADD2) in the column sums of the warp all-reduce (same order of additions: bit-identical sums)
//   8  (off) predicted reduction from the Cholesky factor of the solve that produced d:  with L z = g, L^T p = z, d = -p:
//      d.g + 0.5 d.G d = -0.5 (|z|^2 + lambda |d|^2)  - 16 FMAs, but exact only for an exact solve
// Measured on 1 Mi gauss2 fits (ms per launch): 0: 15.63, 1: 15.12, 3: 14.55, 7: 14.62, 15: 14.32.  Bit 4 removes 32
// instructions from a part of the step that is latency bound (no gain); bit 8 is the fastest but it is NOT the reference's
// predicted reduction in floating point: mean iteration count 10.85 instead of 10.51 and 9x more `max_steps` results on
// the same inputs, so it stays an experiment.
#ifndef LSQ_P2_FAST
#define LSQ_P2_FAST 3
#endif
// Round-2 latency / work trims (bit mask; A/B builds: tools/kernel_variants.sh lsq_gauss2 tag:"-DLSQ_P2_OPT=...", timed and
// checked bit for bit against the baseline by tools/call_lm_ab.sh; results are bit-identical to 0; measurements: DESIGN.md 4.1b):
//   1  the five butterfly rounds of the ||r||^2 reduction are interleaved with the rows of the predicted-reduction quadratic form
//      (the reduction sits behind a divergence check = a basic-block boundary, so the assembler does not overlap them itself)
//   2  Gram entries between parameters whose Jacobian column depends neither on y nor on the problem (PModel::NCONST trailing
//      parameters: the baseline columns 1 and x) are summed ONCE per warp with the same code path and reused
//   4  the f-space Cauchy test (one comparison per row) runs only when the y-space test passed (terminate needs both)
//   8  the two Gram half-sweeps split the entries by parameter GROUP (PModel::kHalf0Mask) instead of by parameter index, so
//      that the first one needs only part of the Jacobian row
//  16  exp split: lo = fma(t, K, -hi) with the negation as an operand modifier instead of a second product t * (-K)
//  32  the first Gram half-sweep leaves j[1] in the (dead) slot of the old accepted residual for the second one
#ifndef LSQ_P2_OPT
#define LSQ_P2_OPT 63
#endif
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // x < 0 or NaN -> NaN, x == 0 -> +inf (reported as `singular`)
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (G + lam I) p = g as spd_solve (common.cuh), fp32 only: hardware rsqrt, newest unknown last in the back-substitution.
template <int N>
__device__ __forceinline__ void spd_solve_fast(const float (&G)[Tri<N>::size], float lam, const float (&g)[N], float (&p)[N], float& zz) {
  float L[Tri<N>::size];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    float s = G[Tri<N>::at(j, j)] + lam;
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[Tri<N>::at(j, k)] * L[Tri<N>::at(j, k)];
    const float inv = rsqrt_approx(s);
    L[Tri<N>::at(j, j)] = inv;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      float t = G[Tri<N>::at(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[Tri<N>::at(i, k)] * L[Tri<N>::at(j, k)];
      L[Tri<N>::at(i, j)] = t * inv;
    }
  }
  float z[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[Tri<N>::at(i, k)] * z[k];
    z[i] = s * L[Tri<N>::at(i, i)];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    float s = z[i];
#pragma unroll
    for (int k = N - 1; k > i; --k) s -= L[Tri<N>::at(k, i)] * p[k];
    p[i] = s * L[Tri<N>::at(i, i)];
  }
  zz = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) zz += z[i] * z[i];
}

// warp_allreduce_array (lsq_kernel.cuh) for fp32 with the 16-row column sums done two columns per instruction.
template <int K> __device__ __forceinline__ void warp_allreduce_array_f2(float (&v)[K], float* sm, int lane) {
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
  float* tot = sm + 32 * L::KP;
  static_assert(L::KQ <= 16, "one pass of 16 quad columns");
  {
    const int q = lane & 15, h = lane >> 4;
    const bool on = q < L::KQ;
    const unsigned mask = __ballot_sync(kFull, on);
    if (on) {
      float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
      for (int l = 0; l < 16; ++l) {
        const float4 w = *reinterpret_cast<const float4*>(sm + (h * 16 + l) * L::KP + q * 4);
        a01 = add2(a01, make_float2(w.x, w.y));
        a23 = add2(a23, make_float2(w.z, w.w));
      }
      float a[4] = {a01.x, a01.y, a23.x, a23.y};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float o = __shfl_xor_sync(mask, a[e], 16);
        a[e] = (h == 0) ? a[e] + o : o + a[e];
      }
      if (h == 0) *reinterpret_cast<float4*>(tot + q * 4) = make_float4(a[0], a[1], a[2], a[3]);
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < L::KQ; ++q) {
    const float4 w = *reinterpret_cast<const float4*>(tot + q * 4);
    const float* ww = reinterpret_cast<const float*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (q * 4 + e < K) v[q * 4 + e] = ww[e];
  }
  __syncwarp();
}

// exp(-0.5 * t) for a pair of non-negative t
__device__ __forceinline__ float2 exp_mhalf2(float2 t) {
  const float K = -0.72134752044448170368f;  // -log2(e) / 2
  const float2 hi = mul2(t, p2(K));
#if LSQ_P2_OPT & 16
  const float2 lo = fma2(t, p2(K), make_float2(-hi.x, -hi.y));  // t K - hi, exact (the negation is an operand modifier)
#else
  const float2 nhi = mul2(t, p2(-K));           // == -hi exactly
  const float2 lo = fma2(t, p2(K), nhi);        // t K - hi, exact
#endif
  const float2 e0 = make_float2(ex2_approx(hi.x), ex2_approx(hi.y));
  return fma2(mul2(e0, lo), p2(0.69314718055994530942f), e0);
}

// model family id 1 (two Gaussians + linear baseline), two rows at a time.  Same parameter layout as Gauss2<float>.
struct Gauss2P2 {
  static constexpr int N = 8;
  static constexpr int NAUX = 2;
  // LSQ_P2_OPT & 8: the first Gram half-sweep takes the entries among the first Gaussian's parameters and the baseline (it never
  // forms the second Gaussian's Jacobian columns: 4 packed instructions per row pair less), the second one everything else
  static constexpr unsigned kHalf0Mask = 0xC7;
  static constexpr int NCONST = 2;  // trailing parameters (b0, b1) whose Jacobian columns (1, x) are the same for every y and problem
  struct Consts {
    float2 i1, c1, i2, c2;    // z = x * i + c,  c = -mu * i
    float2 a1, a2, b0, b1;    // amplitudes and baseline (residual)
    float2 ai1, ai2;          // A * i (Jacobian)
  };
  __device__ static __forceinline__ Consts prepare(const float (&y)[N]) {
    // 1 / sigma by the hardware reciprocal (MUFU.RCP, 1 ulp) instead of the 13-instruction IEEE division: sigma enters
    // only through z = (x - mu) / sigma, whose own rounding is of the same size
    float i1, i2;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i1) : "f"(y[2]));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(i2) : "f"(y[5]));
    Consts c;
    c.i1 = p2(i1), c.c1 = p2(-y[1] * i1), c.i2 = p2(i2), c.c2 = p2(-y[4] * i2);
    c.a1 = p2(y[0]), c.a2 = p2(y[3]), c.b0 = p2(y[6]), c.b1 = p2(y[7]);
    c.ai1 = p2(y[0] * i1), c.ai2 = p2(y[3] * i2);
    return c;
  }
  __device__ static __forceinline__ Consts prepare_const() {  // any y: only the NCONST trailing Jacobian columns are read
    Consts c;
    c.i1 = c.c1 = c.i2 = c.c2 = c.a1 = c.a2 = c.b0 = c.b1 = c.ai1 = c.ai2 = p2(0.f);
    return c;
  }
  __device__ static __forceinline__ float2 residual(const Consts& c, float2 x, float2 d, float2 (&aux)[NAUX]) {
    const float2 z1 = fma2(x, c.i1, c.c1), z2 = fma2(x, c.i2, c.c2);
    const float2 e1 = exp_mhalf2(mul2(z1, z1)), e2 = exp_mhalf2(mul2(z2, z2));
    aux[0] = e1;
    aux[1] = e2;
    const float2 base = fma2(c.b1, x, add2(c.b0, make_float2(-d.x, -d.y)));
    return fma2(c.a1, e1, fma2(c.a2, e2, base));
  }
  __device__ static __forceinline__ void jac(const Consts& c, float2 x, const float2 (&aux)[NAUX], float2 (&j)[N]) {
    const float2 z1 = fma2(x, c.i1, c.c1), z2 = fma2(x, c.i2, c.c2);
    j[0] = aux[0];
    j[1] = mul2(mul2(aux[0], z1), c.ai1);
    j[2] = mul2(j[1], z1);
    j[3] = aux[1];
    j[4] = mul2(mul2(aux[1], z2), c.ai2);
    j[5] = mul2(j[4], z2);
    j[6] = p2(1.0f);
    j[7] = x;
  }
  // LSQ_P2_OPT & 32: the first Gram half-sweep hands j[1] to the second one (through the slot of the old accepted residual,
  // which is dead once the f-space test has read it), so that the second does not form it again
  __device__ static __forceinline__ float2 carry_out(const float2 (&j)[N]) { return j[1]; }
  __device__ static __forceinline__ void jac_carry(const Consts& c, float2 x, const float2 (&aux)[NAUX], float2 carry, float2 (&j)[N]) {
    const float2 z1 = fma2(x, c.i1, c.c1), z2 = fma2(x, c.i2, c.c2);
    j[0] = aux[0];
    j[1] = carry;
    j[2] = mul2(j[1], z1);
    j[3] = aux[1];
    j[4] = mul2(mul2(aux[1], z2), c.ai2);
    j[5] = mul2(j[4], z2);
    j[6] = p2(1.0f);
    j[7] = x;
  }
};

#ifndef LSQ_P2_MINBLOCKS
#define LSQ_P2_MINBLOCKS 4
#endif
#ifndef LSQ_P2_UNROLL_R
#define LSQ_P2_UNROLL_R 4
#endif
#ifndef LSQ_P2_UNROLL_J
#define LSQ_P2_UNROLL_J 4
#endif
#ifndef LSQ_P2_SPLIT
#define LSQ_P2_SPLIT 5  // 0: one sweep with all 44 pair accumulators (88 registers)
#endif
constexpr int kUnrollR = LSQ_P2_UNROLL_R, kUnrollJ = LSQ_P2_UNROLL_J;  // (macros are not expanded inside #pragma)
// NPAIRS > 0: the number of row pairs per lane (m / 64) is a compile-time constant: both row sweeps are fully unrolled and
// the accumulators are initialised by the first pair instead of being zeroed.  NPAIRS == 0: any m % 64 == 0.
template <class PModel> __host__ __device__ constexpr bool p2_const_entry(int a, int b) {
  return (LSQ_P2_OPT & 2) && a >= PModel::N - PModel::NCONST && b >= PModel::N - PModel::NCONST;
}
// which of the two Gram half-sweeps carries entry (a, b) of the lower triangle (b == N: the gradient entry of parameter a)
template <class PModel> __host__ __device__ constexpr int p2_half(int a, int b) {
  if (LSQ_P2_SPLIT == 0) return 0;
  if (LSQ_P2_OPT & 8) {
    const bool a0 = (PModel::kHalf0Mask >> a) & 1, b0 = (b == PModel::N) || ((PModel::kHalf0Mask >> b) & 1);
    return (a0 && b0) ? 0 : 1;
  }
  return a < LSQ_P2_SPLIT ? 0 : 1;
}
template <class PModel, bool TRUST, int NPAIRS>
__global__ void __launch_bounds__(kLsqWarpsPerBlock * 32, LSQ_P2_MINBLOCKS) lsq_solve_kernel_p2(const LsqParams<float> p) {
  using T = float;
  constexpr int N = PModel::N;
  constexpr int NG = Tri<N>::size;
  constexpr int NAUX = PModel::NAUX;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int m = p.m, mp = p.m_pad;  // m % 64 == 0, m_pad == m
  constexpr int SCR = WarpReduceLayout<T, NG + N>::elems;
  T* xs = reinterpret_cast<T*>(smem_raw);
  T* wbase = xs + mp + (size_t)wib * (SCR + (size_t)(3 + NAUX) * mp);
  T* scratch = wbase;
  T* ds = wbase + SCR;
  T* racc = ds + mp;
  T* rtry = racc + mp;
  T* auxs = rtry + mp;
  const int64_t warp0 = (int64_t)blockIdx.x * kLsqWarpsPerBlock + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kLsqWarpsPerBlock;
  const T rtol = p.rtol, atol = p.atol;
  const int npairs = NPAIRS > 0 ? NPAIRS : (m >> 6);  // row pairs per lane

  for (int i = threadIdx.x; i < mp; i += blockDim.x) xs[i] = p.shared[i];
  __syncthreads();
  const float2* xs2 = reinterpret_cast<const float2*>(xs);
  const float2* ds2 = reinterpret_cast<const float2*>(ds);

  // problem-independent Gram entries (LSQ_P2_OPT & 2): the same pair sums, pair -> scalar adds and warp all-reduce as below
  T gconst[(PModel::NCONST * (PModel::NCONST + 1)) / 2 + 1];
  if (LSQ_P2_OPT & 2) {
    T cacc[NG + N];
#pragma unroll
    for (int a = 0; a < NG + N; ++a) cacc[a] = T(0);
    const typename PModel::Consts pc0 = PModel::prepare_const();
    float2 c2[NG];
#pragma unroll
    for (int a = 0; a < NG; ++a) c2[a] = make_float2(0.f, 0.f);
    for (int k = 0; k < npairs; ++k) {
      const int idx = k * 32 + lane;
      float2 aux[NAUX], j[N];
#pragma unroll
      for (int q = 0; q < NAUX; ++q) aux[q] = make_float2(0.f, 0.f);
      PModel::jac(pc0, xs2[idx], aux, j);
#pragma unroll
      for (int a = N - PModel::NCONST; a < N; ++a)
#pragma unroll
        for (int b = N - PModel::NCONST; b <= a; ++b)
          c2[Tri<N>::at(a, b)] = (NPAIRS > 0 && k == 0) ? mul2(j[a], j[b]) : fma2(j[a], j[b], c2[Tri<N>::at(a, b)]);
    }
#pragma unroll
    for (int a = N - PModel::NCONST; a < N; ++a)
#pragma unroll
      for (int b = N - PModel::NCONST; b <= a; ++b) cacc[Tri<N>::at(a, b)] = c2[Tri<N>::at(a, b)].x + c2[Tri<N>::at(a, b)].y;
    warp_allreduce_array<T, NG + N>(cacc, scratch, lane);
    int q = 0;
#pragma unroll
    for (int a = N - PModel::NCONST; a < N; ++a)
#pragma unroll
      for (int b = N - PModel::NCONST; b <= a; ++b) gconst[q++] = cacc[Tri<N>::at(a, b)];
  }

  for (int64_t prob = warp0; prob < p.batch; prob = next_problem_warp(p.queue, prob, nwarps, lane)) {
    T y[N], ye[N];
#pragma unroll
    for (int j = 0; j < N; ++j) y[j] = ye[j] = __ldg(p.y0 + prob * N + j);
    if (p.data_kind != B200OPT_DATA_AS_COMPUTE) {
      load_row16<T>(p.prob, prob, p.prob_stride, m, p.data_kind, p.data_vec != 0, ds, lane);
    } else {
      const T* drow = p.prob + prob * p.prob_stride;
      for (int i = lane; i < m / 4; i += 32) reinterpret_cast<float4*>(ds)[i] = __ldg(reinterpret_cast<const float4*>(drow) + i);
    }
    for (int i = lane; i < mp; i += 32) racc[i] = T(0);
    __syncwarp();

    T G[NG], g[N];
#pragma unroll
    for (int a = 0; a < NG; ++a) G[a] = T(0);
#pragma unroll
    for (int j = 0; j < N; ++j) g[j] = T(0);
    T rr = T(0);
    T delta = T(1);
    T predn = T(0);  // LSQ_P2_FAST & 8 only
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
      const typename PModel::Consts pc = PModel::prepare(ye);
      float2 part = make_float2(0.f, 0.f);
      float2* rtry2 = reinterpret_cast<float2*>(rtry);
      float2* auxs2 = reinterpret_cast<float2*>(auxs);
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
      T d[N];
#pragma unroll
      for (int j = 0; j < N; ++j) d[j] = ye[j] - y[j];

      // ---- search (trust_region.py:86-104 / learning_rate.py:24-35)
      bool accept = true;
      T rre;
      T lin = T(0), quad = T(0);
      if (TRUST && (LSQ_P2_FAST & 2) && !(LSQ_P2_FAST & 8) && (LSQ_P2_OPT & 1) && !first) {  // first step: d = 0, G = 0, g = 0: lin = quad = 0
        // 0.5 d.G d = sum_i d_i (0.5 G_ii d_i + sum_{j<i} G_ij d_j), its rows issued between the rounds of the ||r||^2 butterfly
        T hq = T(0);
        auto quad_row = [&](int i) {
          T gd = T(0.5) * G[Tri<N>::at(i, i)] * d[i];
#pragma unroll
          for (int j = 0; j < N; ++j)
            if (j < i) gd += G[Tri<N>::at(i, j)] * d[j];
          hq += d[i] * gd;
          lin += d[i] * g[i];
        };
        T v = part.x + part.y;
        constexpr int kRowsAfterRound[5] = {(N * 3) / 8, (N * 5) / 8, (N * 6) / 8, (N * 7) / 8, N};  // N = 8: rows 0-2, 3-4, 5, 6, 7
        int i0 = 0;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          v += __shfl_xor_sync(kFull, v, 16 >> r);
#pragma unroll
          for (int i = 0; i < N; ++i)
            if (i >= i0 && i < kRowsAfterRound[r]) quad_row(i);
          i0 = kRowsAfterRound[r];
        }
        rre = v;
        quad = hq + hq;
      } else {
        rre = warp_sum(part.x + part.y);
      }
      if (TRUST) {
        if (LSQ_P2_FAST & 8) {
          lin = predn;
        } else if ((LSQ_P2_FAST & 2) && (LSQ_P2_OPT & 1)) {
          // done above
        } else if (LSQ_P2_FAST & 2) {
          T hq = T(0);  // 0.5 d.G d = sum_i d_i (0.5 G_ii d_i + sum_{j<i} G_ij d_j)
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
        const T pred = predicted_reduction(lin, quad, rr);
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
        // Gram / gradient partials as (even row, odd row) pairs.  The 44 pair accumulators would cost 88 registers, so
        // the rows are swept twice, each sweep carrying about half of them (parameters [0, SPLIT) then [SPLIT, N)): the
        // Jacobian row is recomputed (a handful of packed multiplies), occupancy goes from 12 to 16+ warps per SM.
        T acc[NG + N];
        bool fconv = true;
        const float2* racc2 = reinterpret_cast<const float2*>(racc);
        constexpr int SPLIT = LSQ_P2_SPLIT;
        if ((LSQ_P2_OPT & 4) && yconv) {
          // f-space half of the Cauchy test, only when the y-space half passed (terminate needs both): its own loop, ahead of
          // the sweeps - one warp-uniform branch per accepted step (inside the sweep it was a reconvergence point per row pair),
          // and before the first half-sweep reuses the old residuals' slot
          auto f_pair = [&](int k) {
            const int idx = k * 32 + lane;
            const float2 re = rtry2[idx], ro = racc2[idx];
            fconv = fconv && (t_abs(re.x - ro.x) < atol + rtol * t_abs(re.x)) && (t_abs(re.y - ro.y) < atol + rtol * t_abs(re.y));
          };
          if (NPAIRS > 0) {
#pragma unroll
            for (int k = 0; k < (NPAIRS > 0 ? NPAIRS : 1); ++k) f_pair(k);
          } else {
            for (int k = 0; k < npairs; ++k) f_pair(k);
          }
        }
#pragma unroll
        for (int half = 0; half < (SPLIT > 0 ? 2 : 1); ++half) {
          float2 acc2[NG + N];
          if (NPAIRS == 0) {
#pragma unroll
            for (int a = 0; a < NG + N; ++a) acc2[a] = make_float2(0.f, 0.f);
          }
          auto jac_pair = [&](int k, bool init) {
            const int idx = k * 32 + lane;
            float2 aux[NAUX], j[N];
#pragma unroll
            for (int q = 0; q < NAUX; ++q) aux[q] = auxs2[q * (mp >> 1) + idx];
            const float2 re = rtry2[idx];
            if ((LSQ_P2_OPT & 32) && SPLIT > 0 && half == 1) PModel::jac_carry(pc, xs2[idx], aux, racc2[idx], j);
            else PModel::jac(pc, xs2[idx], aux, j);
#pragma unroll
            for (int a = 0; a < N; ++a) {
#pragma unroll
              for (int b = 0; b <= a; ++b)
                if (p2_half<PModel>(a, b) == half && !p2_const_entry<PModel>(a, b))
                  acc2[Tri<N>::at(a, b)] = init ? mul2(j[a], j[b]) : fma2(j[a], j[b], acc2[Tri<N>::at(a, b)]);
              if (p2_half<PModel>(a, N) == half) acc2[NG + a] = init ? mul2(j[a], re) : fma2(j[a], re, acc2[NG + a]);
            }
            if (half == 0 && !(LSQ_P2_OPT & 4)) {
              const float2 ro = racc2[idx];
              fconv = fconv && (t_abs(re.x - ro.x) < atol + rtol * t_abs(re.x)) && (t_abs(re.y - ro.y) < atol + rtol * t_abs(re.y));
            }
            if ((LSQ_P2_OPT & 32) && SPLIT > 0 && half == 0) reinterpret_cast<float2*>(racc)[idx] = PModel::carry_out(j);
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
#pragma unroll
            for (int b = 0; b <= a; ++b)
              if (p2_half<PModel>(a, b) == half)
                acc[Tri<N>::at(a, b)] = p2_const_entry<PModel>(a, b) ? T(0) : acc2[Tri<N>::at(a, b)].x + acc2[Tri<N>::at(a, b)].y;
            if (p2_half<PModel>(a, N) == half) acc[NG + a] = acc2[NG + a].x + acc2[NG + a].y;
          }
        }
        if (LSQ_P2_FAST & 4) warp_allreduce_array_f2<NG + N>(acc, scratch, lane);
        else warp_allreduce_array<T, NG + N>(acc, scratch, lane);
#pragma unroll
        for (int a = 0; a < NG; ++a) G[a] = acc[a];
        if (LSQ_P2_OPT & 2) {
          int q = 0;
#pragma unroll
          for (int a = N - PModel::NCONST; a < N; ++a)
#pragma unroll
            for (int b = N - PModel::NCONST; b <= a; ++b) G[Tri<N>::at(a, b)] = gconst[q++];
        }
#pragma unroll
        for (int a = 0; a < N; ++a) g[a] = acc[NG + a];
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
      if (LSQ_P2_FAST & 1) {
        T zz;
        spd_solve_fast<N>(G, lam, g, pstep, zz);
        if (LSQ_P2_FAST & 8) {
          T pp = T(0);
#pragma unroll
          for (int j = 0; j < N; ++j) pp += pstep[j] * pstep[j];
          predn = T(-0.5) * (zz + lam * pp);
        }
      } else {
        spd_solve<T, N>(G, lam, g, pstep);
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

// (A cooperative-algebra variant - a CTA of W = 4 / 8 warps iterating W problems in lock step, the 8 x 8 algebra of all W problems
// in W lanes of one warp between two block barriers - was measured and dropped: bit-identical results, 530 instead of 878
// instructions per problem-iteration, but 24.1 / 26.0 ms against 14.55 ms.  The kernel is bound by the LATENCY of a problem's
// step (sweeps + algebra chain) times the 16 problems resident per SM; taking the algebra out of 3 of 4 warps removes issue
// pressure that was not the limit and halves the warps available to hide that latency.  See git history and DESIGN.md section 10.)
// (A register-lean variant - G and g resident in shared memory between reductions, the Gram sweep in three parts, 94 registers
// without spills, FIVE CTAs = 20 problems per SM - was measured and dropped as well: bit-identical results, 15.15 ms against
// 14.55 ms.  The extra shared-memory traffic and Jacobian-row recomputation cost more than the fifth CTA hides: this kernel has
// no single limiter (ncu, lsq_r10: issue slots 62%, FMA pipe 57%, shared-memory wavefronts 45% of peak).)
// (A two-problems-per-warp variant - the row sweeps of each problem on all 32 lanes, one after the other, the 8 x 8 algebra of the
// two problems at the same time in the two half-warps, G and g in shared memory between iterations, 24 problems resident per SM
// - was measured and dropped in round 2: bit-identical results, ~25% fewer instructions per problem-iteration, 13.28 ms against
// 12.88 ms per 1 Mi fits (2.17 against 1.96 ms at 131 k).  With S the latency of a problem's row sweeps and A that of its algebra
// the SM retires (warps x problems per warp) / (problems per warp x S + A) problem-iterations per unit time: 16 / (S + A) here,
// 24 / (2 S + A) there - equal when S = A, which is where this kernel is.  Shared memory (16 KB per warp for two problems) keeps
// the variant at 12 warps per SM.  See git history, profiles/lm_ab_p2x_r2.log and DESIGN.md section 10.)
// (A half-warp-per-problem variant - two problems per warp sharing the replicated 8 x 8 algebra - was measured and dropped:
// 10% fewer instructions but only 22.8 of 32 lanes active on average (accept / reject and iteration counts differ between the
// two problems) and 12 instead of 16 warps per SM: 16.8 ms against 15.6 ms for 1 Mi fits.  See git history and profiles/README.md.)

}  // namespace b200opt
