// Batched BFGS (use_inverse=True) with backtracking Armijo: one CTA of 4 warps per problem, the dense
// inverse Hessian lives in REGISTERS for the whole solve and never touches HBM (or shared memory).
//
// Restates, per problem:
//   driver       optimistix/_iterate.py:194-268
//   init         optimistix/_solver/quasi_newton.py:176-202, :334-340 (H^-1 = I, :41-75), backtracking.py:51-53
//   step         optimistix/_solver/quasi_newton.py:204-300
//   search       optimistix/_solver/backtracking.py:79-99   (slope 0.1, decrease_factor 0.5, step_init 1.0)
//   update       optimistix/_solver/quasi_newton.py:342-415 (inverse branch :371-385, skipped if q.s <= eps)
//   descent      optimistix/_solver/gauss_newton.py:60-62 (newton = H^-1 g), :126-127
//   termination  optimistix/_misc.py:291-310 on (y, f), skipped on the first step (quasi_newton.py:247-249)
//
// B200 design:
//   * H^-1 (n <= 64: 64x64 fp64 = 32 KiB) is spread over the 128 threads' registers: thread (w, lane) owns row
//     16 w + lane/2, columns 32 (lane & 1) .. +31.  An accepted step makes two passes over it: H q, then the fused
//     rank-2 update + new direction H' g.  That is 2 block barriers per ACCEPTED step.
//   * every warp carries the vector state (y, y_eval, g, direction; 2 elements per lane) and the scalar state
//     redundantly, and evaluates f/grad and all dot products itself.  Identical instruction sequences give
//     bit-identical values in all four warps, so control flow is CTA-uniform without communication, and a
//     REJECTED step (74% of all steps on the Rosenbrock workload) needs no barrier at all.
//   * the backtracking search is evaluated four trial points at a time (BFGS_SPEC): warp w evaluates the point the
//     sequential search would try w rejections from now, y - 0.5^w alpha H^-1 g - the four warps' otherwise redundant
//     evaluations become useful.  One barrier later every warp adopts the first accepted trial; `num_steps` advances
//     by the number of trials the sequential search would have made.  The trials are the same floating-point
//     expressions as in the sequential search (0.5^w is exact), so iterates, counts and result codes are unchanged.
//     78% of all steps on the Rosenbrock workload are rejected ones, each a ~900-cycle dependent chain.
//   * HBM traffic per problem: y0 in; y, num_steps, result (and optionally f, H^-1) out.
#pragma once
#include "common.cuh"
#include "models.cuh"

namespace b200opt {

template <typename T> struct MinParams {
  const T* y0;
  T* y_out;
  int32_t* num_steps;
  int32_t* result;
  T* f_out;             // optional
  T* hinv_out;          // optional [batch][n][n]
  int32_t* num_accepted;  // optional
  int64_t batch;
  int n;
  T rtol, atol;
  int norm;
  int max_steps;
  unsigned long long* queue;  // ticket counter of the dynamic problem queue (common.cuh), or NULL: static grid-stride
};

// next problem of a CTA (block-uniform): static stride, or a ticket from the global counter broadcast through shared memory
__device__ __forceinline__ int64_t next_problem_cta(unsigned long long* queue, int64_t prev, long long* slot) {
  if (queue == nullptr) return prev + gridDim.x;
  __syncthreads();  // every thread is done with the previous problem (and with its read of *slot)
  if (threadIdx.x == 0) *slot = (long long)gridDim.x + (long long)atomicAdd(queue, 1ull);
  __syncthreads();
  return *slot;
}

constexpr int kBfgsN = 64;       // padded dimension
constexpr int kBfgsThreads = 128;

// model family id 3: standard chained Rosenbrock. Lane holds elements e0 = lane and e1 = lane + 32.
template <typename T> struct Rosenbrock {
  // this lane's share of f (the caller sums it over the warp) and its two gradient entries
  __device__ static __forceinline__ void eval_parts(T y0, T y1, int n, int lane, T& fpart, T& g0, T& g1) {
    // neighbours: y[i+1]
    const T d0 = __shfl_down_sync(kFull, y0, 1), d1 = __shfl_down_sync(kFull, y1, 1);
    const T w = __shfl_sync(kFull, y1, 0);
    const T yn0 = lane < 31 ? d0 : w;  // y[lane+1]
    const T yn1 = d1;                  // y[lane+33] (unused for lane 31)
    const int i0 = lane, i1 = lane + 32;
    const bool t0 = i0 < n - 1, t1 = i1 < n - 1;  // term i exists for i < n-1
    const T a0 = t0 ? yn0 - y0 * y0 : T(0), b0 = t0 ? T(1) - y0 : T(0);
    const T a1 = t1 ? yn1 - y1 * y1 : T(0), b1 = t1 ? T(1) - y1 : T(0);
    fpart = (T(100) * a0 * a0 + b0 * b0) + (T(100) * a1 * a1 + b1 * b1);
    // a[i-1]
    const T u0 = __shfl_up_sync(kFull, a0, 1), u1 = __shfl_up_sync(kFull, a1, 1);
    const T a31 = __shfl_sync(kFull, a0, 31);
    const T ap0 = lane > 0 ? u0 : T(0);
    const T ap1 = lane > 0 ? u1 : a31;
    g0 = (t0 ? T(-400) * a0 * y0 - T(2) * b0 : T(0)) + (i0 > 0 && i0 < n ? T(200) * ap0 : T(0));
    g1 = (t1 ? T(-400) * a1 * y1 - T(2) * b1 : T(0)) + (i1 < n ? T(200) * ap1 : T(0));
  }
};

template <typename T> __device__ __forceinline__ T warp_dot2(T a0, T a1, T b0, T b1) { return warp_sum(a0 * b0 + a1 * b1); }
// three butterfly sums in lock step: the same additions as three warp_sum calls (bit-identical results), one reduction's
// latency instead of three (an fp64 round is two SHFL + one DADD, ~35 dependent cycles; five rounds per sum)
template <typename T> __device__ __forceinline__ void warp_sum3(T& a, T& b, T& c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T ta = __shfl_xor_sync(kFull, a, o), tb = __shfl_xor_sync(kFull, b, o), tc = __shfl_xor_sync(kFull, c, o);
    a += ta, b += tb, c += tc;
  }
}

#ifndef BFGS_MINBLOCKS
#define BFGS_MINBLOCKS 4
#endif
// The step, as the kernel runs it (round 2; measured history in DESIGN.md section 4.3 / 10):
//   * the rank-2 update of quasi_newton.py:379-385,  H' = H + c1 s s^T - c2 (Hq s^T + s Hq^T),  c1 = (q.s + q.Hq) / (q.s)^2,
//     c2 = 1 / (q.s),  is applied as  H'_ij = H_ij + b_i s_j + s_i a_j  with  a = c1 s - c2 Hq,  b = -c2 Hq:  2 FMAs per element.
//     The b-half needs only c2, which is known before the pass over H that forms Hq: it is issued BEFORE the q.Hq reduction
//     and the division for c1, whose latency it hides.
//   * every trial warp reduces THREE sums in one interleaved butterfly: f, the Armijo slope g.d and - speculatively, for its
//     own trial point - q.s.  The warp whose trial is adopted has thereby already paid for the reduction the update needs.
//   * max_norm termination is the division-free form  |d_i| < atol + rtol |y_i|  for all i  (same truth value as
//     max_i |d_i| / (atol + rtol |y_i|) < 1 except when a quotient rounds to exactly 1; the LM kernel uses the same form),
//     evaluated speculatively by every trial warp, so the accepted step reads it as one flag.  two_norm / rms_norm keep the
//     quotient form after the adoption.
//   * every broadcast vector is read with 128-bit loads; the two dot-product passes over H keep the single summation chain of
//     the reference order (BFGS_NACC interleaved partial sums were measured: no gain, the chains are not the limiter).
#ifndef BFGS_EARLY_B
#define BFGS_EARLY_B 1  // 0: both halves of the rank-2 update inside the second pass (A/B switch)
#endif
#ifndef BFGS_NACC
#define BFGS_NACC 1     // partial sums per dot-product pass (1, 2 or 4): 77.5 / 77.8 / 78.1 ms - the chains are not the limiter
#endif
template <class Model, typename T>
__global__ void __launch_bounds__(kBfgsThreads, BFGS_MINBLOCKS) bfgs_solve_kernel(const MinParams<T> p) {
  constexpr int N = kBfgsN;
  using Vec = typename Num<T>::Vec;
  constexpr int V = Num<T>::kVec;
  struct WarpVecs {  // per-warp private copies: q, s, new gradient, a - one base address, constant offsets
    T q[N], s[N], g[N], a[N];
  };
  __shared__ __align__(16) WarpVecs wv[4];
  __shared__ __align__(16) T xhq[2][N], xnt[2][N];                    // cross-warp exchange, double buffered
  __shared__ __align__(16) T xye[2][4][N], xge[2][4][N];              // the four warps' trial points and gradients, double buffered
  __shared__ __align__(16) T xfe[2][4], xin[2][4];                    // ... their f and q.s
  __shared__ __align__(16) int xflag[2][4];                           // ... bit 0: Armijo accepts, bit 1: termination test passes
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row = 16 * w + (lane >> 1), hb = lane & 1, c0 = 32 * hb;
  const int n = p.n;
  const int e0 = lane, e1 = lane + 32;
  const T rtol = p.rtol, atol = p.atol;
  const bool maxnorm = p.norm == B200OPT_NORM_MAX;
  const T wscale = T(1) / T(1 << w);  // 0.5^w, exact
  WarpVecs& mv = wv[w];
  const Vec* vq_v = reinterpret_cast<const Vec*>(&mv.q[c0]);
  const Vec* vs_v = vq_v + N / V;      // &mv.s[c0]
  const Vec* vg_v = vq_v + 2 * N / V;  // &mv.g[c0]
  const Vec* va_v = vq_v + 3 * N / V;  // &mv.a[c0]

  __shared__ long long next_slot;
  for (int64_t prob = blockIdx.x; prob < p.batch; prob = next_problem_cta(p.queue, prob, &next_slot)) {
    // ---- init (quasi_newton.py:176-202, :334-340)
    T h[32];
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) h[jj] = (c0 + jj == row) ? T(1) : T(0);
    T y0 = e0 < n ? p.y0[prob * n + e0] : T(0), y1 = e1 < n ? p.y0[prob * n + e1] : T(0);
    T g0 = T(0), g1 = T(0), nt0 = T(0), nt1 = T(0);
    T f = T(0), alpha = T(1);
    bool first = true, terminate = false;
    int result = __all_sync(kFull, t_finite(y0) && t_finite(y1)) ? RES_OK : RES_NONFINITE;
    int steps = 0, naccept = 0, buf = 0, par = 0;

    while (!terminate && result == RES_OK && steps < p.max_steps) {
      const bool y_old_finite = __all_sync(kFull, t_finite(y0) && t_finite(y1));
      // ---- four trial points of the backtracking search at once: warp w takes the one w rejections ahead
      const int left = p.max_steps - steps;
      const int nev = y_old_finite ? (left < 4 ? left : 4) : 1;  // a non-finite y ends the solve after ONE step (_iterate.py:216)
      const T aw = alpha * wscale;  // alpha after w rejections (backtracking.py:96-99): alpha is a power of two, the product is exact
      T ye0 = y0 - aw * nt0;        // gauss_newton.py:126-127 (the first step has nt = 0: y_eval = y0)
      T ye1 = y1 - aw * nt1;
      // f and gradient at the trial point (quasi_newton.py:214-216, lin_to_grad :227); Armijo slope (backtracking.py:79-99);
      // q.s of update_hessian (quasi_newton.py:357-360) for THIS trial
      T fe, ge0, ge1;
      Model::eval_parts(ye0, ye1, n, lane, fe, ge0, ge1);
      T d0 = ye0 - y0, d1 = ye1 - y1;
      T q0 = ge0 - g0, q1 = ge1 - g1;
      T pred = g0 * d0 + g1 * d1, inner = q0 * d0 + q1 * d1;
      warp_sum3(fe, pred, inner);
      {
        const T dfw = fe - f;
        const bool acc_w = first || ((dfw <= T(0.1) * pred) && (pred <= T(0)));
        bool term_w = false;
        if (maxnorm) {  // quasi_newton.py:244-249, cauchy_termination on y and on the scalar f (_misc.py:291-310)
          const bool yc = (e0 >= n || t_abs(d0) < atol + rtol * t_abs(ye0)) && (e1 >= n || t_abs(d1) < atol + rtol * t_abs(ye1));
          const bool fc = t_abs(dfw) < atol + rtol * t_abs(fe);
          term_w = __all_sync(kFull, yc) && fc && !first;
        }
        xye[par][w][e0] = ye0, xye[par][w][e1] = ye1;
        xge[par][w][e0] = ge0, xge[par][w][e1] = ge1;
        if (lane == 0) xfe[par][w] = fe, xin[par][w] = inner, xflag[par][w] = (acc_w ? 1 : 0) | (term_w ? 2 : 0);
      }
      __syncthreads();
      const int4 fl = *reinterpret_cast<const int4*>(xflag[par]);
      int kstar = -1;
      if (3 < nev && (fl.w & 1)) kstar = 3;
      if (2 < nev && (fl.z & 1)) kstar = 2;
      if (1 < nev && (fl.y & 1)) kstar = 1;
      if (0 < nev && (fl.x & 1)) kstar = 0;
      const bool accept = kstar >= 0;
      const int par_now = par;
      par ^= 1;  // a slot is rewritten two rounds later: every warp has passed the next round's barrier by then
      if (accept) {
        // adopt trial kstar (every warp, so that the four copies of the vector state stay bit-identical)
        ye0 = xye[par_now][kstar][e0], ye1 = xye[par_now][kstar][e1];
        ge0 = xge[par_now][kstar][e0], ge1 = xge[par_now][kstar][e1];
        fe = xfe[par_now][kstar];
        inner = xin[par_now][kstar];
        const int flk = kstar == 0 ? fl.x : kstar == 1 ? fl.y : kstar == 2 ? fl.z : fl.w;
        d0 = ye0 - y0, d1 = ye1 - y1;
        q0 = ge0 - g0, q1 = ge1 - g1;
        steps += kstar;  // + 1 below
        alpha = T(1);
        const T df = fe - f;
        // ---- update_hessian (quasi_newton.py:342-415)
        const bool update = inner > Num<T>::eps();  // :360
        mv.s[e0] = d0, mv.s[e1] = d1;
        mv.g[e0] = ge0, mv.g[e1] = ge1;
        mv.q[e0] = q0, mv.q[e1] = q1;
        const T c2 = update ? T(1) / inner : T(0);  // term2 coefficient (:380)
        __syncwarp();
        T s_i = T(0);
        if (update) {
          // pass 1: inv_mvp = H q  (:375)
          T pa[BFGS_NACC];
#pragma unroll
          for (int a = 0; a < BFGS_NACC; ++a) pa[a] = T(0);
#pragma unroll
          for (int jv = 0; jv < 32 / V; ++jv) {
            const Vec v = vq_v[jv];
            const T* vv = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int e = 0; e < V; ++e) pa[(jv * V + e) % BFGS_NACC] += h[jv * V + e] * vv[e];
          }
          T acc = BFGS_NACC == 4 ? (pa[0] + pa[1 % BFGS_NACC]) + (pa[2 % BFGS_NACC] + pa[3 % BFGS_NACC]) : BFGS_NACC == 2 ? pa[0] + pa[1 % BFGS_NACC] : pa[0];
          acc += __shfl_xor_sync(kFull, acc, 1);
          if (hb == 0) xhq[buf][row] = acc;
          __syncthreads();
          const T hq0 = xhq[buf][e0], hq1 = xhq[buf][e1];
          s_i = mv.s[row];
          const T b_i = -c2 * xhq[buf][row];
#if BFGS_EARLY_B
#pragma unroll
          for (int jv = 0; jv < 32 / V; ++jv) {  // H += b s^T: needs c2 only, overlaps the reduction and the division below
            const Vec v = vs_v[jv];
            const T* vv = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int e = 0; e < V; ++e) h[jv * V + e] = fma(b_i, vv[e], h[jv * V + e]);
          }
#endif
          const T mvp_inner = warp_dot2(q0, q1, hq0, hq1);   // :376
          const T c1 = (inner + mvp_inner) / (inner * inner);  // term1 coefficient (:379)
          mv.a[e0] = c1 * d0 - c2 * hq0, mv.a[e1] = c1 * d1 - c2 * hq1;
          __syncwarp();
#if !BFGS_EARLY_B
#pragma unroll
          for (int jv = 0; jv < 32 / V; ++jv) {
            const Vec v = vs_v[jv];
            const T* vv = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int e = 0; e < V; ++e) h[jv * V + e] = fma(b_i, vv[e], h[jv * V + e]);
          }
#endif
        }
        // pass 2: H' = H + s a^T (+ b s^T above)  fused with  newton = H' g_new  (gauss_newton.py:60-61)
        T na[BFGS_NACC];
#pragma unroll
        for (int a = 0; a < BFGS_NACC; ++a) na[a] = T(0);
#pragma unroll
        for (int jv = 0; jv < 32 / V; ++jv) {
          const Vec gv = vg_v[jv];
          const T* gg = reinterpret_cast<const T*>(&gv);
          if (update) {
            const Vec av = va_v[jv];
            const T* aa = reinterpret_cast<const T*>(&av);
#pragma unroll
            for (int e = 0; e < V; ++e) h[jv * V + e] = fma(s_i, aa[e], h[jv * V + e]);
          }
#pragma unroll
          for (int e = 0; e < V; ++e) na[(jv * V + e) % BFGS_NACC] += h[jv * V + e] * gg[e];
        }
        T nacc = BFGS_NACC == 4 ? (na[0] + na[1 % BFGS_NACC]) + (na[2 % BFGS_NACC] + na[3 % BFGS_NACC]) : BFGS_NACC == 2 ? na[0] + na[1 % BFGS_NACC] : na[0];
        nacc += __shfl_xor_sync(kFull, nacc, 1);
        if (hb == 0) xnt[buf][row] = nacc;
        __syncthreads();
        nt0 = xnt[buf][e0], nt1 = xnt[buf][e1];
        buf ^= 1;
        // ---- termination (quasi_newton.py:244-249; cauchy_termination on y and on the scalar f)
        if (maxnorm) {
          terminate = (flk & 2) != 0;
        } else {
          NormAcc<T> yn(p.norm);
          if (e0 < n) yn.add(t_abs(d0) / (atol + rtol * t_abs(ye0)));
          if (e1 < n) yn.add(t_abs(d1) / (atol + rtol * t_abs(ye1)));
          yn.warp_combine();
          NormAcc<T> fnrm(p.norm);
          fnrm.add(t_abs(df) / (atol + rtol * t_abs(fe)));
          terminate = (yn.finish(n) < T(1)) && (fnrm.finish(1) < T(1)) && !first;
        }
        ++naccept;
        y0 = ye0, y1 = ye1, f = fe, g0 = ge0, g1 = ge1;
      } else {
        // nev rejections: alpha * 0.5^nev, exact (alpha is a power of two)
        alpha *= nev >= 4 ? T(0.0625) : nev == 3 ? T(0.125) : nev == 2 ? T(0.25) : T(0.5);
        steps += nev - 1;
        terminate = false;
      }
      result = RES_OK;
      if (!y_old_finite) result = RES_NONFINITE;
      first = false;
      ++steps;
    }
    if (result == RES_OK && !terminate) result = RES_NL_MAX_STEPS;

    if (w == 0) {
      if (e0 < n) p.y_out[prob * n + e0] = y0;
      if (e1 < n) p.y_out[prob * n + e1] = y1;
      if (lane == 0) {
        p.num_steps[prob] = steps;
        p.result[prob] = result;
        if (p.f_out) p.f_out[prob] = f;
        if (p.num_accepted) p.num_accepted[prob] = naccept;
      }
    }
    if (p.hinv_out && row < n) {
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
        if (c0 + jj < n) p.hinv_out[(prob * n + row) * n + c0 + jj] = h[jj];
    }
    __syncthreads();  // exchange buffers are reused by the next problem
  }
}

}  // namespace b200opt
