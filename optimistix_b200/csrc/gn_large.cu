// Support kernels of the single large Gauss-Newton problem (BASELINE configs[4]):
//   b200opt_gn_solve        (G + lambda I) p = g by blocked Cholesky in fp64, one CTA (n = 256: 5.6 Mflop, barrier-latency bound)
//   b200opt_linear_residual r = J y - d for the "linear_rows" family, one warp per row, coalesced float4 reads,
//                           fused with the f-space half of cauchy_termination (optimistix/_misc.py:306-309):
//                           max_i |r_new_i - r_old_i| / (atol + rtol |r_new_i|)
#include "common.cuh"
#include "gn_state.cuh"

namespace b200opt {
size_t gn_workspace_solve_offset();
size_t gn_workspace_sync_offset();
int32_t gn_iterate_gram(const float* J, const float* d, float* r0, float* r1, int64_t m_local, int32_t n, const void* state, float* out,
                        void* workspace, const void* const* xchg_bases_dev, void* xchg_local, int32_t world, int32_t rank, float rtol,
                        float atol, int32_t fold_reduce, void* stream);

// Blocked right-looking Cholesky + both triangular solves on ONE CTA (the whole problem is 5.6 Mflop; what matters is
// the number of block-wide barriers on the critical path, not the flop rate):
//   * panels of 32 columns live in shared memory (unscaled working copy P and the finished, scaled factor Lp);
//     a column step costs ONE barrier: the scaled column is written to Lp / the forward substitution is applied to z
//     while the remaining panel columns are updated from the unscaled values, and the thread that produces the next
//     pivot also produces its reciprocal square root;
//   * the trailing matrix (fp64, global memory, L2 resident) is updated once per panel (rank-32) with 64x64 tiles of
//     4x4 register micro-tiles;
//   * the backward substitution walks the panels in reverse: a parallel mat-vec with the already known part of p,
//     then a 32-step in-warp substitution on the diagonal block using the saved reciprocals.
// A non-positive pivot gives NaN/inf, which propagates to p and is reported as `singular` (as lineax does).
constexpr int SOLVE_PW = 32;                 // panel width
constexpr int SOLVE_LD = SOLVE_PW + 1;       // row pitch of the panel buffers (odd: conflict-free row and column access)
constexpr size_t SOLVE_SMEM = (size_t)(256 * SOLVE_LD + 3 * 256 + 3 * SOLVE_PW + 8) * sizeof(double);

// `parts` row-shard partials [G | g | rr | f-statistic bits] (stride `part_stride` floats) are summed here, in rank order
// and in fp64, so that after ONE all-gather every rank holds bit-identical normal equations (and the max of the
// f-space termination statistic) without a second collective.
// `flags` (optional): the peer-exchange flags of this iteration's parity; the kernel first waits until every rank has
// published sequence number `seq` (bounded spin: a missing peer becomes an error code, not a hang).
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Wait until every rank has published sequence number `seq` (bounded: a missing peer becomes an error code, not a hang; the
// bound is wall-clock nanoseconds from %globaltimer, not SM cycles).  Returns 0 on time-out.  All threads call it.
__device__ __forceinline__ int wait_for_peers(const unsigned* flags, int nparts, unsigned seq, int timeout_ms) {
  int ok = 1;
  if ((int)threadIdx.x < nparts) {
    const volatile unsigned* f = flags + threadIdx.x;
    const unsigned long long t0 = global_timer_ns();
    const unsigned long long limit = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 30000) * 1000000ull;
    while (*f != seq) {
      __nanosleep(100);
      if (global_timer_ns() - t0 > limit) {
        ok = 0;
        break;
      }
    }
  }
  __threadfence_system();  // acquire: the peers' stores that preceded their flags are visible to the loads below
  return __syncthreads_and(ok);
}

// Sum of the row-shard partials in rank order, in fp64: the lower triangle of G + lambda I into A (global, L2 resident), g
// into z (shared), and the two scalars.  Ends with a block barrier.
__device__ __forceinline__ void sum_parts64(const float* gram_out, int nparts, int64_t part_stride, int n, double lambda, double* A,
                                            double* z, float* rr_out, float* fs_out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, j = e - i * n;
    if (j <= i) {
      double acc = 0.0;
      for (int q = 0; q < nparts; ++q) acc += (double)__ldcg(gram_out + q * part_stride + e);  // L2: peers write these
      A[e] = acc + (i == j ? lambda : 0.0);
    }
  }
  for (int i = tid; i < n; i += nt) {
    double acc = 0.0;
    for (int q = 0; q < nparts; ++q) acc += (double)__ldcg(gram_out + q * part_stride + n * n + i);
    z[i] = acc;
  }
  if (tid == 0 && (rr_out || fs_out)) {
    double rr = 0.0;
    float fs = 0.f;
    for (int q = 0; q < nparts; ++q) {
      rr += (double)__ldcg(gram_out + q * part_stride + n * n + n);
      fs = nan_max(fs, __ldcg(gram_out + q * part_stride + n * n + n + 1));
    }
    if (rr_out) *rr_out = (float)rr;
    if (fs_out) *fs_out = fs;
  }
  __syncthreads();
}

// Blocked fp64 Cholesky + both triangular solves: A (lower triangle, global) and z (shared) in, pv (shared) out.
__device__ void solve64(int n, double* A, double* sm) {
  double* Lp = sm;                          // [256][SOLVE_LD] finished panel of the factor (rows c0..n-1 used)
  double* z = Lp + 256 * SOLVE_LD;          // [256] right-hand side -> L^-1 g
  double* rinv = z + 256;                   // [256] 1 / L_kk
  double* pv = rinv + 256;                  // [256] solution
  double* zs = pv + 256;                    // [SOLVE_PW] panel right-hand side of the backward sweep
  double* colbuf = zs + SOLVE_PW;           // [2][SOLVE_PW] unscaled column of the diagonal block (ping-pong)
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;

  const int npanels = (n + SOLVE_PW - 1) / SOLVE_PW;
  for (int pb = 0; pb < npanels; ++pb) {
    const int c0 = pb * SOLVE_PW;
    const int pw = min(SOLVE_PW, n - c0);
    // ---- the panel lives in REGISTERS: thread t holds columns 8 cg .. 8 cg + 7 of row c0 + t / 4 (cg = t % 4), so
    // the four threads of a row are adjacent lanes.  Column step c (k = c0 + c), ONE barrier:
    //   l_i   = P[i][c] / sqrt(pivot): from the owner lane of the row by shuffle; the owner also stores it to Lp and
    //           applies the fused forward substitution to z[i]
    //   P[i][cp] -= l_i * l_(c0+cp)  for cp > c: the unscaled column-c entries of the 32 diagonal-block rows were
    //           published to `colbuf` by their owners at the end of step c - 1 (after their last update), together
    //           with the reciprocal square root of the next pivot
    const int prow = c0 + (tid >> 2), cg = tid & 3;
    const bool prow_on = prow < n;
    double pr[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int cp = 8 * cg + e;
      pr[e] = (prow_on && cp < pw && c0 + cp <= prow) ? A[(size_t)prow * n + c0 + cp] : 0.0;
    }
    if (prow_on && prow - c0 < SOLVE_PW && cg == 0) {
      colbuf[prow - c0] = pr[0];
      if (prow == c0) rinv[c0] = rsqrt(pr[0]);
    }
    __syncthreads();
#pragma unroll
    for (int cgc = 0; cgc < 4; ++cgc) {
#pragma unroll
      for (int ec = 0; ec < 8; ++ec) {
        const int c = 8 * cgc + ec;
        if (c < pw) {                              // block-uniform
          const int k = c0 + c;
          const double* cb = colbuf + (c & 1) * SOLVE_PW;
          double* cbn = colbuf + ((c + 1) & 1) * SOLVE_PW;
          const double inv = rinv[k];
          const double zk = z[k] * inv;
          const double li = __shfl_sync(kFull, pr[ec], (lane & ~3) | cgc) * inv;   // L[prow][k]
          if (prow_on && prow >= k) {
            if (cg == cgc) {
              Lp[prow * SOLVE_LD + c] = li;
              if (prow > k) z[prow] -= li * zk;
              else pv[k] = zk;                       // pv holds L^-1 g until the backward sweep overwrites it
            }
            if (prow > k) {
              const double lis = li * inv;           // l_i / sqrt(pivot): the published column entries are unscaled
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int cp = 8 * cg + e;
                if (cp > c) pr[e] = fma(-lis, cb[cp], pr[e]);
              }
              // publish column c + 1 of the diagonal-block rows (now final) and the next reciprocal pivot root
              if (c + 1 < pw && prow - c0 < SOLVE_PW && cg == ((c + 1) >> 3)) {
                const double v = pr[(ec + 1) & 7];
                cbn[prow - c0] = v;
                if (prow == k + 1) rinv[k + 1] = rsqrt(v);
              }
            }
          }
          __syncthreads();
        }
      }
    }
    // ---- store the finished panel (needed by the backward sweep) and apply the rank-pw update to the trailing matrix
    for (int e = tid; e < (n - c0) * SOLVE_PW; e += nt) {
      const int i = c0 + (e >> 5), c = e & 31;
      if (c < pw && c0 + c <= i) A[(size_t)i * n + c0 + c] = Lp[i * SOLVE_LD + c];
    }
    const int t0 = c0 + pw;                      // first trailing row / column
    const int R = n - t0;
    if (R > 0) {
      const int nbt = (R + 63) >> 6;             // 64 x 64 tiles, lower block triangle
      const int ntiles = nbt * (nbt + 1) / 2;
      const int quarter = tid >> 8, t256 = tid & 255, tx = t256 & 15, ty = t256 >> 4;
      for (int tile = quarter; tile < ntiles; tile += 4) {
        int bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= tile) ++bi;
        const int bj = tile - bi * (bi + 1) / 2;
        const int i0 = t0 + 64 * bi + ty, j0 = t0 + 64 * bj + tx;  // lanes run along j: A accesses are coalesced
        double acc[4][4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) acc[e][f] = 0.0;
        for (int c = 0; c < pw; ++c) {
          double li[4], lj[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i0 + 16 * e, j = j0 + 16 * e;
            li[e] = i < n ? Lp[i * SOLVE_LD + c] : 0.0;
            lj[e] = j < n ? Lp[j * SOLVE_LD + c] : 0.0;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[e][f] = fma(li[e], lj[f], acc[e][f]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const int i = i0 + 16 * e, j = j0 + 16 * f;
            if (i < n && j <= i) A[(size_t)i * n + j] -= acc[e][f];
          }
      }
    }
    __syncthreads();
  }

  // ---- backward sweep: L^T p = z  (pv currently holds z = L^-1 g)
  for (int pb = npanels - 1; pb >= 0; --pb) {
    const int c0 = pb * SOLVE_PW;
    const int pw = min(SOLVE_PW, n - c0);
    for (int e = tid; e < (n - c0) * SOLVE_PW; e += nt) {
      const int i = c0 + (e >> 5), c = e & 31;
      Lp[i * SOLVE_LD + c] = (c < pw && c0 + c <= i) ? A[(size_t)i * n + c0 + c] : 0.0;
    }
    __syncthreads();
    {  // warp c: s_c = sum_{i >= c0 + pw} L[i][c0 + c] p_i
      const int c = warp;
      double sacc = 0.0;
      if (c < pw)
        for (int i = c0 + pw + lane; i < n; i += 32) sacc = fma(Lp[i * SOLVE_LD + c], pv[i], sacc);
      sacc = warp_sum(sacc);
      if (lane == 0 && c < pw) zs[c] = pv[c0 + c] - sacc;
    }
    __syncthreads();
    if (warp == 0) {
      double x = lane < pw ? zs[lane] : 0.0;
      for (int c = pw - 1; c >= 0; --c) {
        const double pc = __shfl_sync(kFull, x, c) * rinv[c0 + c];
        if (lane < c) x -= Lp[(c0 + c) * SOLVE_LD + lane] * pc;   // (L^T)[lane][c] = L[c][lane]
        if (lane == c) x = pc;
      }
      if (lane < pw) pv[c0 + lane] = x;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024, 1) gn_solve_kernel(const float* gram_out, int nparts, int64_t part_stride, int n, double lambda,
                                                           double* A, float* p_out, int32_t* status, float* summary,
                                                           const unsigned* flags, unsigned seq) {
  extern __shared__ double sm[];
  double* z = sm + 256 * SOLVE_LD;
  double* pv = z + 512;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (flags && !wait_for_peers(flags, nparts, seq, 2000)) {  // legacy entry point: 2 s
    if (tid == 0 && status) *status = B200OPT_EXCHANGE_TIMEOUT;
    return;
  }
  sum_parts64(gram_out, nparts, part_stride, n, lambda, A, z, summary ? summary : nullptr, summary ? summary + 1 : nullptr);
  solve64(n, A, sm);
  int ok = 1;
  for (int i = tid; i < n; i += nt) {
    p_out[i] = (float)pv[i];
    if (!isfinite(pv[i])) ok = 0;
  }
  ok = __syncthreads_and(ok);
  if (tid == 0 && status) *status = ok ? RES_OK : RES_SINGULAR;  // lineax: non-finite solution => singular
}


// =====================================================================================================================
// Fast path of the state-driven iteration: Jacobi-scaled fp32 blocked Cholesky in shared memory + fp64 refinement.
// The normal equations of configs[4] come from TF32 operands (10-bit mantissas): an fp64 factorisation is far more exact
// than its input, and it is what bounded the 8-GPU iteration (0.28 ms on one SM against a 1.6 ms Gram pass).  Here the
// 256 x 256 system is equilibrated (D G D with D = diag(G)^-1/2: unit diagonal, which removes the column scaling of J
// from the condition number), factorised in fp32 as 8 x 8 blocks of 32 x 32 held in shared memory, and the solution is
// refined against the fp64 system: rho = g - G p (fp64), p += L^-T L^-1 rho.  The size of the refinement step measures
// the quality of the fp32 factor: if it is not small (or a pivot is not positive) the kernel falls back to the fp64
// factorisation, so an ill-conditioned system gets exactly the answer it got before.
//   per block column k:  (a) one warp factorises the 32 x 32 diagonal block in registers (shuffles, no block barrier),
//                        (b) one thread per row below solves its row against the block (forward substitution); meanwhile
//                            one warp inverts the diagonal block (used by the triangular solves at the end),
//                        (c) all warps apply the rank-32 update to the trailing blocks, 4 x 4 register tiles, 128-bit
//                            shared-memory loads (rows interleaved by 8: conflict free).
// Three block barriers per block column; the triangular solves are mat-vecs with the inverted diagonal blocks.
constexpr int FB = 32;                  // block size
constexpr int FNB = 8;                  // 256 / 32
constexpr int FP = 36;                  // row pitch of a block in floats: 16-byte aligned rows, LDS.128 conflict free
constexpr int FBLK = FB * FP;           // floats per block
constexpr int FNBLK = FNB * (FNB + 1) / 2;
__device__ __forceinline__ int fblk(int i, int j) { return (i * (i + 1) / 2 + j) * FBLK; }  // block (i, j), j <= i
struct FastSmem {
  float S[FNBLK * FBLK];    // lower block triangle of D G D, overwritten by its Cholesky factor
  float Li[FNB * FBLK];     // inverses of the diagonal blocks of the factor
  float T[FBLK];            // transpose of the current diagonal block (column access for the row solves)
  double g[256];            // right-hand side (fp64, unscaled)
  double p[256];            // solution (fp64, unscaled)
  double rho[256];          // refinement residual
  float sc[256];            // D
  float dinv[256];          // 1 / L_ii
  float rhs[256];           // working right-hand side of the triangular solves (scaled system)
  float zk[32];             // current block of the triangular-solve solution
  float red[64];
  int flags[4];
};
constexpr size_t FAST_SMEM = sizeof(FastSmem);
static_assert(FAST_SMEM <= 227 * 1024, "fast solve does not fit the shared memory of one CTA");

// L L^T x = rhs on the scaled system (fs.rhs in, fs.rhs out = x), using the inverted diagonal blocks.
__device__ void fast_tri_solves(FastSmem& fs) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // forward: L z = rhs
  for (int k = 0; k < FNB; ++k) {
    {  // z_k = Li_kk rhs_k: warp i forms row i
      float v = (lane <= warp) ? fs.Li[k * FBLK + warp * FP + lane] * fs.rhs[32 * k + lane] : 0.f;
      v = warp_sum(v);
      if (lane == 0) fs.zk[warp] = v;
    }
    __syncthreads();
    if (tid < 32) fs.rhs[32 * k + tid] = fs.zk[tid];
    // rhs_i -= L_ik z_k for the block rows below: one thread per row (128-bit loads of the row: conflict free at pitch 36, z_k as
    // broadcasts) - a warp per row with a shuffle reduction took up to seven rounds of ~200 cycles per block column
    const int R = 32 * (FNB - 1 - k);
    if (tid < R) {
      const int ib = k + 1 + (tid >> 5), rr = tid & 31;
      const float* Lr = fs.S + fblk(ib, k) + rr * FP;
      float acc = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 l = *reinterpret_cast<const float4*>(Lr + 4 * c4);
        const float4 z = *reinterpret_cast<const float4*>(fs.zk + 4 * c4);
        acc = fmaf(l.x, z.x, acc);
        acc = fmaf(l.y, z.y, acc);
        acc = fmaf(l.z, z.z, acc);
        acc = fmaf(l.w, z.w, acc);
      }
      fs.rhs[32 * ib + rr] -= acc;
    }
    __syncthreads();
  }
  // backward: L^T x = z
  for (int k = FNB - 1; k >= 0; --k) {
    {  // x_k = Li_kk^T z_k: warp i forms sum_t Li[t][i] z[t], t >= i
      float v = (lane >= warp) ? fs.Li[k * FBLK + lane * FP + warp] * fs.rhs[32 * k + lane] : 0.f;
      v = warp_sum(v);
      if (lane == 0) fs.zk[warp] = v;
    }
    __syncthreads();
    if (tid < 32) fs.rhs[32 * k + tid] = fs.zk[tid];
    // rhs_j -= L_kj^T x_k for the block columns to the left: one thread per (block column, column)
    if (tid < 32 * k) {
      const int jb = tid >> 5, c = tid & 31;
      const float* L = fs.S + fblk(k, jb);
      float acc = 0.f;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) acc = fmaf(L[r * FP + c], fs.zk[r], acc);
      fs.rhs[32 * jb + c] -= acc;
    }
    __syncthreads();
  }
}

// Returns 1 when fs.p holds an accurate solution of (G + lambda I) p = g, 0 when the caller must use the fp64 path.
// A: the fp64 sum of the partials (full symmetric matrix, global memory), written here.
__device__ int solve32(const float* gram_out, int nparts, int64_t part_stride, double lambda, double* A, FastSmem& fs,
                       float* rr_out, float* fs_out) {
  constexpr int n = 256;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 4) fs.flags[tid] = 0;
  // ---- diagonal, scaling, right-hand side, scalars
  if (tid < n) {
    double acc = 0.0, gg = 0.0;
    for (int q = 0; q < nparts; ++q) {
      acc += (double)__ldcg(gram_out + q * part_stride + tid * n + tid);
      gg += (double)__ldcg(gram_out + q * part_stride + n * n + tid);
    }
    acc += lambda;
    fs.sc[tid] = (float)rsqrt(acc);  // acc <= 0 or NaN -> NaN / inf: caught below
    fs.g[tid] = gg;
  }
  if (tid == 0) {
    double rr = 0.0;
    float fsv = 0.f;
    for (int q = 0; q < nparts; ++q) {
      rr += (double)__ldcg(gram_out + q * part_stride + n * n + n);
      fsv = nan_max(fsv, __ldcg(gram_out + q * part_stride + n * n + n + 1));
    }
    *rr_out = (float)rr;
    *fs_out = fsv;
  }
  __syncthreads();
  {
    int bad = 0;
    if (tid < n) bad = !(fs.sc[tid] > 0.f && fs.sc[tid] < 3.0e38f);
    if (__syncthreads_or(bad)) return 0;
  }
  // ---- fp64 sum of the partials -> A (full, for the refinement mat-vec) and the scaled fp32 lower block triangle
  for (int e4 = tid; e4 < n * n / 4; e4 += 1024) {
    const int e = 4 * e4, i = e >> 8, j = e & 255;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int q = 0; q < nparts; ++q) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(gram_out + q * part_stride + e));
      a0 += (double)v.x, a1 += (double)v.y, a2 += (double)v.z, a3 += (double)v.w;
    }
    if (i >= j && i < j + 4) {  // the diagonal entry lies in this quad
      const int d = i - j;
      if (d == 0) a0 += lambda;
      else if (d == 1) a1 += lambda;
      else if (d == 2) a2 += lambda;
      else a3 += lambda;
    }
    reinterpret_cast<double2*>(A + e)[0] = make_double2(a0, a1);
    reinterpret_cast<double2*>(A + e)[1] = make_double2(a2, a3);
    if ((j >> 5) <= (i >> 5)) {  // block (i/32, j/32) of the lower block triangle (diagonal blocks are stored in full)
      const double si = (double)fs.sc[i];
      float4 w;
      w.x = (float)(a0 * si * (double)fs.sc[j]);
      w.y = (float)(a1 * si * (double)fs.sc[j + 1]);
      w.z = (float)(a2 * si * (double)fs.sc[j + 2]);
      w.w = (float)(a3 * si * (double)fs.sc[j + 3]);
      *reinterpret_cast<float4*>(fs.S + fblk(i >> 5, j >> 5) + (i & 31) * FP + (j & 31)) = w;
    }
  }
  __syncthreads();

  // ---- blocked Cholesky
  for (int k = 0; k < FNB; ++k) {
    float* D = fs.S + fblk(k, k);
    if (warp == 0) {  // (a) diagonal block: lane = row, the row lives in registers
      float a[32];
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 v = *reinterpret_cast<const float4*>(D + lane * FP + 4 * c4);
        a[4 * c4] = v.x, a[4 * c4 + 1] = v.y, a[4 * c4 + 2] = v.z, a[4 * c4 + 3] = v.w;
      }
      int badp = 0;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float piv = __shfl_sync(kFull, a[c], c);
        if (!(piv > 0.f)) badp = 1;  // uniform: every lane sees the same pivot
        const float inv = rsqrtf(piv);
        const float l = a[c] * inv;   // lane c: sqrt(piv); lanes > c: L[lane][c]; lanes < c: unused
        a[c] = l;
        if (lane == c) fs.dinv[32 * k + c] = inv;
#pragma unroll
        for (int j = c + 1; j < 32; ++j) {
          const float lj = __shfl_sync(kFull, l, j);
          a[j] = fmaf(-l, lj, a[j]);  // meaningful for lane >= j
        }
      }
      if (badp && lane == 0) fs.flags[0] = 1;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        float4 v;
        v.x = (4 * c4 <= lane) ? a[4 * c4] : 0.f;
        v.y = (4 * c4 + 1 <= lane) ? a[4 * c4 + 1] : 0.f;
        v.z = (4 * c4 + 2 <= lane) ? a[4 * c4 + 2] : 0.f;
        v.w = (4 * c4 + 3 <= lane) ? a[4 * c4 + 3] : 0.f;
        *reinterpret_cast<float4*>(D + lane * FP + 4 * c4) = v;  // upper triangle zeroed
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) fs.T[c * FP + lane] = (c <= lane) ? a[c] : 0.f;  // T[c][j] = L[j][c]
    }
    __syncthreads();
    if (fs.flags[0]) return 0;  // non-positive pivot in fp32: let the fp64 path decide
    // (b) rows below the diagonal block: x L_kk^T = a, one thread per row; warp 31 inverts the diagonal block meanwhile
    const int R = 32 * (FNB - 1 - k);
    if (tid < R) {
      float* row = fs.S + fblk(k + 1 + (tid >> 5), k) + (tid & 31) * FP;
      float a[32];
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 v = *reinterpret_cast<const float4*>(row + 4 * c4);
        a[4 * c4] = v.x, a[4 * c4 + 1] = v.y, a[4 * c4 + 2] = v.z, a[4 * c4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float x = a[c] * fs.dinv[32 * k + c];
        a[c] = x;
#pragma unroll
        for (int j = c + 1; j < 32; ++j) a[j] = fmaf(-x, fs.T[c * FP + j], a[j]);  // T[c][j] = L[j][c]: broadcast, contiguous in j
      }
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        *reinterpret_cast<float4*>(row + 4 * c4) = make_float4(a[4 * c4], a[4 * c4 + 1], a[4 * c4 + 2], a[4 * c4 + 3]);
    } else if (warp == 31) {
      // lane j computes column j of L_kk^-1 by forward substitution: x_i = (delta_ij - sum_{t<i} L_it x_t) / L_ii
      float x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float sacc = (i == lane) ? 1.f : 0.f;
#pragma unroll
        for (int t = 0; t < i; ++t) sacc = fmaf(-D[i * FP + t], x[t], sacc);  // x[t] == 0 for t < lane
        x[i] = sacc * fs.dinv[32 * k + i];
        fs.Li[k * FBLK + i * FP + lane] = x[i];
      }
    }
    __syncthreads();
    // (c) trailing update: C_ij -= L_ik L_jk^T for k < j <= i; 64 threads per 32 x 32 block, 4 x 4 tiles with rows
    // interleaved by 8 (thread (tr, tc) owns rows tr + 8 a and columns tc + 8 b): conflict-free 128-bit loads
    {
      const int nb = FNB - 1 - k;
      const int nblk = nb * (nb + 1) / 2;
      const int t64 = tid & 63, tr = t64 >> 3, tc = t64 & 7;
      for (int blk = tid >> 6; blk < nblk; blk += 16) {
        int bi = 0;
        while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
        const int bj = blk - bi * (bi + 1) / 2;
        const int I = k + 1 + bi, Jb = k + 1 + bj;
        const float* Lrow = fs.S + fblk(I, k) + tr * FP;
        const float* Lcol = fs.S + fblk(Jb, k) + tc * FP;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll
        for (int t = 0; t < 32; t += 4) {
          float4 li[4], lj[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) li[a] = *reinterpret_cast<const float4*>(Lrow + 8 * a * FP + t);
#pragma unroll
          for (int b = 0; b < 4; ++b) lj[b] = *reinterpret_cast<const float4*>(Lcol + 8 * b * FP + t);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              acc[a][b] = fmaf(li[a].x, lj[b].x, acc[a][b]);
              acc[a][b] = fmaf(li[a].y, lj[b].y, acc[a][b]);
              acc[a][b] = fmaf(li[a].z, lj[b].z, acc[a][b]);
              acc[a][b] = fmaf(li[a].w, lj[b].w, acc[a][b]);
            }
        }
        float* C = fs.S + fblk(I, Jb);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) C[(tr + 8 * a) * FP + tc + 8 * b] -= acc[a][b];
      }
    }
    __syncthreads();
  }

  // ---- solve + refinement against the fp64 system
  if (tid < n) {
    fs.rhs[tid] = (float)(fs.g[tid] * (double)fs.sc[tid]);
    fs.p[tid] = 0.0;
  }
  __syncthreads();
  float step_ratio = 0.f;
  for (int pass = 0; pass < 3; ++pass) {
    fast_tri_solves(fs);  // fs.rhs: scaled right-hand side -> scaled solution
    // p += D x; measure the step against the solution
    float dmax = 0.f, pmax = 0.f;
    if (tid < n) {
      const double dx = (double)fs.rhs[tid] * (double)fs.sc[tid];
      fs.p[tid] += dx;
      dmax = fabsf((float)dx);
      pmax = fabsf((float)fs.p[tid]);
      if (!(dmax == dmax)) dmax = __int_as_float(0x7f800000);
    }
    dmax = warp_nanmax_nonneg(dmax);
    pmax = warp_nanmax_nonneg(pmax);
    if (lane == 0) fs.red[warp] = dmax, fs.red[32 + warp] = pmax;
    __syncthreads();
    float dm = 0.f, pm = 0.f;
    for (int w = 0; w < 8; ++w) dm = fmaxf(dm, fs.red[w]), pm = fmaxf(pm, fs.red[32 + w]);
    step_ratio = dm / fmaxf(pm, 1e-37f);
    // a refinement step of relative size s leaves an error of about s^2 (the factor's relative accuracy is s): s <= 1e-4
    // is below the resolution of the fp32 output
    if (pass > 0 && step_ratio <= 1e-4f) break;
    if (pass == 2) break;
    // rho = g - A p in fp64 (A full symmetric, row-major in global memory / L2): warp w takes rows w, w + 32, ...
    for (int i = warp; i < n; i += 32) {
      const double* Ai = A + (size_t)i * n;
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc = fma(Ai[lane + 32 * c], fs.p[lane + 32 * c], acc);
      acc = warp_sum(acc);
      if (lane == 0) fs.rho[i] = fs.g[i] - acc;
    }
    __syncthreads();
    if (tid < n) fs.rhs[tid] = (float)(fs.rho[tid] * (double)fs.sc[tid]);
    __syncthreads();
  }
  // the last refinement step must be small against the solution: otherwise the fp32 factor is too inexact for this matrix
  return (step_ratio <= 1e-3f) ? 1 : 0;
}

// (G + lambda I) p = g for n = 256 with either solver; returns the fp64 solution vector (shared memory, valid after the
// function's final barrier).  *fast_out: 1 when the fp32 + refinement path produced it.
__device__ const double* solve_normal_equations(const float* parts, int nparts, int64_t part_stride, double lambda, double* A,
                                                unsigned char* smem_raw, int solver, float* rr_out, float* fs_out, int* fast_out) {
  int fast = 0;
  const double* pv = nullptr;
  if (solver == 0) {
    FastSmem& fs = *reinterpret_cast<FastSmem*>(smem_raw);
    fast = solve32(parts, nparts, part_stride, lambda, A, fs, rr_out, fs_out);  // uniform over the block
    pv = fs.p;
  }
  if (!fast) {
    double* sm = reinterpret_cast<double*>(smem_raw);
    double* z = sm + 256 * SOLVE_LD;
    __syncthreads();
    sum_parts64(parts, nparts, part_stride, 256, lambda, A, z, rr_out, fs_out);
    solve64(256, A, sm);
    pv = z + 512;
  }
  __syncthreads();
  *fast_out = fast;
  return pv;
}

// b200opt_gn_solve_fast: the solver of the state-driven iteration as a stand-alone call (n = 256)
__global__ void __launch_bounds__(1024, 1) gn_solve_fast_kernel(const float* parts, int nparts, int64_t part_stride, double lambda, double* A,
                                                                int solver, float* p_out, int32_t* status, float* summary, int32_t* info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_rr, s_fs;
  int fast = 0;
  const double* pv = solve_normal_equations(parts, nparts, part_stride, lambda, A, smem_raw, solver, &s_rr, &s_fs, &fast);
  const int tid = threadIdx.x;
  int ok = 1;
  if (tid < 256) {
    p_out[tid] = (float)pv[tid];
    ok = isfinite(pv[tid]);
  }
  ok = __syncthreads_and(ok);
  if (tid == 0) {
    if (status) *status = ok ? RES_OK : RES_SINGULAR;
    if (summary) summary[0] = s_rr, summary[1] = s_fs;
    if (info) *info = fast;
  }
}

// One iteration's second half (b200opt_gn_iterate): wait for the peers, sum, solve, and the rest of
// AbstractGaussNewton.step + the driver's bookkeeping on the device state (gauss_newton.py:276-334, _iterate.py:216-257).
__global__ void __launch_bounds__(1024, 1) gn_step_kernel(const float* local_out, const float* xchg_parts, int nparts, int64_t part_stride,
                                                          double lambda, double* A, GnState* state, float* fstat_slot,
                                                          const unsigned* flags_base, unsigned* seq_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int n = 256;
  const int tid = threadIdx.x;
  if (gn_frozen(state)) return;
  const float* parts = local_out;
  if (flags_base) {
    const unsigned seq = *seq_counter + 1u;
    const int parity = (int)(seq & 1u);
    if (!wait_for_peers(flags_base + parity * nparts, nparts, seq, state->timeout_ms)) {
      if (tid == 0) state->result = B200OPT_EXCHANGE_TIMEOUT;
      return;
    }
    parts = xchg_parts + (size_t)parity * nparts * part_stride;
  } else {
    nparts = 1;
  }
  __shared__ float s_rr, s_fs;
  int fast = 0;
  const double* pv = solve_normal_equations(parts, nparts, part_stride, lambda, A, smem_raw, state->solver, &s_rr, &s_fs, &fast);
  // ---- accepted branch of the step (LearningRate(1.0) always accepts): Cauchy test, new iterate, result code
  int fin = 1, yconv = 1, pfin = 1;
  float pnew = 0.f, ye = 0.f;
  if (tid < n) {
    const float yo = state->y[tid];
    ye = state->y_eval[tid];
    fin = isfinite(yo);                                                    // _iterate.py:216,238: the carry's y before this step
    yconv = fabsf(ye - yo) / (state->atol + state->rtol * fabsf(ye)) < 1.f;  // _misc.py:306-308 (NaN compares false)
    pnew = (float)pv[tid];
    pfin = isfinite(pv[tid]);
  }
  fin = __syncthreads_and(fin);
  yconv = __syncthreads_and(yconv);
  pfin = __syncthreads_and(pfin);
  if (tid < n) {
    state->y[tid] = ye;
    state->y_eval[tid] = ye - pnew;  // gauss_newton.py:126-127 with step size 1
  }
  if (tid == 0) {
    const float fsv = s_fs;
    state->terminate = (yconv && fsv < 1.f) ? 1 : 0;
    int result = pfin ? RES_OK : RES_SINGULAR;  // lineax: non-finite solution => singular (gauss_newton.py:316-318)
    if (!fin) result = RES_NONFINITE;
    state->result = result;
    state->steps += 1;
    state->rr = s_rr;
    state->fstat = fsv;
    if (state->solver == 0 && !fast) state->fallbacks += 1;
    *fstat_slot = 0.f;                  // the next pass's atomicMax target
    if (seq_counter) *seq_counter += 1u;
  }
}

__global__ void gn_init_kernel(GnState* state, const float* y0, int n, float rtol, float atol, int max_steps, int timeout_ms, int solver,
                               float* fstat_slot, unsigned* sync) {
  __shared__ int fin_s;
  const int tid = threadIdx.x;
  if (tid == 0) fin_s = 1;
  __syncthreads();
  const float v = tid < n ? y0[tid] : 0.f;  // unknowns beyond n are padding: zero, and they stay zero
  state->y[tid] = v;
  state->y_eval[tid] = v;
  if (!isfinite(v)) fin_s = 0;
  __syncthreads();
  if (tid == 0) {
    state->steps = 0;
    state->terminate = 0;
    state->result = fin_s ? RES_OK : RES_NONFINITE;  // _iterate.py:214-219
    state->max_steps = max_steps;
    state->rtol = rtol;
    state->atol = atol;
    state->rr = 0.f;
    state->fstat = 0.f;
    state->timeout_ms = timeout_ms;
    state->solver = solver;
    state->fallbacks = 0;
    state->pad = 0;
    *fstat_slot = 0.f;
    sync[0] = 0;
    sync[1] = 0;
  }
}

// r_new = J y - d (one warp per row); optional fused termination statistic against r_old.
__global__ void __launch_bounds__(256) linear_residual_kernel(const float* __restrict__ J, const float* __restrict__ y,
                                                              const float* __restrict__ d, int64_t m, int n, float* r_new,
                                                              const float* r_old, float rtol, float atol, unsigned* fmax_bits) {
  __shared__ float ys[256];
  for (int i = threadIdx.x; i < n; i += blockDim.x) ys[i] = y[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float worst = 0.f;
  for (int64_t row = warp; row < m; row += nwarps) {
    const float4* jr = reinterpret_cast<const float4*>(J + row * n);
    float acc = 0.f;
    for (int c = lane; c < n / 4; c += 32) {
      const float4 v = __ldg(jr + c);
      acc += v.x * ys[4 * c] + v.y * ys[4 * c + 1] + v.z * ys[4 * c + 2] + v.w * ys[4 * c + 3];
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float r = acc - d[row];
      if (r_old) {
        const float q = fabsf(r - r_old[row]) / (atol + rtol * fabsf(r));
        worst = nan_max(worst, q);
      }
      r_new[row] = r;
    }
  }
  if (fmax_bits && lane == 0) {
    // non-negative floats (and +NaN, which must win) order like their bit patterns
    atomicMax(fmax_bits, __float_as_uint(worst != worst ? __int_as_float(0x7fc00000) : worst));
  }
}

}  // namespace b200opt

using namespace b200opt;

extern "C" {

int32_t b200opt_gn_solve(const float* gram_out, int32_t n, double lambda, float* p_out, int32_t* status_out, void* workspace,
                         void* stream) {
  return b200opt_gn_solve_parts(gram_out, 1, 0, n, lambda, p_out, status_out, nullptr, workspace, stream);
}

static int32_t gn_solve_impl(const float* parts, int32_t nparts, int64_t part_stride, int32_t n, double lambda, float* p_out,
                             int32_t* status_out, float* summary_out, void* workspace, void* stream, const unsigned* xchg_flags,
                             unsigned xchg_seq) {
  const float* gram_out = parts;
  if (n < 1 || n > 256 || !gram_out || !p_out || !workspace || nparts < 1 || (nparts > 1 && part_stride < (int64_t)n * n + n + 2)) {
    set_error("gn_solve: need 1 <= n <= 256, nparts >= 1, part_stride >= n*n + n + 2 and non-NULL buffers");
    return B200OPT_EINVAL;
  }
  // the fp64 scratch sits after the per-CTA partials of b200opt_gn_gram in the same workspace
  double* A = reinterpret_cast<double*>(static_cast<char*>(workspace) + (size_t)256 * (256 * 256 + 256 + 4) * 4);
  static LaunchCache cache;
  if (cache.configure(gn_solve_kernel, 1024, SOLVE_SMEM) < 0) return B200OPT_ECUDA;
  gn_solve_kernel<<<1, 1024, SOLVE_SMEM, static_cast<cudaStream_t>(stream)>>>(gram_out, nparts, part_stride, n, lambda, A, p_out, status_out,
                                                                              summary_out, xchg_flags, xchg_seq);
  return check_cuda(cudaGetLastError(), "gn_solve_kernel launch");
}

int32_t b200opt_gn_solve_parts(const float* parts, int32_t nparts, int64_t part_stride, int32_t n, double lambda, float* p_out,
                               int32_t* status_out, float* summary_out, void* workspace, void* stream) {
  return gn_solve_impl(parts, nparts, part_stride, n, lambda, p_out, status_out, summary_out, workspace, stream, nullptr, 0);
}

int32_t b200opt_gn_solve_xchg(const void* xchg_local, int32_t world, int32_t parity, uint32_t seq, int32_t n, double lambda,
                              float* p_out, int32_t* status_out, float* summary_out, void* workspace, void* stream) {
  if (!xchg_local || world < 1 || world > 64 || (parity != 0 && parity != 1) || n != 256) {
    set_error("gn_solve_xchg: bad arguments");
    return B200OPT_EINVAL;
  }
  const int64_t stride = (int64_t)n * n + n + 4;  // gram::XCHG_STRIDE
  const unsigned char* base = static_cast<const unsigned char*>(xchg_local);
  const float* parts = reinterpret_cast<const float*>(base) + (size_t)parity * world * stride;
  const unsigned* flags = reinterpret_cast<const unsigned*>(base + (size_t)2 * world * stride * sizeof(float)) + parity * world;
  return gn_solve_impl(parts, world, stride, n, lambda, p_out, status_out, summary_out, workspace, stream, flags, seq);
}

int32_t b200opt_gn_solve_fast(const float* parts, int32_t nparts, int64_t part_stride, int32_t n, double lambda, int32_t solver,
                              float* p_out, int32_t* status_out, float* summary_out, int32_t* info_out, void* workspace, void* stream) {
  if (n != 256 || !parts || !p_out || !workspace || nparts < 1 || part_stride < (int64_t)n * n + n + 2 || (solver != 0 && solver != 1) ||
      (reinterpret_cast<uintptr_t>(parts) & 15) || (part_stride & 3)) {
    set_error("gn_solve_fast: need n == 256, nparts >= 1, part_stride >= n*n + n + 2 (a multiple of 4), 16-byte aligned parts");
    return B200OPT_EINVAL;
  }
  double* A = reinterpret_cast<double*>(static_cast<char*>(workspace) + gn_workspace_solve_offset());
  const size_t smem = FAST_SMEM > SOLVE_SMEM ? FAST_SMEM : SOLVE_SMEM;
  static LaunchCache cache;
  if (cache.configure(gn_solve_fast_kernel, 1024, smem) < 0) return B200OPT_ECUDA;
  gn_solve_fast_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(parts, nparts, part_stride, lambda, A, solver, p_out, status_out,
                                                                            summary_out, info_out);
  return check_cuda(cudaGetLastError(), "gn_solve_fast_kernel launch");
}

int64_t b200opt_gn_state_size(void) { return (int64_t)sizeof(GnState); }

int32_t b200opt_gn_init(void* state, const float* y0, int32_t n, double rtol, double atol, int32_t max_steps, int32_t solver,
                        int32_t timeout_ms, float* r_old0, int64_t m_local, float* out, void* workspace, void* stream) {
  if (!state || !y0 || !out || !workspace || !r_old0 || n < 4 || n > 256 || (n & 3) || max_steps < 0 || m_local < 0 || (solver != 0 && solver != 1)) {
    set_error("gn_init: need 4 <= n <= 256 with n %% 4 == 0, non-NULL buffers, max_steps >= 0 and solver in {0, 1}");
    return B200OPT_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // gauss_newton.py:229-230: the carry starts from a zero-filled f_info, so the first f-space test compares against zeros
  if (m_local && check_cuda(cudaMemsetAsync(r_old0, 0, sizeof(float) * (size_t)m_local, st), "gn_init memset")) return B200OPT_ECUDA;
  unsigned* sync = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + gn_workspace_sync_offset());
  gn_init_kernel<<<1, 256, 0, st>>>(static_cast<GnState*>(state), y0, n, (float)rtol, (float)atol, max_steps, timeout_ms, solver,
                                    out + 256 * 256 + 256 + 1, sync);
  return check_cuda(cudaGetLastError(), "gn_init_kernel launch");
}

int32_t b200opt_gn_iterate(const float* J, const float* d, float* r0, float* r1, int64_t m_local, int32_t n, void* state, float* out,
                           void* workspace, const void* const* xchg_bases_dev, void* xchg_local, int32_t world, int32_t rank,
                           double rtol, double atol, double lambda, int32_t fold_reduce, void* stream) {
  if (!J || !d || !r0 || !r1 || !state || !out || !workspace || n < 4 || n > 256 || (n & 3) || m_local < 4) {
    set_error("gn_iterate: need 4 <= n <= 256 with n %% 4 == 0, m_local >= 4 and non-NULL buffers");
    return B200OPT_EINVAL;
  }
  if ((xchg_bases_dev != nullptr) != (xchg_local != nullptr) || (xchg_bases_dev && (world < 1 || world > 64 || rank < 0 || rank >= world))) {
    set_error("gn_iterate: inconsistent exchange arguments");
    return B200OPT_EINVAL;
  }
  if ((reinterpret_cast<uintptr_t>(r0) & 15) || (reinterpret_cast<uintptr_t>(r1) & 15)) {
    set_error("gn_iterate: the residual buffers must be 16-byte aligned");
    return B200OPT_EINVAL;
  }
  if (int32_t e = gn_iterate_gram(J, d, r0, r1, m_local, n, state, out, workspace, xchg_bases_dev, xchg_local, world, rank, (float)rtol,
                                  (float)atol, fold_reduce, stream))
    return e;
  double* A = reinterpret_cast<double*>(static_cast<char*>(workspace) + gn_workspace_solve_offset());
  const size_t smem = FAST_SMEM > SOLVE_SMEM ? FAST_SMEM : SOLVE_SMEM;
  static LaunchCache cache;
  if (cache.configure(gn_step_kernel, 1024, smem) < 0) return B200OPT_ECUDA;
  const int64_t stride = 256 * 256 + 256 + 4;  // gram::XCHG_STRIDE
  const float* xparts = nullptr;
  const unsigned* flags = nullptr;
  unsigned* seqc = nullptr;
  if (xchg_local) {
    unsigned char* base = static_cast<unsigned char*>(xchg_local);
    xparts = reinterpret_cast<const float*>(base);
    unsigned* tail = reinterpret_cast<unsigned*>(base + (size_t)2 * world * stride * sizeof(float));
    flags = tail;
    seqc = tail + 2 * world + 1;
  }
  gn_step_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(out, xparts, xchg_local ? world : 1, stride, lambda, A,
                                                                      static_cast<GnState*>(state), out + 256 * 256 + 256 + 1, flags, seqc);
  return check_cuda(cudaGetLastError(), "gn_step_kernel launch");
}

int32_t b200opt_linear_residual(const float* J, const float* y, const float* d, int64_t m_local, int32_t n, float* r,
                                const float* r_old, float rtol, float atol, uint32_t* fmax_bits, void* stream) {
  if (n < 4 || n > 256 || (n & 3) || m_local < 0 || (m_local > 0 && (!J || !y || !d || !r))) {
    set_error("linear_residual: need 4 <= n <= 256, n %% 4 == 0 and non-NULL buffers");
    return B200OPT_EINVAL;
  }
  if (m_local == 0) return 0;
  int64_t blocks = (m_local + 7) / 8;
  const int64_t cap = (int64_t)148 * 8;
  if (blocks > cap) blocks = cap;
  linear_residual_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(J, y, d, m_local, n, r, r_old, rtol, atol,
                                                                                         fmax_bits);
  return check_cuda(cudaGetLastError(), "linear_residual_kernel launch");
}

}  // extern "C"
