// Support kernels of the single large Gauss-Newton problem (BASELINE configs[4]):
//   b200opt_gn_solve        (G + lambda I) p = g by blocked Cholesky in fp64, one CTA (n = 256: 5.6 Mflop, barrier-latency bound)
//   b200opt_linear_residual r = J y - d for the "linear_rows" family, one warp per row, coalesced float4 reads,
//                           fused with the f-space half of cauchy_termination (optimistix/_misc.py:306-309):
//                           max_i |r_new_i - r_old_i| / (atol + rtol |r_new_i|)
#include "common.cuh"

namespace b200opt {

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
__global__ void __launch_bounds__(1024, 1) gn_solve_kernel(const float* gram_out, int nparts, int64_t part_stride, int n, double lambda,
                                                           double* A, float* p_out, int32_t* status, float* summary,
                                                           const unsigned* flags, unsigned seq) {
  extern __shared__ double sm[];
  double* Lp = sm;                          // [256][SOLVE_LD] finished panel of the factor (rows c0..n-1 used)
  double* z = Lp + 256 * SOLVE_LD;          // [256] right-hand side -> L^-1 g
  double* rinv = z + 256;                   // [256] 1 / L_kk
  double* pv = rinv + 256;                  // [256] solution
  double* zs = pv + 256;                    // [SOLVE_PW] panel right-hand side of the backward sweep
  double* colbuf = zs + SOLVE_PW;           // [2][SOLVE_PW] unscaled column of the diagonal block (ping-pong)
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  if (flags) {
    int ok = 1;
    if (tid < nparts) {
      const volatile unsigned* f = flags + tid;
      const long long t0 = clock64();
      while (*f != seq) {
        __nanosleep(100);
        if (clock64() - t0 > 4000000000LL) {  // ~2 s: a peer never published
          ok = 0;
          break;
        }
      }
    }
    __threadfence_system();  // acquire: the peers' stores that preceded their flags are visible to the loads below
    ok = __syncthreads_and(ok);
    if (!ok) {
      if (tid == 0 && status) *status = B200OPT_EXCHANGE_TIMEOUT;
      return;
    }
  }
  // lower triangle of G + lambda I in fp64
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
  if (summary && tid == 0) {
    double rr = 0.0;
    float fs = 0.f;
    for (int q = 0; q < nparts; ++q) {
      rr += (double)__ldcg(gram_out + q * part_stride + n * n + n);
      fs = nan_max(fs, __ldcg(gram_out + q * part_stride + n * n + n + 1));
    }
    summary[0] = (float)rr;
    summary[1] = fs;
  }
  __syncthreads();

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
  int ok = 1;
  for (int i = tid; i < n; i += nt) {
    p_out[i] = (float)pv[i];
    if (!isfinite(pv[i])) ok = 0;
  }
  ok = __syncthreads_and(ok);
  if (tid == 0 && status) *status = ok ? RES_OK : RES_SINGULAR;  // lineax: non-finite solution => singular
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
  static thread_local bool attr_set = false;  // per host thread: the entry points are re-entrant
  if (!attr_set) {
    if (check_cuda(cudaFuncSetAttribute(gn_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM),
                   "cudaFuncSetAttribute(gn_solve_kernel)"))
      return B200OPT_ECUDA;
    attr_set = true;
  }
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
