// Support kernels of the single large Gauss-Newton problem (BASELINE configs[4]):
//   b200opt_gn_solve        (G + lambda I) p = g by Cholesky in fp64, one CTA (n = 256: 5.6 Mflop, latency bound)
//   b200opt_linear_residual r = J y - d for the "linear_rows" family, one warp per row, coalesced float4 reads,
//                           fused with the f-space half of cauchy_termination (optimistix/_misc.py:306-309):
//                           max_i |r_new_i - r_old_i| / (atol + rtol |r_new_i|)
#include "common.cuh"

namespace b200opt {

// Right-looking Cholesky on an n x n fp64 matrix held in global memory (L2 resident, 512 KiB at n = 256).
__global__ void __launch_bounds__(1024, 1) gn_solve_kernel(const float* gram_out, int n, double lambda, double* A, float* p_out,
                                                           int32_t* status) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ double colk[256];
  __shared__ double z[256];
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, j = e % n;
    A[e] = (double)gram_out[e] + (i == j ? lambda : 0.0);
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    if (tid == 0) colk[k] = sqrt(A[k * n + k]);  // NaN for a non-positive pivot -> non-finite solution -> singular
    __syncthreads();
    const double dk = colk[k];
    for (int i = k + 1 + tid; i < n; i += nt) colk[i] = A[i * n + k] / dk;
    __syncthreads();
    // trailing update of the lower triangle: A[i][j] -= L[i][k] L[j][k], k < j <= i
    const int rem = n - k - 1;
    for (int e = tid; e < rem * rem; e += nt) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j <= i) A[i * n + j] -= colk[i] * colk[j];
    }
    for (int i = k + tid; i < n; i += nt) A[i * n + k] = colk[i];  // store column k of L
    __syncthreads();
  }
  // forward / backward substitution (sequential in k, parallel over rows)
  for (int i = tid; i < n; i += nt) z[i] = (double)gram_out[n * n + i];
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    if (tid == 0) z[k] = z[k] / A[k * n + k];
    __syncthreads();
    const double zk = z[k];
    for (int i = k + 1 + tid; i < n; i += nt) z[i] -= A[i * n + k] * zk;
    __syncthreads();
  }
  for (int k = n - 1; k >= 0; --k) {
    if (tid == 0) z[k] = z[k] / A[k * n + k];
    __syncthreads();
    const double zk = z[k];
    for (int i = tid; i < k; i += nt) z[i] -= A[k * n + i] * zk;
    __syncthreads();
  }
  int ok = 1;
  for (int i = tid; i < n; i += nt) {
    p_out[i] = (float)z[i];
    if (!isfinite(z[i])) ok = 0;
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
  if (n < 1 || n > 256 || !gram_out || !p_out || !workspace) {
    set_error("gn_solve: need 1 <= n <= 256 and non-NULL buffers");
    return B200OPT_EINVAL;
  }
  // the fp64 scratch sits after the per-CTA partials of b200opt_gn_gram in the same workspace
  double* A = reinterpret_cast<double*>(static_cast<char*>(workspace) + (size_t)256 * (256 * 256 + 256 + 4) * 4);
  gn_solve_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(gram_out, n, lambda, A, p_out, status_out);
  return check_cuda(cudaGetLastError(), "gn_solve_kernel launch");
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
