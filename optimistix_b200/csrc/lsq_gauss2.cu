#include <stdlib.h>

#include "lsq_kernel_p2.cuh"
#include "lsq_launch.cuh"
namespace b200opt {

// packed-fp32 fast path (lsq_kernel_p2.cuh); everything it does not cover goes to the scalar kernel
template <class PModel, bool TRUST, int NPAIRS> static int lsq_launch_p2(const LsqParams<float>& p, cudaStream_t stream) {
  auto kern = lsq_solve_kernel_p2<PModel, TRUST, NPAIRS>;
  const size_t per_warp = WarpReduceLayout<float, Tri<PModel::N>::size + PModel::N>::elems + (size_t)(3 + PModel::NAUX) * p.m_pad;
  const size_t smem = sizeof(float) * ((size_t)p.m_pad + kLsqWarpsPerBlock * per_warp);
  if (smem > 227 * 1024) return -1;  // caller falls back
  static LaunchCache cache;
  const int blocks_per_sm = cache.configure(kern, kLsqWarpsPerBlock * 32, smem);
  if (blocks_per_sm < 0) return B200OPT_ECUDA;
  const int64_t need = (p.batch + kLsqWarpsPerBlock - 1) / kLsqWarpsPerBlock;
  const int64_t cap = (int64_t)device_sm_count() * blocks_per_sm;
  const int grid = (int)(need < cap ? need : cap);
  QueueCounter qc;
  LsqParams<float> pq = p;
  pq.queue = nullptr;
  if (need > 4 * cap) {  // dynamic problem queue after the first resident wave (common.cuh); not worth its set-up for a wave or two
    qc.acquire(stream, p.queue);  // p.queue carries the caller's scratch (options.scratch), or NULL
    pq.queue = qc.ptr;
  }
  kern<<<grid, kLsqWarpsPerBlock * 32, smem, stream>>>(pq);
  return check_cuda(cudaGetLastError(), "lsq_solve_kernel_p2 launch");
}

int lsq_launch_gauss2(const LsqArgs& a) {
  const bool d16 = a.opts->data_dtype != B200OPT_DATA_AS_COMPUTE;  // 16-bit rows: any alignment (the loader has a scalar path)
  const bool aligned = d16 || ((a.prob_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.prob) % 16) == 0));
  if (a.opts->linear_solver == B200OPT_LINSOLVE_NORMAL_CHOLESKY && a.dtype == B200OPT_F32 && a.n == Gauss2P2::N && a.m % 64 == 0 && aligned && a.opts->norm == B200OPT_NORM_MAX &&
      !a.opts->root_near_zero) {
    LsqParams<float> p;
    p.y0 = static_cast<const float*>(a.y0);
    p.shared = static_cast<const float*>(a.shared);
    p.prob = static_cast<const float*>(a.prob);
    p.prob_stride = a.prob_stride;
    p.y_out = static_cast<float*>(a.y_out);
    p.num_steps = a.num_steps;
    p.result = a.result;
    p.step_size_out = static_cast<float*>(a.step_size_out);
    p.num_accepted = a.num_accepted;
    p.batch = a.batch;
    p.m = a.m;
    p.m_pad = a.m;
    p.rtol = (float)a.opts->rtol;
    p.atol = (float)a.opts->atol;
    p.norm = a.opts->norm;
    p.max_steps = a.opts->max_steps;
    p.root_near_zero = 0;
    p.vec_ok = 1;
    p.data_kind = a.opts->data_dtype;
    p.queue = static_cast<unsigned long long*>(a.opts->scratch);
    p.data_vec = (a.m % 8 == 0) && (a.prob_stride % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.prob) % 16) == 0);
    const bool lm = a.solver == B200OPT_SOLVER_LEVENBERG_MARQUARDT;
    int rc;
    if (a.m == 256) rc = lm ? lsq_launch_p2<Gauss2P2, true, 4>(p, a.stream) : lsq_launch_p2<Gauss2P2, false, 4>(p, a.stream);  // configs[1]
    else rc = lm ? lsq_launch_p2<Gauss2P2, true, 0>(p, a.stream) : lsq_launch_p2<Gauss2P2, false, 0>(p, a.stream);
    if (rc >= 0) return rc;
  }
  return lsq_launch_model<Gauss2>(a);
}
}  // namespace b200opt
