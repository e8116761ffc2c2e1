// Host-side launcher for the batched LM/GN kernel: picks rows-per-lane, grid size and shared memory.
#pragma once
#include "lsq_kernel.cuh"

namespace b200opt {

struct LsqArgs {
  int solver, dtype;
  int64_t batch;
  int n, m;
  const void *y0, *shared, *prob;
  int64_t prob_stride;
  const b200opt_options_t* opts;
  void* y_out;
  int32_t *num_steps, *result;
  void* step_size_out;
  int32_t* num_accepted;
  cudaStream_t stream;
};

int device_sm_count();

template <class Model, typename T, bool TRUST, bool QRMODE = false> int lsq_launch_one(const LsqParams<T>& p, cudaStream_t stream) {
  auto kern = lsq_solve_kernel<Model, T, TRUST, QRMODE>;
  const size_t smem = lsq_smem_bytes<Model, T>(p.m_pad, QRMODE);
  if (smem > 227 * 1024) {
    set_error("least_squares: m=%d needs %zu bytes of shared memory per block (limit 227 KB)", p.m, smem);
    return B200OPT_EINVAL;
  }
  static LaunchCache cache;  // per instantiation; per device inside
  const int blocks_per_sm = cache.configure(kern, kLsqWarpsPerBlock * 32, smem);
  if (blocks_per_sm < 0) return B200OPT_ECUDA;
  const int64_t need = (p.batch + kLsqWarpsPerBlock - 1) / kLsqWarpsPerBlock;
  const int64_t cap = (int64_t)device_sm_count() * blocks_per_sm;  // persistent: one resident wave, then the problem queue
  const int grid = (int)(need < cap ? need : cap);
  QueueCounter qc;
  LsqParams<T> pq = p;
  pq.queue = nullptr;
  if (need > 4 * cap) {  // several waves of problems: later ones are handed out dynamically (a launch of a wave or two runs
                         // ~0.1 ms: the stream-ordered counter allocation would cost more than the balance gains)
    qc.acquire(stream, p.queue);  // p.queue carries the caller's scratch (options.scratch), or NULL
    pq.queue = qc.ptr;
  }
  kern<<<grid, kLsqWarpsPerBlock * 32, smem, stream>>>(pq);
  return check_cuda(cudaGetLastError(), "lsq_solve_kernel launch");
}

template <class Model, typename T, bool TRUST> int lsq_launch_rpl(const LsqParams<T>& p, cudaStream_t stream) {
  return lsq_launch_one<Model, T, TRUST>(p, stream);
}

template <template <typename> class ModelT, typename T> int lsq_launch_typed(const LsqArgs& a) {
  using Model = ModelT<T>;
  if (a.n != Model::N) {
    set_error("least_squares: model expects n=%d parameters, got n=%d", Model::N, a.n);
    return B200OPT_EINVAL;
  }
  LsqParams<T> p;
  p.y0 = static_cast<const T*>(a.y0);
  p.shared = static_cast<const T*>(a.shared);
  p.prob = static_cast<const T*>(a.prob);
  p.prob_stride = a.prob_stride;
  p.y_out = static_cast<T*>(a.y_out);
  p.num_steps = a.num_steps;
  p.result = a.result;
  p.step_size_out = static_cast<T*>(a.step_size_out);
  p.num_accepted = a.num_accepted;
  p.batch = a.batch;
  p.m = a.m;
  p.m_pad = (a.m + 31) / 32 * 32;
  p.rtol = (T)a.opts->rtol;
  p.atol = (T)a.opts->atol;
  p.norm = a.opts->norm;
  p.max_steps = a.opts->max_steps;
  p.root_near_zero = a.opts->root_near_zero;
  constexpr int V = Num<T>::kVec;
  p.vec_ok = (a.m % V == 0) && (a.prob_stride % V == 0) && ((reinterpret_cast<uintptr_t>(a.prob) % (V * sizeof(T))) == 0);
  p.data_kind = a.opts->data_dtype;
  p.queue = static_cast<unsigned long long*>(a.opts->scratch);
  p.data_vec = (a.m % 8 == 0) && (a.prob_stride % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.prob) % 16) == 0);
  const bool lm = a.solver == B200OPT_SOLVER_LEVENBERG_MARQUARDT;
  if (a.opts->linear_solver == B200OPT_LINSOLVE_QR)  // the reference's own linear algebra (see lsq_kernel.cuh, QR mode)
    return lm ? lsq_launch_one<Model, T, true, true>(p, a.stream) : lsq_launch_one<Model, T, false, true>(p, a.stream);
  if (lm) return lsq_launch_rpl<Model, T, true>(p, a.stream);
  return lsq_launch_rpl<Model, T, false>(p, a.stream);
}

template <template <typename> class ModelT> int lsq_launch_model(const LsqArgs& a) {
  if (a.dtype == B200OPT_F32) return lsq_launch_typed<ModelT, float>(a);
  if (a.dtype == B200OPT_F64) return lsq_launch_typed<ModelT, double>(a);
  set_error("least_squares: unknown dtype %d", a.dtype);
  return B200OPT_EINVAL;
}

// one translation unit per model family (parallel nvcc)
int lsq_launch_exp_decay(const LsqArgs& a);
int lsq_launch_gauss2(const LsqArgs& a);

}  // namespace b200opt
