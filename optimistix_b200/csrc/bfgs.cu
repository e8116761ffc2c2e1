// Batched BFGS entry point (declarations + reference citations: include/b200opt.h).
#include "bfgs_kernel.cuh"

namespace b200opt {
int device_sm_count();

template <class Model, typename T>
static int min_launch(int64_t batch, int n, const void* y0, const b200opt_options_t* o, void* y_out, int32_t* num_steps,
                      int32_t* result, void* f_out, void* hinv_out, int32_t* nacc, cudaStream_t stream) {
  MinParams<T> p;
  p.y0 = static_cast<const T*>(y0);
  p.y_out = static_cast<T*>(y_out);
  p.num_steps = num_steps;
  p.result = result;
  p.f_out = static_cast<T*>(f_out);
  p.hinv_out = static_cast<T*>(hinv_out);
  p.num_accepted = nacc;
  p.batch = batch;
  p.n = n;
  p.rtol = (T)o->rtol;
  p.atol = (T)o->atol;
  p.norm = o->norm;
  p.max_steps = o->max_steps;
  auto kern = bfgs_solve_kernel<Model, T>;
  static LaunchCache cache;
  const int blocks_per_sm = cache.configure(kern, kBfgsThreads, 0);
  if (blocks_per_sm < 0) return B200OPT_ECUDA;
  const int64_t cap = (int64_t)device_sm_count() * blocks_per_sm;
  const int grid = (int)(batch < cap ? batch : cap);
  QueueCounter qc;
  p.queue = nullptr;
  if (batch > 4 * cap) {  // dynamic problem queue after the first resident wave (common.cuh)
    qc.acquire(stream, o->scratch);
    p.queue = qc.ptr;
  }
  kern<<<grid, kBfgsThreads, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "bfgs_solve_kernel launch");
}
}  // namespace b200opt

using namespace b200opt;

extern "C" int32_t b200opt_minimise(int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n, const void* y0,
                                    const void* shared_args, const void* problem_args, int64_t problem_args_stride,
                                    const b200opt_options_t* opts, void* y_out, int32_t* num_steps, int32_t* result,
                                    void* f_out, void* hessian_inv_out, int32_t* num_accepted_out, void* stream) {
  (void)shared_args, (void)problem_args, (void)problem_args_stride;
  if (!opts || opts->norm < 0 || opts->norm > 2 || opts->max_steps < 0 || batch < 0) {
    set_error("minimise: bad options or negative batch");
    return B200OPT_EINVAL;
  }
  if (solver != B200OPT_SOLVER_BFGS) {
    set_error("minimise: solver %d is not a minimiser", solver);
    return B200OPT_EINVAL;
  }
  if (model_id != MODEL_ROSENBROCK) {
    set_error("minimise: model %d is not a minimisation family", model_id);
    return B200OPT_EINVAL;
  }
  if (n < 2 || n > kBfgsN) {
    set_error("minimise: rosenbrock needs 2 <= n <= %d, got %d", kBfgsN, n);
    return B200OPT_EINVAL;
  }
  if (dtype != B200OPT_F32 && dtype != B200OPT_F64) {
    set_error("minimise: unknown dtype %d", dtype);
    return B200OPT_EINVAL;
  }
  if (batch == 0) return 0;
  if (!y0 || !y_out || !num_steps || !result) {
    set_error("minimise: NULL buffer");
    return B200OPT_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200OPT_F64)
    return min_launch<Rosenbrock<double>, double>(batch, n, y0, opts, y_out, num_steps, result, f_out, hessian_inv_out,
                                                  num_accepted_out, st);
  return min_launch<Rosenbrock<float>, float>(batch, n, y0, opts, y_out, num_steps, result, f_out, hessian_inv_out,
                                              num_accepted_out, st);
}
