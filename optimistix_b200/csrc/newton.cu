// Batched Newton / Chord entry points (declarations + reference citations: include/b200opt.h).
#include <vector>

#include "newton_kernel.cuh"

namespace b200opt {
int device_sm_count();
int root_launch_jit(int model_id, int solver, int dtype, int64_t batch, int n, const void* y0, const void* shared, const void* prob,
                    int64_t stride, const void* lower, const void* upper, const b200opt_options_t* o, void* y_out, int32_t* num_steps,
                    int32_t* result, cudaStream_t stream);

template <class Model, typename T>
static int root_launch(int solver, int64_t batch, int n, const void* y0, const void* shared, const void* prob,
                       int64_t stride, const void* lower, const void* upper, const b200opt_options_t* o, void* y_out,
                       int32_t* num_steps, int32_t* result, cudaStream_t stream) {
  if (n != Model::N) {
    set_error("root_find: model expects n=%d, got n=%d", Model::N, n);
    return B200OPT_EINVAL;
  }
  RootParams<T> p;
  p.y0 = static_cast<const T*>(y0);
  p.shared = static_cast<const T*>(shared);
  p.prob = static_cast<const T*>(prob);
  p.prob_stride = stride;
  p.lower = static_cast<const T*>(lower);
  p.upper = static_cast<const T*>(upper);
  p.y_out = static_cast<T*>(y_out);
  p.num_steps = num_steps;
  p.result = result;
  p.batch = batch;
  p.n = n;
  p.rtol = (T)o->rtol;
  p.atol = (T)o->atol;
  p.kappa = (T)o->kappa;
  p.norm = o->norm;
  p.max_steps = o->max_steps;
  p.cauchy = o->cauchy_termination;
  p.chord = solver == B200OPT_SOLVER_CHORD;
  auto kern = root_solve_kernel<Model, T>;
  const size_t smem = root_smem_bytes<Model, T>();
  static LaunchCache cache;
  const int blocks_per_sm = cache.configure(kern, kRootWarpsPerBlock * 32, smem);
  if (blocks_per_sm < 0) return B200OPT_ECUDA;
  const int64_t need = (batch + kRootWarpsPerBlock - 1) / kRootWarpsPerBlock;
  const int64_t cap = (int64_t)device_sm_count() * blocks_per_sm;
  const int grid = (int)(need < cap ? need : cap);
  kern<<<grid, kRootWarpsPerBlock * 32, smem, stream>>>(p);
  return check_cuda(cudaGetLastError(), "root_solve_kernel launch");
}
}  // namespace b200opt

using namespace b200opt;

extern "C" int32_t b200opt_root_find(int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n,
                                     const void* y0, const void* shared_args, const void* problem_args,
                                     int64_t problem_args_stride, const void* lower, const void* upper,
                                     const b200opt_options_t* opts, void* y_out, int32_t* num_steps, int32_t* result,
                                     void* stream) {
  if (!opts || opts->norm < 0 || opts->norm > 2 || opts->max_steps < 0 || batch < 0) {
    set_error("root_find: bad options or negative batch");
    return B200OPT_EINVAL;
  }
  if (solver != B200OPT_SOLVER_NEWTON && solver != B200OPT_SOLVER_CHORD) {
    set_error("root_find: solver %d is not a root finder", solver);
    return B200OPT_EINVAL;
  }
  if (model_id != MODEL_IMPLICIT_EULER && model_id < B200OPT_JIT_MODEL_BASE) {
    set_error("root_find: model %d is not a root-finding family", model_id);
    return B200OPT_EINVAL;
  }
  if (dtype != B200OPT_F32 && dtype != B200OPT_F64) {
    set_error("root_find: unknown dtype %d", dtype);
    return B200OPT_EINVAL;
  }
  if (model_id >= B200OPT_JIT_MODEL_BASE) {  // run-time compiled family (b200opt_register_root_family)
    if (batch == 0) return 0;
    if (!y0 || !shared_args || !problem_args || !y_out || !num_steps || !result) {
      set_error("root_find: NULL buffer");
      return B200OPT_EINVAL;
    }
    return root_launch_jit(model_id, solver, dtype, batch, n, y0, shared_args, problem_args, problem_args_stride, lower, upper, opts, y_out,
                           num_steps, result, static_cast<cudaStream_t>(stream));
  }
  if (problem_args_stride < n + 1) {
    set_error("root_find: implicit_euler needs problem_args_stride >= n + 1");
    return B200OPT_EINVAL;
  }
  if (batch == 0) return 0;
  if (!y0 || !shared_args || !problem_args || !y_out || !num_steps || !result) {
    set_error("root_find: NULL buffer");
    return B200OPT_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == B200OPT_F64)
    return root_launch<ImplicitEuler<double>, double>(solver, batch, n, y0, shared_args, problem_args, problem_args_stride,
                                                       lower, upper, opts, y_out, num_steps, result, st);
  return root_launch<ImplicitEuler<float>, float>(solver, batch, n, y0, shared_args, problem_args, problem_args_stride, lower,
                                                   upper, opts, y_out, num_steps, result, st);
}
