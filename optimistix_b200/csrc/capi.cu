// C ABI of the B200 Optimistix hot path (declarations + reference citations: include/b200opt.h).
#include <chrono>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <vector>

#include "lsq_launch.cuh"

namespace b200opt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return B200OPT_ECUDA;
}

int device_sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}

static const b200opt_model_info_t kModels[MODEL_COUNT] = {
    {MODEL_EXP_DECAY, B200OPT_KIND_LEAST_SQUARES, 3, 3, 512, 0, "exp_decay", "y = [a, b, c]; r_i = a*exp(-b*t_i) + c - d_i",
     "t[m] abscissa grid", "d[m] data"},
    {MODEL_GAUSS2, B200OPT_KIND_LEAST_SQUARES, 8, 8, 512, 0, "gauss2",
     "y = [A1, mu1, s1, A2, mu2, s2, b0, b1]; r_i = A1 exp(-(x_i-mu1)^2/(2 s1^2)) + A2 exp(-(x_i-mu2)^2/(2 s2^2)) + b0 + b1 x_i - d_i",
     "x[m] abscissa grid", "d[m] data"},
    {MODEL_IMPLICIT_EULER, B200OPT_KIND_ROOT, 32, 32, 0, 0, "implicit_euler",
     "F(y) = y - y_prev - dt*(A y + y*(B y)), n = 32", "A[n][n] then B[n][n], row-major", "y_prev[n] then dt"},
    {MODEL_ROSENBROCK, B200OPT_KIND_MINIMISE, 0, 64, 0, 0, "rosenbrock",
     "f(y) = sum_i 100 (y_{i+1} - y_i^2)^2 + (1 - y_i)^2 (standard chained Rosenbrock), 2 <= n <= 64", "none", "none"},
    {MODEL_LINEAR_ROWS, B200OPT_KIND_LEAST_SQUARES, 256, 256, 0, 0, "linear_rows",
     "r = J y - d with J materialised row-major in HBM (single large problem; b200opt_gn_gram path)", "J[m][n]", "d[m]"},
};

static const char* kResultNames[] = {"successful", "max_steps_reached", "singular", "breakdown", "stagnation",
                                     "nonlinear_max_steps_reached", "nonlinear_divergence", "nonfinite"};

static int validate_common(const char* who, int dtype, int64_t batch, const b200opt_options_t* opts) {
  if (dtype != B200OPT_F32 && dtype != B200OPT_F64) {
    set_error("%s: unknown dtype %d", who, dtype);
    return B200OPT_EINVAL;
  }
  if (batch < 0) {
    set_error("%s: negative batch", who);
    return B200OPT_EINVAL;
  }
  if (!opts) {
    set_error("%s: opts is NULL", who);
    return B200OPT_EINVAL;
  }
  if (opts->norm < 0 || opts->norm > 2) {
    set_error("%s: unknown norm %d", who, opts->norm);
    return B200OPT_EINVAL;
  }
  if (opts->max_steps < 0) {
    set_error("%s: max_steps must be >= 0", who);
    return B200OPT_EINVAL;
  }
  return 0;
}

// ---- chunked host pipeline shared by the *_host entry points --------------------------------------------------
struct HostPipe {
  cudaStream_t st[2] = {nullptr, nullptr};
  std::vector<void*> allocs;
  ~HostPipe() {
    for (auto s : st)
      if (s) cudaStreamDestroy(s);
    for (auto a : allocs) cudaFree(a);
  }
  void* alloc(size_t bytes) {
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    allocs.push_back(ptr);
    return ptr;
  }
};

}  // namespace b200opt

using namespace b200opt;

extern "C" {

int32_t b200opt_version(void) { return B200OPT_VERSION; }
const char* b200opt_last_error(void) { return g_err; }
const char* b200opt_result_name(int32_t c) { return (c >= 0 && c < 8) ? kResultNames[c] : "unknown"; }
int32_t b200opt_model_count(void) { return MODEL_COUNT; }
int32_t b200opt_model_info_get(int32_t id, b200opt_model_info_t* out) {
  if (id < 0 || id >= MODEL_COUNT || !out) {
    set_error("model_info_get: bad model id %d", id);
    return B200OPT_EINVAL;
  }
  *out = kModels[id];
  return 0;
}
int32_t b200opt_model_id(const char* name) {
  if (!name) return -1;
  for (int i = 0; i < MODEL_COUNT; ++i)
    if (!strcmp(kModels[i].name, name)) return i;
  return -1;
}
void b200opt_default_options(b200opt_options_t* o) {
  if (!o) return;
  o->rtol = 1e-8;
  o->atol = 1e-8;
  o->norm = B200OPT_NORM_MAX;
  o->max_steps = 256;
  o->cauchy_termination = 1;
  o->root_near_zero = 0;
  o->kappa = 1e-2;
}

int32_t b200opt_least_squares(int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n, int32_t m,
                              const void* y0, const void* shared_args, const void* problem_args,
                              int64_t problem_args_stride, const b200opt_options_t* opts, void* y_out,
                              int32_t* num_steps, int32_t* result, void* step_size_out, int32_t* num_accepted_out,
                              void* stream) {
  if (int e = validate_common("least_squares", dtype, batch, opts)) return e;
  if (solver != B200OPT_SOLVER_LEVENBERG_MARQUARDT && solver != B200OPT_SOLVER_GAUSS_NEWTON) {
    set_error("least_squares: solver %d is not a least-squares solver", solver);
    return B200OPT_EINVAL;
  }
  if (m < 1 || problem_args_stride < m) {
    set_error("least_squares: need m >= 1 and problem_args_stride >= m (m=%d stride=%lld)", m, (long long)problem_args_stride);
    return B200OPT_EINVAL;
  }
  if (batch == 0) return 0;
  if (!y0 || !shared_args || !problem_args || !y_out || !num_steps || !result) {
    set_error("least_squares: NULL buffer");
    return B200OPT_EINVAL;
  }
  LsqArgs a{solver, dtype, batch, n, m, y0, shared_args, problem_args, problem_args_stride, opts, y_out,
            num_steps, result, step_size_out, num_accepted_out, static_cast<cudaStream_t>(stream)};
  switch (model_id) {
    case MODEL_EXP_DECAY: return lsq_launch_exp_decay(a);
    case MODEL_GAUSS2: return lsq_launch_gauss2(a);
    default:
      set_error("least_squares: model %d is not a batched least-squares family", model_id);
      return B200OPT_EINVAL;
  }
}

int32_t b200opt_least_squares_host(int32_t device, int32_t solver, int32_t model_id, int32_t dtype, int64_t batch,
                                   int32_t n, int32_t m, const void* y0, const void* shared_args,
                                   int64_t shared_args_count, const void* problem_args, int64_t problem_args_stride,
                                   const b200opt_options_t* opts, void* y_out, int32_t* num_steps, int32_t* result,
                                   int64_t chunk) {
  if (int e = validate_common("least_squares_host", dtype, batch, opts)) return e;
  if (batch == 0) return 0;
  if (check_cuda(cudaSetDevice(device), "cudaSetDevice")) return B200OPT_ECUDA;
  const bool timing = getenv("B200OPT_HOST_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  const size_t es = dtype == B200OPT_F32 ? 4 : 8;
  if (chunk <= 0) chunk = 1 << 16;
  if (chunk > batch) chunk = batch;
  HostPipe hp;
  void* d_shared = hp.alloc(es * shared_args_count);
  void *d_y0[2], *d_prob[2], *d_y[2];
  int32_t *d_steps[2], *d_res[2];
  bool ok = d_shared != nullptr;
  for (int s = 0; s < 2 && ok; ++s) {
    ok = ok && cudaStreamCreateWithFlags(&hp.st[s], cudaStreamNonBlocking) == cudaSuccess;
    d_y0[s] = hp.alloc(es * chunk * n);
    d_prob[s] = hp.alloc(es * chunk * problem_args_stride);
    d_y[s] = hp.alloc(es * chunk * n);
    d_steps[s] = static_cast<int32_t*>(hp.alloc(4 * chunk));
    d_res[s] = static_cast<int32_t*>(hp.alloc(4 * chunk));
    ok = ok && d_y0[s] && d_prob[s] && d_y[s] && d_steps[s] && d_res[s];
  }
  if (!ok) {
    set_error("least_squares_host: device allocation failed");
    return B200OPT_ECUDA;
  }
  if (check_cuda(cudaMemcpy(d_shared, shared_args, es * shared_args_count, cudaMemcpyHostToDevice), "H2D shared"))
    return B200OPT_ECUDA;
  const double t_alloc = now();
  const char* h_y0 = static_cast<const char*>(y0);
  const char* h_prob = static_cast<const char*>(problem_args);
  char* h_y = static_cast<char*>(y_out);
  int c = 0;
  for (int64_t off = 0; off < batch; off += chunk, ++c) {
    const int s = c & 1;
    const int64_t nb = (batch - off < chunk) ? batch - off : chunk;
    cudaMemcpyAsync(d_y0[s], h_y0 + es * off * n, es * nb * n, cudaMemcpyHostToDevice, hp.st[s]);
    cudaMemcpyAsync(d_prob[s], h_prob + es * off * problem_args_stride, es * nb * problem_args_stride,
                    cudaMemcpyHostToDevice, hp.st[s]);
    int e = b200opt_least_squares(solver, model_id, dtype, nb, n, m, d_y0[s], d_shared, d_prob[s], problem_args_stride,
                                  opts, d_y[s], d_steps[s], d_res[s], nullptr, nullptr, hp.st[s]);
    if (e) return e;
    cudaMemcpyAsync(h_y + es * off * n, d_y[s], es * nb * n, cudaMemcpyDeviceToHost, hp.st[s]);
    cudaMemcpyAsync(num_steps + off, d_steps[s], 4 * nb, cudaMemcpyDeviceToHost, hp.st[s]);
    cudaMemcpyAsync(result + off, d_res[s], 4 * nb, cudaMemcpyDeviceToHost, hp.st[s]);
  }
  const double t_enq = now();
  for (int s = 0; s < 2; ++s)
    if (check_cuda(cudaStreamSynchronize(hp.st[s]), "least_squares_host sync")) return B200OPT_ECUDA;
  if (timing) {
    cudaPointerAttributes pa{};
    cudaPointerGetAttributes(&pa, problem_args);
    fprintf(stderr, "[b200opt host] alloc+shared %.2f ms, enqueue %.2f ms, drain %.2f ms, chunks %d, prob memtype %d (1=host-pinned, 0=unregistered)\n",
            t_alloc - t_begin, t_enq - t_alloc, now() - t_enq, c, (int)pa.type);
  }
  return 0;
}

}  // extern "C"
