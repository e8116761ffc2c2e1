// C ABI of the B200 Optimistix hot path (declarations + reference citations: include/b200opt.h).
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "lsq_launch.cuh"

namespace b200opt {

static thread_local char g_err[512] = "";
thread_local unsigned long long* g_queue_scratch = nullptr;  // see QueueCounter (common.cuh)

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

int jit_family_count();
int jit_family_id(const char* name);
bool jit_family_info(int index, b200opt_model_info_t* out);
int lsq_launch_jit(const LsqArgs& a, int model_id);

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
    {MODEL_LINEAR_ROWS, B200OPT_KIND_LEAST_SQUARES, 0, 256, 0, 0, "linear_rows",
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
  if (opts->linear_solver != B200OPT_LINSOLVE_NORMAL_CHOLESKY && opts->linear_solver != B200OPT_LINSOLVE_QR) {
    set_error("%s: unknown linear_solver %d", who, opts->linear_solver);
    return B200OPT_EINVAL;
  }
  if (opts->data_dtype < B200OPT_DATA_AS_COMPUTE || opts->data_dtype > B200OPT_DATA_I16) {
    set_error("%s: unknown data_dtype %d", who, opts->data_dtype);
    return B200OPT_EINVAL;
  }
  return 0;
}

// ---- chunked host pipeline shared by the *_host entry points --------------------------------------------------
// kStreams streams, each with its own device staging buffers: chunk c+1's H2D overlaps chunk c's kernel and chunk
// c-1's D2H; a ring deeper than 2 keeps the copy engines busy while the host drains a finished slot.  Page-locked caller memory is copied directly; pageable OUTPUT arrays go through page-locked bounce
// buffers so that no D2H ever blocks the enqueueing thread (a blocking copy would serialise the pipeline).
// The staging buffers are cached per host thread (grow-only) and released by b200opt_host_cache_release().
constexpr int kStreams = 4;

// Pageable -> page-locked staging copies done by a few helper threads.  cudaMemcpyAsync from pageable memory goes through the
// driver's own single staging buffer (measured 11 GB/s: 100 ms per 1 Mi gauss2 fits for a caller who did not pin anything);
// copying each chunk into our page-locked slab with several threads first and letting the copy engine read from there keeps
// the pipeline at the host-memory rate instead.  One pool per host thread that uses the host entry points, created on first
// use with pageable inputs, joined by b200opt_host_cache_release().
class CopyPool {
 public:
  explicit CopyPool(int nthreads) : stop_(false), pending_(0), generation_(0) {
    for (int t = 0; t < nthreads; ++t) workers_.emplace_back([this, t] { run(t); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int size() const { return (int)workers_.size(); }
  // dst[0..bytes) = src[0..bytes), split over the pool and the calling thread; returns when every byte has been copied
  void copy(void* dst, const void* src, size_t bytes) {
    const int parts = size() + 1;
    if (bytes < (size_t)1 << 20 || parts == 1) {
      memcpy(dst, src, bytes);
      return;
    }
    const size_t per = ((bytes / parts) + 4095) & ~size_t(4095);
    {
      std::lock_guard<std::mutex> lk(mu_);
      dst_ = static_cast<char*>(dst), src_ = static_cast<const char*>(src), bytes_ = bytes, per_ = per;
      pending_ = size();
      ++generation_;
    }
    cv_.notify_all();
    const size_t off = per * size();  // the caller takes the last slice
    if (off < bytes) memcpy(dst_ + off, src_ + off, bytes - off);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void run(int t) {
    unsigned long long seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return generation_ != seen; });
      seen = generation_;
      if (stop_) return;
      const size_t off = per_ * t;
      char* d = dst_;
      const char* s = src_;
      const size_t n = off < bytes_ ? (bytes_ - off < per_ ? bytes_ - off : per_) : 0;
      lk.unlock();
      if (n) memcpy(d + off, s + off, n);
      lk.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  bool stop_;
  int pending_;
  unsigned long long generation_;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0, per_ = 0;
};

struct HostCache {
  int device = -1;
  cudaStream_t st[kStreams] = {};
  cudaEvent_t done[kStreams] = {};
  void* dev[kStreams] = {};
  size_t dev_bytes = 0;
  void* pin[kStreams] = {};
  size_t pin_bytes = 0;
  void* pin_in[kStreams] = {};   // page-locked staging of PAGEABLE inputs (filled by the copy pool)
  size_t pin_in_bytes = 0;
  cudaEvent_t in_done[kStreams] = {};  // the H2D copies out of pin_in[s] have completed: the slab may be refilled
  CopyPool* pool = nullptr;
  void* shared = nullptr;
  size_t shared_bytes = 0;
  void release() {
    if (device < 0) return;
    int current = -1;
    cudaGetDevice(&current);
    cudaSetDevice(device);  // the staging buffers are freed on the device that owns them ...
    for (int s = 0; s < kStreams; ++s) {
      if (st[s]) cudaStreamDestroy(st[s]);
      if (done[s]) cudaEventDestroy(done[s]);
      if (dev[s]) cudaFree(dev[s]);
      if (pin[s]) cudaFreeHost(pin[s]);
      if (pin_in[s]) cudaFreeHost(pin_in[s]);
      if (in_done[s]) cudaEventDestroy(in_done[s]);
      st[s] = nullptr, done[s] = nullptr, dev[s] = nullptr, pin[s] = nullptr, pin_in[s] = nullptr, in_done[s] = nullptr;
    }
    delete pool;
    pool = nullptr;
    pin_in_bytes = 0;
    if (shared) cudaFree(shared);
    shared = nullptr;
    dev_bytes = pin_bytes = shared_bytes = 0;
    device = -1;
    if (current >= 0) cudaSetDevice(current);  // ... and the caller's current device is left as it was
  }
  ~HostCache() { /* process teardown: the CUDA context may already be gone; leak rather than crash */ }
};
static thread_local HostCache g_host;

struct HostIn {
  const void* h;
  size_t bytes_per_problem;
};
struct HostOut {
  void* h;
  size_t bytes_per_problem;
};

static bool is_pinned(const void* ptr) {
  cudaPointerAttributes pa{};
  if (cudaPointerGetAttributes(&pa, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeManaged;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

// launch(nb, d_shared, d_in[], d_out[], stream) enqueues the solve of one chunk.
template <class Launch>
static int host_pipeline(const char* who, int device, int64_t batch, int64_t chunk, const void* shared_h, size_t shared_bytes,
                         const std::vector<HostIn>& ins, const std::vector<HostOut>& outs, Launch launch) {
  struct DeviceGuard {  // the host entry points work on `device` and leave the caller's current device untouched
    int prev = -1;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() {
      if (prev >= 0) cudaSetDevice(prev);
    }
  } guard;
  if (check_cuda(cudaSetDevice(device), "cudaSetDevice")) return B200OPT_ECUDA;
  HostCache& hc = g_host;
  if (hc.device != device) {
    hc.release();
    if (check_cuda(cudaSetDevice(device), "cudaSetDevice")) return B200OPT_ECUDA;  // streams and buffers below belong to `device`
    hc.device = device;
    for (int s = 0; s < kStreams; ++s) {
      if (check_cuda(cudaStreamCreateWithFlags(&hc.st[s], cudaStreamNonBlocking), "stream create")) return B200OPT_ECUDA;
      if (check_cuda(cudaEventCreateWithFlags(&hc.done[s], cudaEventDisableTiming), "event create")) return B200OPT_ECUDA;
      if (check_cuda(cudaEventCreateWithFlags(&hc.in_done[s], cudaEventDisableTiming), "event create")) return B200OPT_ECUDA;
    }
  }
  if (chunk <= 0) chunk = 1 << 15;
  if (chunk > batch) chunk = batch;
  // carve one device slab and one pinned slab per stream
  std::vector<size_t> in_off(ins.size()), out_off(outs.size()), bounce_off(outs.size());
  std::vector<bool> out_pinned(outs.size());
  std::vector<bool> in_staged(ins.size());
  std::vector<size_t> stage_off(ins.size());
  size_t dev_need = 0, pin_need = 0, pin_in_need = 0;
  for (size_t i = 0; i < ins.size(); ++i) {
    in_off[i] = dev_need, dev_need += align256(ins[i].bytes_per_problem * chunk);
    // a pageable input that is worth it (>= 1 MiB per chunk) is staged through page-locked memory by the copy pool
    in_staged[i] = ins[i].h != nullptr && ins[i].bytes_per_problem * (size_t)chunk >= ((size_t)1 << 20) && !is_pinned(ins[i].h);
    if (in_staged[i]) stage_off[i] = pin_in_need, pin_in_need += align256(ins[i].bytes_per_problem * chunk);
  }
  const size_t queue_off = dev_need;  // 8 bytes per stream for the ticket counter of the dynamic problem queue
  dev_need += 256;
  for (size_t i = 0; i < outs.size(); ++i) {
    out_off[i] = dev_need, dev_need += align256(outs[i].bytes_per_problem * chunk);
    out_pinned[i] = is_pinned(outs[i].h);
    if (!out_pinned[i]) bounce_off[i] = pin_need, pin_need += align256(outs[i].bytes_per_problem * chunk);
  }
  if (dev_need > hc.dev_bytes) {
    for (int s = 0; s < kStreams; ++s) {
      if (hc.dev[s]) cudaFree(hc.dev[s]);
      hc.dev[s] = nullptr;
      if (check_cuda(cudaMalloc(&hc.dev[s], dev_need), "staging cudaMalloc")) return hc.dev_bytes = 0, B200OPT_ECUDA;
    }
    hc.dev_bytes = dev_need;
  }
  if (pin_need > hc.pin_bytes) {
    for (int s = 0; s < kStreams; ++s) {
      if (hc.pin[s]) cudaFreeHost(hc.pin[s]);
      hc.pin[s] = nullptr;
      if (check_cuda(cudaMallocHost(&hc.pin[s], pin_need), "bounce cudaMallocHost")) return hc.pin_bytes = 0, B200OPT_ECUDA;
    }
    hc.pin_bytes = pin_need;
  }
  if (pin_in_need > hc.pin_in_bytes) {
    for (int s = 0; s < kStreams; ++s) {
      if (hc.pin_in[s]) cudaFreeHost(hc.pin_in[s]);
      hc.pin_in[s] = nullptr;
      if (check_cuda(cudaMallocHost(&hc.pin_in[s], pin_in_need), "input staging cudaMallocHost")) return hc.pin_in_bytes = 0, B200OPT_ECUDA;
    }
    hc.pin_in_bytes = pin_in_need;
  }
  if (pin_in_need && !hc.pool) {
    unsigned hw = std::thread::hardware_concurrency();
    int helpers = hw > 16 ? 7 : (hw > 4 ? (int)hw / 2 - 1 : 1);  // the calling thread copies too
    hc.pool = new CopyPool(helpers);
  }
  bool in_busy[kStreams] = {};
  if (shared_bytes > hc.shared_bytes) {
    if (hc.shared) cudaFree(hc.shared);
    hc.shared = nullptr;
    if (check_cuda(cudaMalloc(&hc.shared, shared_bytes), "shared cudaMalloc")) return hc.shared_bytes = 0, B200OPT_ECUDA;
    hc.shared_bytes = shared_bytes;
  }
  if (shared_bytes) {
    if (check_cuda(cudaMemcpyAsync(hc.shared, shared_h, shared_bytes, cudaMemcpyHostToDevice, hc.st[0]), "H2D shared"))
      return B200OPT_ECUDA;
    cudaEventRecord(hc.done[0], hc.st[0]);
    for (int s = 1; s < kStreams; ++s) cudaStreamWaitEvent(hc.st[s], hc.done[0], 0);
  }
  struct Pending {
    int64_t off = -1, nb = 0;
  } pending[kStreams];
  auto drain = [&](int s) -> int {  // copy stream s's finished bounce buffers into the caller's pageable arrays
    if (pending[s].off < 0) return 0;
    if (check_cuda(cudaEventSynchronize(hc.done[s]), who)) return B200OPT_ECUDA;
    for (size_t i = 0; i < outs.size(); ++i)
      if (!out_pinned[i])
        memcpy(static_cast<char*>(outs[i].h) + outs[i].bytes_per_problem * pending[s].off,
               static_cast<char*>(hc.pin[s]) + bounce_off[i], outs[i].bytes_per_problem * pending[s].nb);
    pending[s].off = -1;
    return 0;
  };
  std::vector<const void*> d_in(ins.size());
  std::vector<void*> d_out(outs.size());
  int c = 0;
  for (int64_t off = 0; off < batch; off += chunk, ++c) {
    const int s = c % kStreams;
    const int64_t nb = (batch - off < chunk) ? batch - off : chunk;
    if (int e = drain(s)) return e;  // the bounce buffers of this stream are about to be reused
    char* base = static_cast<char*>(hc.dev[s]);
    bool staged_any = false;
    for (size_t i = 0; i < ins.size(); ++i) {
      d_in[i] = base + in_off[i];
      if (!ins[i].bytes_per_problem || !ins[i].h) continue;  // h == NULL: device-only scratch of that size per problem
      const char* src = static_cast<const char*>(ins[i].h) + ins[i].bytes_per_problem * off;
      if (in_staged[i]) {
        if (in_busy[s] && !staged_any) {  // the copy engine may still be reading this slot's previous contents
          if (check_cuda(cudaEventSynchronize(hc.in_done[s]), who)) return B200OPT_ECUDA;
        }
        staged_any = true;
        char* stage = static_cast<char*>(hc.pin_in[s]) + stage_off[i];
        hc.pool->copy(stage, src, ins[i].bytes_per_problem * nb);
        src = stage;
      }
      cudaMemcpyAsync(base + in_off[i], src, ins[i].bytes_per_problem * nb, cudaMemcpyHostToDevice, hc.st[s]);
    }
    if (staged_any) {
      cudaEventRecord(hc.in_done[s], hc.st[s]);
      in_busy[s] = true;
    }
    for (size_t i = 0; i < outs.size(); ++i) d_out[i] = base + out_off[i];
    g_queue_scratch = reinterpret_cast<unsigned long long*>(base + queue_off);
    const int launch_rc = launch(nb, hc.shared, d_in.data(), d_out.data(), hc.st[s]);
    g_queue_scratch = nullptr;
    if (launch_rc) return launch_rc;
    for (size_t i = 0; i < outs.size(); ++i) {
      void* dst = out_pinned[i] ? static_cast<void*>(static_cast<char*>(outs[i].h) + outs[i].bytes_per_problem * off)
                                : static_cast<void*>(static_cast<char*>(hc.pin[s]) + bounce_off[i]);
      cudaMemcpyAsync(dst, d_out[i], outs[i].bytes_per_problem * nb, cudaMemcpyDeviceToHost, hc.st[s]);
    }
    cudaEventRecord(hc.done[s], hc.st[s]);
    pending[s].off = off, pending[s].nb = nb;
  }
  for (int s = 0; s < kStreams; ++s) {
    if (int e = drain(s)) return e;
    if (check_cuda(cudaStreamSynchronize(hc.st[s]), who)) return B200OPT_ECUDA;
  }
  return check_cuda(cudaGetLastError(), who);
}

// rows [nb][stride] from up to four per-problem part arrays (b200opt_root_find_host_parts); unused trailing elements are zero
struct PackParts {
  const void* src[4];
  int elems[4];
  int n;
};
template <typename T> __global__ void pack_parts_kernel(T* dst, int64_t stride, int64_t nb, PackParts pp) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nb * stride) return;
  const int64_t prob = e / stride;
  int j = (int)(e - prob * stride);
  T v = T(0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q < pp.n) {
      if (j >= 0 && j < pp.elems[q]) v = static_cast<const T*>(pp.src[q])[prob * pp.elems[q] + j];
      j -= pp.elems[q];
    }
  }
  dst[e] = v;
}
}  // namespace b200opt

using namespace b200opt;

extern "C" {

int32_t b200opt_version(void) { return B200OPT_VERSION; }
const char* b200opt_last_error(void) { return g_err; }
const char* b200opt_result_name(int32_t c) { return (c >= 0 && c < 8) ? kResultNames[c] : "unknown"; }
int32_t b200opt_model_count(void) { return MODEL_COUNT + jit_family_count(); }
int32_t b200opt_model_info_get(int32_t id, b200opt_model_info_t* out) {
  if (out && id >= B200OPT_JIT_MODEL_BASE && jit_family_info(id - B200OPT_JIT_MODEL_BASE, out)) return 0;
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
  return jit_family_id(name);
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
  o->linear_solver = B200OPT_LINSOLVE_NORMAL_CHOLESKY;
  o->data_dtype = B200OPT_DATA_AS_COMPUTE;
  o->scratch = nullptr;
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
      if (model_id >= B200OPT_JIT_MODEL_BASE) return lsq_launch_jit(a, model_id);  // run-time compiled family
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
  if (!y0 || !shared_args || !problem_args || !y_out || !num_steps || !result) {
    set_error("least_squares_host: NULL buffer");
    return B200OPT_EINVAL;
  }
  const size_t es = dtype == B200OPT_F32 ? 4 : 8;
  const size_t ds = opts->data_dtype == B200OPT_DATA_AS_COMPUTE ? es : 2;  // 16-bit data rows: half (a quarter) of the bytes per problem
  return host_pipeline(
      "least_squares_host", device, batch, chunk, shared_args, es * shared_args_count,
      {{y0, es * n}, {problem_args, ds * (size_t)problem_args_stride}}, {{y_out, es * n}, {num_steps, 4}, {result, 4}},
      [&](int64_t nb, const void* d_shared, const void** d_in, void** d_out, cudaStream_t st) {
        return b200opt_least_squares(solver, model_id, dtype, nb, n, m, d_in[0], d_shared, d_in[1], problem_args_stride, opts,
                                     d_out[0], static_cast<int32_t*>(d_out[1]), static_cast<int32_t*>(d_out[2]), nullptr,
                                     nullptr, st);
      });
}

void b200opt_host_cache_release(void) { g_host.release(); }

int32_t b200opt_minimise_host(int32_t device, int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n,
                              const void* y0, const void* shared_args, int64_t shared_args_count, const void* problem_args,
                              int64_t problem_args_stride, const b200opt_options_t* opts, void* y_out, int32_t* num_steps,
                              int32_t* result, int64_t chunk) {
  (void)shared_args, (void)shared_args_count, (void)problem_args, (void)problem_args_stride;
  if (int e = validate_common("minimise_host", dtype, batch, opts)) return e;
  if (batch == 0) return 0;
  if (!y0 || !y_out || !num_steps || !result) {
    set_error("minimise_host: NULL buffer");
    return B200OPT_EINVAL;
  }
  const size_t es = dtype == B200OPT_F32 ? 4 : 8;
  return host_pipeline("minimise_host", device, batch, chunk, nullptr, 0, {{y0, es * n}},
                       {{y_out, es * n}, {num_steps, 4}, {result, 4}},
                       [&](int64_t nb, const void*, const void** d_in, void** d_out, cudaStream_t st) {
                         return b200opt_minimise(solver, model_id, dtype, nb, n, d_in[0], nullptr, nullptr, 0, opts, d_out[0],
                                                 static_cast<int32_t*>(d_out[1]), static_cast<int32_t*>(d_out[2]), nullptr,
                                                 nullptr, nullptr, st);
                       });
}

int32_t b200opt_root_find_host(int32_t device, int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n,
                               const void* y0, const void* shared_args, int64_t shared_args_count, const void* problem_args,
                               int64_t problem_args_stride, const b200opt_options_t* opts, void* y_out, int32_t* num_steps,
                               int32_t* result, int64_t chunk) {
  if (int e = validate_common("root_find_host", dtype, batch, opts)) return e;
  if (batch == 0) return 0;
  if (!y0 || !shared_args || !problem_args || !y_out || !num_steps || !result) {
    set_error("root_find_host: NULL buffer");
    return B200OPT_EINVAL;
  }
  const size_t es = dtype == B200OPT_F32 ? 4 : 8;
  return host_pipeline(
      "root_find_host", device, batch, chunk, shared_args, es * shared_args_count,
      {{y0, es * n}, {problem_args, es * (size_t)problem_args_stride}}, {{y_out, es * n}, {num_steps, 4}, {result, 4}},
      [&](int64_t nb, const void* d_shared, const void** d_in, void** d_out, cudaStream_t st) {
        return b200opt_root_find(solver, model_id, dtype, nb, n, d_in[0], d_shared, d_in[1], problem_args_stride, nullptr,
                                 nullptr, opts, d_out[0], static_cast<int32_t*>(d_out[1]), static_cast<int32_t*>(d_out[2]), st);
      });
}

int32_t b200opt_root_find_host_parts(int32_t device, int32_t solver, int32_t model_id, int32_t dtype, int64_t batch, int32_t n,
                                     const void* y0, const void* shared_args, int64_t shared_args_count, const void* const* parts,
                                     const int32_t* part_elems, int32_t n_parts, int64_t problem_args_stride,
                                     const b200opt_options_t* opts, void* y_out, int32_t* num_steps, int32_t* result, int64_t chunk) {
  if (int e = validate_common("root_find_host_parts", dtype, batch, opts)) return e;
  if (batch == 0) return 0;
  if (!y0 || !shared_args || !parts || !part_elems || !y_out || !num_steps || !result || n_parts < 1 || n_parts > 4) {
    set_error("root_find_host_parts: NULL buffer or n_parts outside 1..4");
    return B200OPT_EINVAL;
  }
  int64_t total = 0;
  for (int q = 0; q < n_parts; ++q) {
    if (!parts[q] || part_elems[q] < 1) {
      set_error("root_find_host_parts: part %d is NULL or empty", q);
      return B200OPT_EINVAL;
    }
    total += part_elems[q];
  }
  if (total > problem_args_stride) {
    set_error("root_find_host_parts: the parts hold %lld elements per problem, the stride is %lld", (long long)total,
              (long long)problem_args_stride);
    return B200OPT_EINVAL;
  }
  const size_t es = dtype == B200OPT_F32 ? 4 : 8;
  std::vector<HostIn> ins = {{y0, es * n}, {nullptr, es * (size_t)problem_args_stride}};  // [1]: the packed rows, device only
  for (int q = 0; q < n_parts; ++q) ins.push_back({parts[q], es * (size_t)part_elems[q]});
  return host_pipeline(
      "root_find_host_parts", device, batch, chunk, shared_args, es * shared_args_count, ins, {{y_out, es * n}, {num_steps, 4}, {result, 4}},
      [&](int64_t nb, const void* d_shared, const void** d_in, void** d_out, cudaStream_t st) {
        PackParts pp;
        pp.n = n_parts;
        for (int q = 0; q < 4; ++q) pp.src[q] = q < n_parts ? d_in[2 + q] : nullptr, pp.elems[q] = q < n_parts ? part_elems[q] : 0;
        void* rows = const_cast<void*>(d_in[1]);
        const int64_t cells = nb * problem_args_stride;
        const unsigned grid = (unsigned)((cells + 255) / 256);
        if (dtype == B200OPT_F64) pack_parts_kernel<double><<<grid, 256, 0, st>>>(static_cast<double*>(rows), problem_args_stride, nb, pp);
        else pack_parts_kernel<float><<<grid, 256, 0, st>>>(static_cast<float*>(rows), problem_args_stride, nb, pp);
        if (check_cuda(cudaGetLastError(), "pack_parts_kernel launch")) return (int)B200OPT_ECUDA;
        return (int)b200opt_root_find(solver, model_id, dtype, nb, n, d_in[0], d_shared, rows, problem_args_stride, nullptr, nullptr, opts,
                                      d_out[0], static_cast<int32_t*>(d_out[1]), static_cast<int32_t*>(d_out[2]), st);
      });
}

}  // extern "C"
