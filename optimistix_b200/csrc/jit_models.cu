// Run-time registration of residual-model families (the "plugin registry" of the north_star, user side).
//
// The compiled-in families (models.cuh) cover BASELINE.json's configs.  A user family is the SAME kind of object - a struct
// of __device__ functions `prepare / residual / jac_row` (see models.cuh) - given as CUDA C++ source text.  It is compiled
// here with NVRTC for sm_100a against the very headers the library was built from (lsq_kernel.cuh), loaded with the driver
// API, and from then on `b200opt_least_squares[_host]` runs it in the same persistent one-warp-per-problem kernel as the
// built-in families: the reference's `fn(y, args)` becomes device code instead of a per-iteration round trip through the
// generic step kernel.  NVRTC is dlopen()ed on first use, so the library itself has no link-time dependency on it.
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "lsq_launch.cuh"
#include "newton_kernel.cuh"

namespace b200opt {

namespace {

struct Nvrtc {
  void* handle = nullptr;
  decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
  decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
  decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
  decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
  decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
  decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
  decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
  decltype(&nvrtcAddNameExpression) AddNameExpression = nullptr;
  decltype(&nvrtcGetLoweredName) GetLoweredName = nullptr;
  decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
  bool load() {
    if (handle) return true;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (handle) break;
    }
    if (!handle) return false;
#define B200OPT_SYM(f) f = reinterpret_cast<decltype(f)>(dlsym(handle, "nvrtc" #f))
    B200OPT_SYM(CreateProgram);
    B200OPT_SYM(CompileProgram);
    B200OPT_SYM(DestroyProgram);
    B200OPT_SYM(GetProgramLogSize);
    B200OPT_SYM(GetProgramLog);
    B200OPT_SYM(GetCUBINSize);
    B200OPT_SYM(GetCUBIN);
    B200OPT_SYM(AddNameExpression);
    B200OPT_SYM(GetLoweredName);
    B200OPT_SYM(GetErrorString);
#undef B200OPT_SYM
    return CreateProgram && CompileProgram && DestroyProgram && GetProgramLogSize && GetProgramLog && GetCUBINSize && GetCUBIN &&
           AddNameExpression && GetLoweredName && GetErrorString;
  }
};

struct JitFamily {
  const char* name = "";  // strdup'ed once, never freed: b200opt_model_info_get hands the pointer out
  int kind = B200OPT_KIND_LEAST_SQUARES;
  int n = 0, naux = 0, shared_elems = 0;
  CUmodule module = nullptr;
  CUfunction fn[2][2] = {};  // least squares: [dtype f32/f64][GN/LM]; root: [dtype][0]
};

// Driver API entry points, resolved through the runtime (cudaGetDriverEntryPoint) so that libb200opt.so does not link
// against libcuda.so.1 and still loads on a machine without a driver (the CPU-side symbol test relies on that).
struct Driver {
  bool ready = false;
  decltype(&cuModuleLoadData) ModuleLoadData = nullptr;
  decltype(&cuModuleGetFunction) ModuleGetFunction = nullptr;
  decltype(&cuModuleUnload) ModuleUnload = nullptr;
  decltype(&cuFuncSetAttribute) FuncSetAttribute = nullptr;
  decltype(&cuOccupancyMaxActiveBlocksPerMultiprocessor) OccupancyMaxActiveBlocksPerMultiprocessor = nullptr;
  decltype(&cuLaunchKernel) LaunchKernel = nullptr;
  decltype(&cuGetErrorString) GetErrorString = nullptr;
  template <class F> static bool get(const char* name, F& f) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
    f = reinterpret_cast<F>(sym);
    return f != nullptr;
  }
  bool load() {
    if (ready) return true;
    ready = get("cuModuleLoadData", ModuleLoadData) && get("cuModuleGetFunction", ModuleGetFunction) &&
            get("cuModuleUnload", ModuleUnload) && get("cuFuncSetAttribute", FuncSetAttribute) &&
            get("cuOccupancyMaxActiveBlocksPerMultiprocessor", OccupancyMaxActiveBlocksPerMultiprocessor) &&
            get("cuLaunchKernel", LaunchKernel) && get("cuGetErrorString", GetErrorString);
    return ready;
  }
};

std::mutex g_mu;
Nvrtc g_nvrtc;
Driver g_drv;
std::vector<JitFamily> g_families;

int check_cu(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return 0;
  const char* msg = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &msg);
  set_error("CUDA driver error in %s: %s", what, msg ? msg : "?");
  return B200OPT_ECUDA;
}

// WarpReduceLayout<T, K>::elems for run-time K (lsq_kernel.cuh)
size_t reduce_scratch_elems(size_t esize, int K) {
  const int V = (int)(16 / esize);
  const int KQ = (K + V - 1) / V, KPAD = KQ * V;
  const int KP = KPAD + V * ((KQ % 2 == 0) ? 1 : 0);
  return (size_t)32 * KP + KPAD;
}

}  // namespace

int jit_family_count() {
  std::lock_guard<std::mutex> lk(g_mu);
  return (int)g_families.size();
}

int jit_family_id(const char* name) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t i = 0; i < g_families.size(); ++i)
    if (!strcmp(g_families[i].name, name)) return B200OPT_JIT_MODEL_BASE + (int)i;
  return -1;
}

bool jit_family_info(int index, b200opt_model_info_t* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (index < 0 || index >= (int)g_families.size()) return false;
  const JitFamily& f = g_families[index];
  if (f.kind == B200OPT_KIND_ROOT)
    *out = b200opt_model_info_t{B200OPT_JIT_MODEL_BASE + index, B200OPT_KIND_ROOT, f.n, f.n, 0, f.shared_elems, f.name,
                                "run-time compiled family (b200opt_register_root_family)", "shared[shared_args] staged in shared memory",
                                "prob[stride] per problem"};
  else
    *out = b200opt_model_info_t{B200OPT_JIT_MODEL_BASE + index, B200OPT_KIND_LEAST_SQUARES, f.n, f.n, 0, 0, f.name,
                                "run-time compiled family (b200opt_register_curve_family)", "x[m] abscissa grid", "d[m] data"};
  return true;
}

int lsq_launch_jit(const LsqArgs& a, int model_id) {
  JitFamily fam;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    const int idx = model_id - B200OPT_JIT_MODEL_BASE;
    if (idx < 0 || idx >= (int)g_families.size()) {
      set_error("least_squares: unknown run-time model id %d", model_id);
      return B200OPT_EINVAL;
    }
    fam = g_families[idx];
  }
  if (fam.kind != B200OPT_KIND_LEAST_SQUARES) {
    set_error("least_squares: model %s is not a least-squares family", fam.name);
    return B200OPT_EINVAL;
  }
  if (a.n != fam.n) {
    set_error("least_squares: model %s expects n=%d parameters, got n=%d", fam.name, fam.n, a.n);
    return B200OPT_EINVAL;
  }
  if (a.opts->linear_solver != B200OPT_LINSOLVE_NORMAL_CHOLESKY) {
    set_error("least_squares: run-time compiled families run the normal-equation kernel (linear_solver = QR is compiled for the built-in families)");
    return B200OPT_EINVAL;
  }
  const bool f64 = a.dtype == B200OPT_F64;
  const size_t es = f64 ? 8 : 4;
  const int m_pad = (a.m + 31) / 32 * 32;
  const int K = fam.n * (fam.n + 1) / 2 + fam.n;
  const size_t per_warp = reduce_scratch_elems(es, K) + (size_t)(3 + fam.naux) * m_pad;
  const size_t smem = es * ((size_t)m_pad + kLsqWarpsPerBlock * per_warp);
  if (smem > 227 * 1024) {
    set_error("least_squares: m=%d needs %zu bytes of shared memory per block (limit 227 KB)", a.m, smem);
    return B200OPT_EINVAL;
  }
  CUfunction fn = fam.fn[f64 ? 1 : 0][a.solver == B200OPT_SOLVER_LEVENBERG_MARQUARDT ? 1 : 0];
  if (check_cu(g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem), "cuFuncSetAttribute")) return B200OPT_ECUDA;
  int occ = 0;
  if (check_cu(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kLsqWarpsPerBlock * 32, smem), "occupancy")) return B200OPT_ECUDA;
  const int64_t need = (a.batch + kLsqWarpsPerBlock - 1) / kLsqWarpsPerBlock;
  const int64_t cap = (int64_t)device_sm_count() * (occ > 0 ? occ : 1);
  const unsigned grid = (unsigned)(need < cap ? need : cap);
  // LsqParams<T> is the kernel's single by-value parameter; the two instantiations share the code below
  auto launch = [&](auto tag) -> int {
    using T = decltype(tag);
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
    p.m_pad = m_pad;
    p.rtol = (T)a.opts->rtol;
    p.atol = (T)a.opts->atol;
    p.norm = a.opts->norm;
    p.max_steps = a.opts->max_steps;
    p.root_near_zero = a.opts->root_near_zero;
    constexpr int V = Num<T>::kVec;
    p.vec_ok = (a.m % V == 0) && (a.prob_stride % V == 0) && ((reinterpret_cast<uintptr_t>(a.prob) % (V * sizeof(T))) == 0);
    p.data_kind = a.opts->data_dtype;
    p.data_vec = (a.m % 8 == 0) && (a.prob_stride % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.prob) % 16) == 0);
    QueueCounter qc;
    p.queue = nullptr;
    if (need > 4 * cap) {
      qc.acquire(a.stream, a.opts->scratch);
      p.queue = qc.ptr;
    }
    void* params[] = {&p};
    return check_cu(g_drv.LaunchKernel(fn, grid, 1, 1, kLsqWarpsPerBlock * 32, 1, 1, (unsigned)smem, a.stream, params, nullptr),
                    "cuLaunchKernel(run-time family)");
  };
  return f64 ? launch(double{}) : launch(float{});
}

int root_launch_jit(int model_id, int solver, int dtype, int64_t batch, int n, const void* y0, const void* shared, const void* prob,
                    int64_t stride, const void* lower, const void* upper, const b200opt_options_t* o, void* y_out, int32_t* num_steps,
                    int32_t* result, cudaStream_t stream) {
  JitFamily fam;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    const int idx = model_id - B200OPT_JIT_MODEL_BASE;
    if (idx < 0 || idx >= (int)g_families.size() || g_families[idx].kind != B200OPT_KIND_ROOT) {
      set_error("root_find: model %d is not a registered root-finding family", model_id);
      return B200OPT_EINVAL;
    }
    fam = g_families[idx];
  }
  if (n != fam.n) {
    set_error("root_find: model %s expects n=%d, got n=%d", fam.name, fam.n, n);
    return B200OPT_EINVAL;
  }
  const bool f64 = dtype == B200OPT_F64;
  const size_t es = f64 ? 8 : 4;
  const size_t smem = es * fam.shared_elems + kRootWarpsPerBlock * (es * (2 * fam.n + 4) + 32);  // root_smem_bytes
  CUfunction fn = fam.fn[f64 ? 1 : 0][0];
  if (check_cu(g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem), "cuFuncSetAttribute")) return B200OPT_ECUDA;
  int occ = 0;
  if (check_cu(g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kRootWarpsPerBlock * 32, smem), "occupancy")) return B200OPT_ECUDA;
  const int64_t need = (batch + kRootWarpsPerBlock - 1) / kRootWarpsPerBlock;
  const int64_t cap = (int64_t)device_sm_count() * (occ > 0 ? occ : 1);
  const unsigned grid = (unsigned)(need < cap ? need : cap);
  auto launch = [&](auto tag) -> int {
    using T = decltype(tag);
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
    void* params[] = {&p};
    return check_cu(g_drv.LaunchKernel(fn, grid, 1, 1, kRootWarpsPerBlock * 32, 1, 1, (unsigned)smem, stream, params, nullptr),
                    "cuLaunchKernel(run-time root family)");
  };
  return f64 ? launch(double{}) : launch(float{});
}

namespace {

// compile `src`, load it, resolve `nexpr` name expressions into fam.fn (row-major over [2][2]); shared by both kinds
int compile_and_load(const char* name, const std::string& src, const char* csrc_dir, const char* const* exprs, int nexpr, JitFamily& fam) {
  if (!g_nvrtc.load()) {
    set_error("register family: libnvrtc could not be loaded (%s)", dlerror() ? dlerror() : "symbols missing");
    return B200OPT_ENOTIMPL;
  }
  // make sure the primary context exists and is current (the rest of the library uses the runtime API)
  if (check_cuda(cudaFree(nullptr), "context")) return B200OPT_ECUDA;
  if (!g_drv.load()) {
    set_error("register family: the CUDA driver entry points could not be resolved");
    return B200OPT_ECUDA;
  }
  nvrtcProgram prog = nullptr;
  nvrtcResult r = g_nvrtc.CreateProgram(&prog, src.c_str(), (std::string(name) + ".cu").c_str(), 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) {
    set_error("register family: nvrtcCreateProgram: %s", g_nvrtc.GetErrorString(r));
    return B200OPT_ECUDA;
  }
  for (int i = 0; i < nexpr; ++i) g_nvrtc.AddNameExpression(prog, exprs[i]);
  const std::string inc = std::string("-I") + csrc_dir;
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", inc.c_str()};
  r = g_nvrtc.CompileProgram(prog, 4, opts);
  if (r != NVRTC_SUCCESS) {
    size_t n_log = 0;
    g_nvrtc.GetProgramLogSize(prog, &n_log);
    std::string log(n_log, '\0');
    if (n_log) g_nvrtc.GetProgramLog(prog, &log[0]);
    set_error("register family (%s): compilation failed: %.400s", name, log.c_str());
    g_nvrtc.DestroyProgram(&prog);
    return B200OPT_EINVAL;
  }
  size_t n_bin = 0;
  g_nvrtc.GetCUBINSize(prog, &n_bin);
  std::vector<char> cubin(n_bin);
  g_nvrtc.GetCUBIN(prog, cubin.data());
  if (check_cu(g_drv.ModuleLoadData(&fam.module, cubin.data()), "cuModuleLoadData")) {
    g_nvrtc.DestroyProgram(&prog);
    return B200OPT_ECUDA;
  }
  for (int i = 0; i < nexpr; ++i) {
    const char* lowered = nullptr;
    g_nvrtc.GetLoweredName(prog, exprs[i], &lowered);
    CUfunction* slot = &fam.fn[0][0] + (nexpr == 4 ? i : 2 * i);
    if (!lowered || check_cu(g_drv.ModuleGetFunction(slot, fam.module, lowered), "cuModuleGetFunction")) {
      g_nvrtc.DestroyProgram(&prog);
      g_drv.ModuleUnload(fam.module);
      return B200OPT_ECUDA;
    }
  }
  g_nvrtc.DestroyProgram(&prog);
  return 0;
}

bool name_taken(const char* name) {
  for (size_t i = 0; i < g_families.size(); ++i)
    if (!strcmp(g_families[i].name, name)) return true;
  return false;
}

}  // namespace

}  // namespace b200opt

using namespace b200opt;

extern "C" {

int32_t b200opt_register_root_family(const char* name, int32_t n, int32_t shared_elems, const char* struct_body, const char* csrc_dir) {
  if (!name || !struct_body || !csrc_dir || n < 4 || n > 32 || (n & 3) || shared_elems < 0 || shared_elems > 24 * 1024) {
    set_error("register_root_family: need a name, a struct body, the csrc directory, n in {4, 8, ..., 32} and 0 <= shared_elems <= 24576");
    return B200OPT_EINVAL;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  if (name_taken(name)) {
    set_error("register_root_family: a family named %s is already registered", name);
    return B200OPT_EINVAL;
  }
  char head[1024];
  snprintf(head, sizeof(head),
           "#include \"newton_kernel.cuh\"\nnamespace b200opt {\ntemplate <typename T> struct UserRoot {\n  static constexpr int N = %d;\n"
           "  static constexpr int kSharedElems = %d;\n  struct Ctx { const T* prob; };\n"
           "  __device__ static __forceinline__ void stage_shared(const T* g, T* s, int tid, int nthreads) {\n"
           "    for (int e = tid; e < kSharedElems; e += nthreads) s[e] = g[e];\n  }\n"
           "  __device__ static __forceinline__ Ctx load(const T* prob, int) { return Ctx{prob}; }\n",
           n, shared_elems);
  const std::string src = std::string(head) + struct_body + "\n};\n}  // namespace b200opt\n";
  static const char* kExpr[2] = {"b200opt::root_solve_kernel<b200opt::UserRoot<float>, float>",
                                 "b200opt::root_solve_kernel<b200opt::UserRoot<double>, double>"};
  JitFamily fam;
  fam.kind = B200OPT_KIND_ROOT;
  fam.n = n;
  fam.shared_elems = shared_elems;
  if (int e = compile_and_load(name, src, csrc_dir, kExpr, 2, fam)) return e;
  fam.name = strdup(name);
  g_families.push_back(fam);
  return B200OPT_JIT_MODEL_BASE + (int)g_families.size() - 1;
}

int32_t b200opt_register_curve_family(const char* name, int32_t n, int32_t naux, const char* struct_body, const char* csrc_dir) {
  if (!name || !struct_body || !csrc_dir || n < 1 || n > 16 || naux < 1 || naux > 8) {
    set_error("register_curve_family: need a name, a struct body, the csrc directory, 1 <= n <= 16 and 1 <= naux <= 8");
    return B200OPT_EINVAL;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  if (name_taken(name)) {
    set_error("register_curve_family: a family named %s is already registered", name);
    return B200OPT_EINVAL;
  }
  char head[512];
  snprintf(head, sizeof(head),
           "#include \"lsq_kernel.cuh\"\nnamespace b200opt {\ntemplate <typename T> struct UserModel {\n  static constexpr int N = %d;\n"
           "  static constexpr int NAUX = %d;\n",
           n, naux);
  const std::string src = std::string(head) + struct_body + "\n};\n}  // namespace b200opt\n";
  static const char* kExpr[4] = {
      "b200opt::lsq_solve_kernel<b200opt::UserModel<float>, float, false>", "b200opt::lsq_solve_kernel<b200opt::UserModel<float>, float, true>",
      "b200opt::lsq_solve_kernel<b200opt::UserModel<double>, double, false>",
      "b200opt::lsq_solve_kernel<b200opt::UserModel<double>, double, true>"};
  JitFamily fam;
  fam.kind = B200OPT_KIND_LEAST_SQUARES;
  fam.n = n;
  fam.naux = naux;
  if (int e = compile_and_load(name, src, csrc_dir, kExpr, 4, fam)) return e;
  fam.name = strdup(name);
  g_families.push_back(fam);
  return B200OPT_JIT_MODEL_BASE + (int)g_families.size() - 1;
}

}  // extern "C"
