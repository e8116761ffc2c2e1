// Shared device helpers for the B200 (sm_100a) Optimistix hot path.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#endif

#include "../../include/b200opt.h"

namespace b200opt {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

enum Result : int {
  RES_OK = B200OPT_RESULT_SUCCESSFUL,
  RES_SINGULAR = B200OPT_RESULT_SINGULAR,
  RES_NL_MAX_STEPS = B200OPT_RESULT_NONLINEAR_MAX_STEPS_REACHED,
  RES_DIVERGED = B200OPT_RESULT_NONLINEAR_DIVERGENCE,
  RES_NONFINITE = B200OPT_RESULT_NONFINITE,
};

// ---- numeric traits (jnp.finfo) ---------------------------------------------------------------------------
template <typename T> struct Num;
template <> struct Num<float> {
  static __host__ __device__ constexpr float eps() { return 1.1920928955078125e-07f; }
  static __host__ __device__ constexpr float max() { return 3.4028234663852886e+38f; }
#ifdef __CUDACC_RTC__
  static __device__ float inf() { return __int_as_float(0x7f800000); }
#else
  static __host__ __device__ constexpr float inf() { return __builtin_huge_valf(); }
#endif
  static constexpr int precision = 6;  // jnp.finfo(float32).precision
  using Vec = float4;
  static constexpr int kVec = 4;
};
template <> struct Num<double> {
  static __host__ __device__ constexpr double eps() { return 2.220446049250313e-16; }
  static __host__ __device__ constexpr double max() { return 1.7976931348623157e+308; }
#ifdef __CUDACC_RTC__
  static __device__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
#else
  static __host__ __device__ constexpr double inf() { return __builtin_huge_val(); }
#endif
  static constexpr int precision = 15;
  using Vec = double2;
  static constexpr int kVec = 2;
};

__device__ __forceinline__ float t_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float t_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double t_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ float t_exp(float x) { return expf(x); }
__device__ __forceinline__ double t_exp(double x) { return exp(x); }
__device__ __forceinline__ float t_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double t_abs(double x) { return fabs(x); }
__device__ __forceinline__ float t_tanh(float x) { return tanhf(x); }
__device__ __forceinline__ double t_tanh(double x) { return tanh(x); }
template <typename T> __device__ __forceinline__ bool t_finite(T x) { return isfinite(x); }

// jnp.max semantics: NaN propagates (fmax would drop it and wrongly report convergence).
template <typename T> __device__ __forceinline__ T nan_max(T a, T b) {
  return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}

// ---- warp collectives (deterministic butterfly; every lane ends with the same value) -------------------------
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
template <typename T> __device__ __forceinline__ T warp_nanmax(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// NaN-propagating max of NON-NEGATIVE values (x >= +0, +inf or NaN - what the scaled norms below feed it): such floats
// order like their bit patterns read as unsigned integers, with every NaN above +inf, so the warp maximum is one integer
// REDUX (two for a double: high words, then the low words of the lanes that hold the maximal high word) instead of five
// shuffle rounds of compare-and-select.  Exact, like any max.
__device__ __forceinline__ float warp_nanmax_nonneg(float v) {
  return __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(v)));
}
__device__ __forceinline__ double warp_nanmax_nonneg(double v) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(kFull, hi);
  const unsigned ml = __reduce_max_sync(kFull, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}

// ---- predicted reduction of ClassicalTrustRegion (trust_region.py:182-186) from the carried Gram quantities ------------
// The reference evaluates  0.5 (|J d + r|^2 - |r|^2)  by forming both squared norms in floating point and subtracting
// them: its value is quantised to the spacing of |r|^2 and is exactly 0 once the step is too small to move |r|^2.  The
// accept test `df <= 0.01 pred` near convergence depends on that (df is 0 or +-1 spacing there): with the exactly
// evaluated d.g + 0.5 d.G d < 0 a step with df == 0 is rejected, with the reference's quantity it is accepted, and the
// exact form costs +2-3 % iterations on the LM workloads (measured with the oracle: tests/test_oracle_pred_modes.py).
// So |J d + r|^2 is formed through the Gram identity  r.r + (2 d.g + d.G d)  and then rounded and subtracted exactly as the
// reference does.  `lin` = d.g, `quad` = d.G d, `rr` = r.r.  The adds are explicit (no contraction, no re-association).
__device__ __forceinline__ float predicted_reduction(float lin, float quad, float rr) {
  const float jt = __fadd_rn(rr, __fmaf_rn(2.f, lin, quad));  // |J d + r|^2 at the precision of |r|^2
  return 0.5f * __fsub_rn(jt, rr);
}
__device__ __forceinline__ double predicted_reduction(double lin, double quad, double rr) {
  const double jt = __dadd_rn(rr, __fma_rn(2.0, lin, quad));
  return 0.5 * __dsub_rn(jt, rr);
}

// ---- 16-bit per-problem data (b200opt_options_t.data_dtype) -------------------------------------------------------------
// Detector counts and half-precision samples are what measured curves usually are; shipping them as 16-bit values halves the
// bytes every problem moves over PCIe and HBM.  Each value is converted EXACTLY to the compute type on load (every uint16 /
// int16 / float16 is representable in fp32), so the solve is the solve of the converted array, bit for bit.
template <typename T> __device__ __forceinline__ T data16_to(unsigned short raw, int kind) {
  if (kind == B200OPT_DATA_U16) return (T)raw;
  if (kind == B200OPT_DATA_I16) return (T)(short)raw;
  // IEEE binary16 -> fp32 by hand (no cuda_fp16.h under NVRTC): exact, subnormals and inf / NaN included
  const unsigned sign = (unsigned)(raw & 0x8000u) << 16, ex = (raw >> 10) & 0x1fu, man = raw & 0x3ffu;
  float f;
  if (ex == 0) f = __uint_as_float(sign) + (sign ? -1.f : 1.f) * (float)man * 5.9604644775390625e-08f;  // man * 2^-24
  else if (ex == 31) f = __uint_as_float(sign | 0x7f800000u | (man << 13));
  else f = __uint_as_float(sign | ((ex + 112u) << 23) | (man << 13));
  return (T)f;
}
// one data row of m 16-bit values -> m values of T in shared memory (128-bit loads when `vec`: row 16-byte aligned, m % 8 == 0)
template <typename T> __device__ __forceinline__ void load_row16(const void* rows, int64_t row_index, int64_t stride, int m, int kind, bool vec,
                                                                 T* dst, int lane) {
  const unsigned short* src = static_cast<const unsigned short*>(rows) + row_index * stride;
  if (vec) {
    for (int i = lane; i < m / 8; i += 32) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(src) + i);
      const unsigned v[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dst[8 * i + 2 * e] = data16_to<T>((unsigned short)(v[e] & 0xffffu), kind);
        dst[8 * i + 2 * e + 1] = data16_to<T>((unsigned short)(v[e] >> 16), kind);
      }
    }
  } else {
    for (int i = lane; i < m; i += 32) dst[i] = data16_to<T>(__ldg(src + i), kind);
  }
}

// ---- dynamic problem queue of the persistent kernels -------------------------------------------------------------------
// A worker (warp or CTA) takes problem `worker index` first; every further problem is a ticket from a global counter
// (zero at launch).  Iteration counts vary a lot between problems (LM gauss2: mean 10, a few at max_steps = 256), so a static
// grid-stride assignment leaves most SMs idle while the unluckiest warps finish (simulated on the measured count
// distribution: 15 % tail with static assignment, < 2 % with the queue).  Results do not depend on the assignment.
// queue == NULL: static grid-stride sequence.
__device__ __forceinline__ int64_t next_problem_warp(unsigned long long* queue, int64_t prev, int64_t nworkers, int lane) {
  if (queue == nullptr) return prev + nworkers;
  unsigned long long t = 0;
  if (lane == 0) t = atomicAdd(queue, 1ull);
  t = __shfl_sync(kFull, t, 0);
  return nworkers + (int64_t)t;
}

// ---- norms over a small per-lane-replicated vector and over warp-distributed rows -----------------------------
// acc_* implement max_norm / two_norm / rms_norm (lineax.internal, re-exported at optimistix/_misc.py:41-45)
template <typename T> struct NormAcc {
  T v;
  int kind;
  __device__ __forceinline__ explicit NormAcc(int k) : v(T(0)), kind(k) {}
  __device__ __forceinline__ void add(T x) {  // x >= 0 or NaN
    if (kind == B200OPT_NORM_MAX) v = nan_max(v, x);
    else v += x * x;
  }
  // combine across the warp (rows distributed over lanes)
  __device__ __forceinline__ void warp_combine() {
    if (kind == B200OPT_NORM_MAX) v = warp_nanmax_nonneg(v);
    else v = warp_sum(v);
  }
  __device__ __forceinline__ T finish(int count) const {
    if (kind == B200OPT_NORM_MAX) return v;
    if (kind == B200OPT_NORM_TWO) return t_sqrt(v);
    return t_sqrt(v / T(count));
  }
};

// ---- dense SPD solve in registers: (G + lam I) p = g, G packed lower-triangular, fully unrolled ---------------
template <int N> struct Tri {
  static constexpr int size = N * (N + 1) / 2;
  static __host__ __device__ constexpr int at(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
};

template <typename T, int N>
__device__ __forceinline__ void spd_solve(const T (&G)[Tri<N>::size], T lam, const T (&g)[N], T (&p)[N]) {
  T L[Tri<N>::size];  // strictly-lower entries of the Cholesky factor; diagonal slot holds 1/L_jj
#pragma unroll
  for (int j = 0; j < N; ++j) {
    T s = G[Tri<N>::at(j, j)] + lam;
#pragma unroll
    for (int k = 0; k < j; ++k) s -= L[Tri<N>::at(j, k)] * L[Tri<N>::at(j, k)];
    const T inv = t_rsqrt(s);  // s < 0 or NaN -> NaN, s == 0 -> inf: reported as `singular` by the caller
    L[Tri<N>::at(j, j)] = inv;
#pragma unroll
    for (int i = j + 1; i < N; ++i) {
      T t = G[Tri<N>::at(i, j)];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= L[Tri<N>::at(i, k)] * L[Tri<N>::at(j, k)];
      L[Tri<N>::at(i, j)] = t * inv;
    }
  }
  T z[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T s = g[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[Tri<N>::at(i, k)] * z[k];
    z[i] = s * L[Tri<N>::at(i, i)];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    T s = z[i];
#pragma unroll
    for (int k = i + 1; k < N; ++k) s -= L[Tri<N>::at(k, i)] * p[k];
    p[i] = s * L[Tri<N>::at(i, i)];
  }
}

#ifndef __CUDACC_RTC__
// thread-local last-error string (host side)
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
}  // namespace b200opt
#include <mutex>
namespace b200opt {

// Launch configuration of ONE kernel (one `static LaunchCache` per launcher instantiation), kept PER DEVICE ORDINAL and
// shared by all host threads: cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device, process-wide property of a
// kernel, so a host thread that drives cuda:0 and then cuda:1 (the `_host` entry points take a device argument) has to
// raise it on each device, and the value may only ever grow (another thread's launch may rely on the larger one).
// Ticket counter of the dynamic problem queue (next_problem_warp): allocated, zeroed and freed in stream order on the caller's
// stream - nothing persists between calls and nothing is shared between launches.  If the stream-ordered allocator is not
// available the kernels run their static schedule (ptr stays NULL).
// The chunked host pipeline (capi.cu) launches on four streams; a block freed in stream order on one stream cannot be handed to
// another without a dependency, so per-chunk cudaMallocAsync serialised the pipeline (measured: e2e 28 ms instead of 21 ms per
// 1 Mi fits).  The pipeline therefore lends each launch 8 bytes of its own per-stream staging slab through this thread-local.
extern thread_local unsigned long long* g_queue_scratch;
struct QueueCounter {
  unsigned long long* ptr = nullptr;
  cudaStream_t stream = nullptr;
  bool owned = false;
  void acquire(cudaStream_t s, void* caller_scratch = nullptr) {
    stream = s;
    unsigned long long* lent = g_queue_scratch != nullptr ? g_queue_scratch : static_cast<unsigned long long*>(caller_scratch);
    if (lent != nullptr) {  // caller-owned scratch (b200opt_options_t.scratch, or the host pipeline's slab): zero it in stream order
      if (cudaMemsetAsync(lent, 0, sizeof(unsigned long long), s) == cudaSuccess) ptr = lent;
      else cudaGetLastError();
      return;
    }
    owned = true;
    void* q = nullptr;
    if (cudaMallocAsync(&q, sizeof(unsigned long long), s) != cudaSuccess) {
      cudaGetLastError();
      return;
    }
    if (cudaMemsetAsync(q, 0, sizeof(unsigned long long), s) != cudaSuccess) {
      cudaGetLastError();
      cudaFreeAsync(q, s);
      return;
    }
    ptr = static_cast<unsigned long long*>(q);
  }
  ~QueueCounter() {
    if (ptr && owned) cudaFreeAsync(ptr, stream);
  }
};

struct LaunchCache {
  static constexpr int kMaxDevices = 64;
  std::mutex mu;
  size_t smem_set[kMaxDevices];  // largest dynamic shared-memory size the attribute was raised to on this device
  size_t occ_smem[kMaxDevices];  // shared-memory size the cached occupancy belongs to
  int blocks[kMaxDevices];       // resident blocks per SM at occ_smem
  LaunchCache() {
    for (int d = 0; d < kMaxDevices; ++d) smem_set[d] = 0, occ_smem[d] = ~size_t(0), blocks[d] = 1;
  }
  // Makes `smem` bytes of dynamic shared memory launchable on the current device and returns the number of resident
  // blocks per SM for (threads, smem) (>= 1), or -1 after set_error().
  template <class Kernel> int configure(Kernel kern, int threads, size_t smem) {
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return -1;
    if (dev < 0 || dev >= kMaxDevices) {
      set_error("device ordinal %d outside the launch cache (%d devices)", dev, kMaxDevices);
      return -1;
    }
    std::lock_guard<std::mutex> lock(mu);
    if (smem > smem_set[dev] && smem > 48 * 1024) {
      if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                     "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)"))
        return -1;
      smem_set[dev] = smem;
    }
    if (occ_smem[dev] != smem) {
      int occ = 0;
      if (check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem), "cudaOccupancyMaxActiveBlocksPerMultiprocessor"))
        return -1;
      blocks[dev] = occ > 0 ? occ : 1;
      occ_smem[dev] = smem;
    }
    return blocks[dev];
  }
};
#endif

}  // namespace b200opt
