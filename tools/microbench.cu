// Pipe-throughput microbenchmark for the roofline denominators the batched solvers are bound by
// ("no measured fp64, fp32-FMA peak; the build must measure an FMA-loop microbenchmark").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench tools/microbench.cu
// Prints one JSON object; each figure is lanes*ops per clock per SM from in-kernel clock64, and the
// resulting TFLOP/s at the clock observed with CUDA events.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

template <int MODE> __global__ void k(float* out, long long* cyc, float seed) {
  float a[CH], b = seed, c = seed * 0.5f;
  double da[CH], db = seed, dc = seed * 0.5;
  float2 fa[CH], fb = make_float2(seed, seed), fc = make_float2(c, c);
  for (int i = 0; i < CH; ++i) {
    a[i] = seed + i + threadIdx.x;
    da[i] = a[i];
    fa[i] = make_float2(a[i], a[i] + 1.f);
  }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
      if (MODE == 1) fa[i] = __ffma2_rn(fa[i], fb, fc);
      if (MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(da[i]) : "d"(db), "d"(dc));
      if (MODE == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 4) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1 + (i & 15));
      if (MODE == 5) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(b));  // 2 distinct regs
      if (MODE == 6) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      if (MODE == 7) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CH; ++i) s += a[i] + (float)da[i] + fa[i].x + fa[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, double ops_per_instr, int blocks, int threads, float* out, long long* cyc) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<blocks, threads>>>(out, cyc, 1.0001f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<blocks, threads>>>(out, cyc, 1.0001f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(blocks);
  cudaMemcpy(h.data(), cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (auto v : h) mean += v;
  mean /= blocks;
  double lane_ops = (double)ITERS * CH * threads * ops_per_instr;  // per block == per SM (1 block/SM)
  double per_clk = lane_ops / mean;
  double total = lane_ops * blocks / (ms * 1e-3);
  printf("  \"%s\": {\"lane_ops_per_clk_per_sm\": %.1f, \"warp_instr_per_clk_per_sm\": %.3f, \"gops\": %.1f, \"tflops_if_fma\": %.2f, \"ms\": %.3f},\n",
         name, per_clk, per_clk / 32 / ops_per_instr, total / 1e9, 2 * total / 1e12, ms);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int blocks = p.multiProcessorCount, threads = 1024;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"threads_per_sm\": %d,\n", p.name, blocks, threads);
  run<0>("ffma_3reg", 1, blocks, threads, out, cyc);
  run<5>("ffma_2reg", 1, blocks, threads, out, cyc);
  run<1>("ffma2_packed", 2, blocks, threads, out, cyc);
  run<6>("fadd", 1, blocks, threads, out, cyc);
  run<2>("dfma", 1, blocks, threads, out, cyc);
  run<3>("mufu_ex2", 1, blocks, threads, out, cyc);
  run<7>("mufu_rsq", 1, blocks, threads, out, cyc);
  run<4>("shfl", 1, blocks, threads, out, cyc);
  printf("  \"note\": \"1 block of 1024 threads per SM, %d independent chains/thread, %d iterations\"\n}\n", CH, ITERS);
  return 0;
}
