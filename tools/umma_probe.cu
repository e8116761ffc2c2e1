// Probe of tcgen05.mma kind::tf32 operand descriptors (MN-major, SWIZZLE_128B) - debugging aid, not product code.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <cmath>

constexpr int N = 256, KROWS = 64, GROUP_BYTES = KROWS * 128, TILE_BYTES = 8 * GROUP_BYTES;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t version) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)version << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(const float* J, float* out, int variant) {
  extern __shared__ unsigned char raw[];
  unsigned char* tile = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // place J[r][c] (r < 64, c < 256) the way TMA SWIZZLE_128B does for 32-column boxes
  for (int e = tid; e < KROWS * N; e += blockDim.x) {
    const int r = e / N, c = e % N;
    const int grp = c >> 5, unit = (c & 31) >> 2, sub = c & 3;
    float* dst = (float*)(tile + grp * GROUP_BYTES + r * 128 + (((variant == 3) ? unit : (unit ^ (r & 7))) << 4) + (sub << 2));
    if (variant >= 5) {  // SWIZZLE_128B with 32-byte atoms: the 32B-atom index within the 128B row is XORed with (row & 3)
      const int atom = (c & 31) >> 3, w = c & 7;
      dst = (float*)(tile + grp * GROUP_BYTES + r * 128 + ((atom ^ (r & 3)) << 5) + (w << 2));
    }
    *dst = J[r * N + c];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t lbo = GROUP_BYTES, sbo = 1024, layout = 2, version = 1;
    if (variant == 1) { lbo = 1024; sbo = GROUP_BYTES; }
    if (variant == 2) idesc &= ~((1u << 15) | (1u << 16));
    if (variant == 3) layout = 0;
    if (variant == 4) version = 0;
    if (variant == 5) { layout = 1; sbo = 512; }
    if (variant == 6) { layout = 1; sbo = 512; lbo = 512; sbo = GROUP_BYTES; }
    const uint32_t base = smem_u32(tile);
    for (int k = 0; k < 8; ++k) {
      uint64_t da = make_desc(base + k * 1024, lbo, sbo, layout, version);
      uint32_t acc = k > 0;
      asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                   "l"(da), "l"(da), "r"(idesc), "r"(acc)
                   : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMA
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 256; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
  std::vector<float> J(KROWS * N);
  for (int i = 0; i < KROWS * N; ++i) J[i] = (float)((i * 2654435761u >> 20) % 17) - 8.0f;  // small integers: exact in tf32
  std::vector<double> ref(128 * 256, 0.0);
  for (int a = 0; a < 128; ++a)
    for (int b = 0; b < 256; ++b) {
      double s = 0;
      for (int r = 0; r < KROWS; ++r) s += (double)J[r * N + a] * J[r * N + b];
      ref[a * 256 + b] = s;
    }
  float *dJ, *dout;
  cudaMalloc(&dJ, J.size() * 4);
  cudaMalloc(&dout, 128 * 256 * 4);
  cudaMemcpy(dJ, J.data(), J.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES + 2048);
  for (int variant = 4; variant < 7; ++variant) {
    cudaMemset(dout, 0xff, 128 * 256 * 4);
    probe<<<1, 128, TILE_BYTES + 2048>>>(dJ, dout, variant);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(128 * 256);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    int nz = 0, bad = 0;
    double maxerr = 0;
    for (int i = 0; i < 128 * 256; ++i) {
      if (out[i] != 0.f) nz++;
      double err = fabs(out[i] - ref[i]);
      if (!(err <= 1e-3)) bad++;
      if (err > maxerr) maxerr = err;
    }
    printf("variant %d: cuda=%s nonzero=%d bad=%d maxerr=%g  out[0..3]=%g %g %g %g  ref=%g %g %g %g\n", variant, cudaGetErrorString(e), nz,
           bad, maxerr, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
    if (e != cudaSuccess) break;
  }
  return 0;
}
