// Large row-sharded Gauss-Newton Gram on the 5th-generation tensor cores (sm_100a: tcgen05 + TMEM + TMA).
//
// Replaces, for ONE large least-squares problem (BASELINE configs[4]: m = 64Mi rows, n = 256 parameters), the
// O(m n^2) part of NewtonDescent's `lx.linear_solve(J, r)` (optimistix/_solver/gauss_newton.py:64-78, which
// resolves to a QR of the m x n operator) by the normal equations the north_star asks for:
//     G = J^T J  (n x n),   g = J^T r  (n),   rr = r^T r
// per row shard; the caller all-reduces [G | g | rr] across ranks (NCCL) and solves (G + lambda I) p = g.
//
// Kernel design (persistent, one CTA per SM, 320 threads):
//   * J [m][256] fp32 row-major is streamed ONCE from HBM by TMA in tiles of 64 rows x 256 columns (64 KiB), as 8
//     boxes of 32 columns (128 bytes = the swizzle span; CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), 3-stage mbarrier ring
//     (192 KiB of shared memory).  Tile rows / stages are compile-time (GRAM_KROWS / GRAM_STAGES; 32 x 6 and 32 x 7
//     measured slower: the per-tile synchronisation of the CUDA-core warps costs more than the shallower ring).
//   * the same shared-memory tile is BOTH MMA operands: A = J_tile^T (M = parameter index) and B = J_tile
//     (N = parameter index), both "MN-major" (the parameter index is the contiguous one), K = rows, kind::tf32
//     (K = 8 rows per instruction), fp32 accumulation in TMEM.  Symmetry: only the blocks
//     (rows 0-127 x cols 0-255) and (rows 128-255 x cols 128-255) are computed - 384 of 512 TMEM columns, 0.75 of
//     the MMA work.  The accumulators stay in TMEM across ALL tiles of the CTA; there is one epilogue per CTA.
//   * g = J^T r, rr = r^T r and - in the fused mode of b200opt_gn_gram_linear - the residual r = J y - d itself come
//     from the CUDA cores: 8 warps make ONE pass over the tile while the tensor core works on it.  A lane keeps 32
//     columns of one row in registers (conflict-free 128-bit reads under the swizzle), forms its share of the row's
//     dot product with y, gets the row's residual by a 3-step butterfly over the 8 lanes of the row and accumulates
//     J[row][c] * r[row] into its 32 column partials.  Shared memory is the contended resource (the tensor core streams
//     its operands from it as well: 20 KB of operand reads per 8 KB of tile), which is why the tile is read only once.
//   * per-CTA partials go to a workspace; a second small kernel adds them in a fixed order (deterministic) and
//     mirrors the triangle.
// Roofline: HBM-bound by construction (68.7 GB of J per pass at configs[4]); tensor pipe ~55-68% busy at the HBM rate.
// Measured (profiles/README.md): fused pass 85% of the measured HBM copy rate, unfused Gram 96%.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "gn_state.cuh"

namespace b200opt {
int device_sm_count();
// workspace layout (b200opt_gn_gram_workspace_size): [256 per-CTA partials][fp64 n x n scratch of the solve][counters, 4 KiB]
size_t gn_workspace_solve_offset() { return (size_t)256 * (256 * 256 + 256 + 4) * 4; }
size_t gn_workspace_sync_offset() { return gn_workspace_solve_offset() + (size_t)256 * 256 * 8; }

namespace gram {

constexpr int N = 256;            // parameters (columns of J)
#ifndef GRAM_KROWS
#define GRAM_KROWS 64
#endif
#ifndef GRAM_STAGES
#define GRAM_STAGES 3
#endif
constexpr int KROWS = GRAM_KROWS;   // rows of J per tile
constexpr int STAGES = GRAM_STAGES; // STAGES * KROWS * 1 KiB of shared memory in flight per SM
constexpr int BOX_COLS = 32;      // 32 fp32 = 128 bytes = one SWIZZLE_128B span
constexpr int GROUPS = N / BOX_COLS;                    // 8 column groups per tile
constexpr int GROUP_BYTES = KROWS * 128;                // 8 KiB: one TMA box in shared memory
constexpr int TILE_BYTES = GROUPS * GROUP_BYTES;        // 64 KiB
constexpr int THREADS = 320;                            // warp 0: TMA, warp 1: MMA/TMEM, warps 2-9: g, rr, epilogue
constexpr int TMEM_COLS = 512;                          // 384 used
constexpr int GRED_FLOATS = 8 * N + 8;                   // end-of-kernel reduction of g and rr over the 8 g-warps
constexpr size_t SMEM_PAYLOAD = (size_t)STAGES * TILE_BYTES + 2 * STAGES * KROWS * 4 + GRED_FLOATS * 4 + 256;
// the tiles need 1024-byte alignment (swizzle atom); reserve slack for rounding the base up unless that would exceed the
// 227 KiB limit, in which case the kernel traps if the dynamic shared-memory window is not already aligned
constexpr size_t SMEM_ALIGN_SLACK = (SMEM_PAYLOAD + 1024 <= 232448) ? 1024 : 0;
constexpr size_t SMEM_BYTES = SMEM_ALIGN_SLACK + SMEM_PAYLOAD;
static_assert(SMEM_BYTES <= 232448, "tile ring does not fit the 227 KiB of shared memory per CTA");
static_assert(KROWS % 32 == 0, "a g-warp pass covers 32 rows of the tile");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// UMMA shared-memory descriptor, MN-major TF32 operand.  MN-major 32-bit operands have exactly one legal layout,
// SWIZZLE_128B_BASE32B ("128-byte swizzle with 32-byte atoms": within each 128-byte row the index of the 32-byte
// atom is XORed with (row & 3); measured with tools/umma_probe.cu - the plain SWIZZLE_128B layout yields zeros).
// cute::UMMA::SmemDescriptor bit layout:
//   [0,14) start >> 4 | [16,30) leading byte offset >> 4 (stride between 32-column groups)
//   [32,46) stride byte offset >> 4 (stride between 4-row swizzle atoms = 512 B) | [46,48) version = 1
//   [61,64) layout type = 1 (SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((GROUP_BYTES >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, TF32 x TF32, both MN-major, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(int n_dim) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n_dim >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Exchange buffer of the row-sharded problem (one per rank, mapped into every peer by CUDA IPC):
//   float    slots[2 parities][world][XCHG_STRIDE]   slot [p][q] = rank q's [G | g | rr | f-stat] of an iteration with parity p
//   uint32   flags[2][world]                           flags[p][q] = sequence number of the last block rank q published there
//   uint32   done                                      local "last block" counter of the publishing kernel
//   uint32   seq_counter                               local: sequence number of the last COMPLETED exchange (state-driven path)
constexpr int XCHG_STRIDE = N * N + N + 4;
__host__ __device__ inline size_t xchg_flags_offset(int world) { return (size_t)2 * world * XCHG_STRIDE * sizeof(float); }
__host__ __device__ inline size_t xchg_bytes(int world) { return xchg_flags_offset(world) + (size_t)(2 * world + 4) * sizeof(unsigned); }

struct PeerXchg {
  unsigned char* const* bases;  // device array [world]: every rank's exchange buffer as mapped in THIS process (NULL: no exchange)
  int world, rank, parity;
  unsigned seq;
  const unsigned* seq_counter;  // state-driven path: seq = *seq_counter + 1, parity = seq & 1 (the solve kernel advances it)
};

struct Params {
  const float* r;      // [m] residual (plain mode), or the data vector d (fused mode: r = J y - d is formed in-kernel)
  const float* y;      // [N] fused mode: current iterate; NULL = plain mode
  const float* r_old;  // [m] fused mode, optional: previous residual for the f-space termination statistic
  float* r_out;        // [m] fused mode: residual written back
  unsigned* fmax_bits; // fused mode, optional: max_i |r_i - r_old_i| / (atol + rtol |r_i|) as IEEE bits (atomicMax)
  float rtol, atol;
  int64_t m;
  float* ws;           // per-CTA partials: [grid][N*N + N + 1] (+ padding to a multiple of 4 floats)
  int num_tiles;
  int n_cols;            // columns of J that exist (4 <= n_cols <= 256, a multiple of 4); the kernel works on the matrix padded
                         // with zero columns to 256: column groups beyond n_cols are zeroed once and never loaded
  // ---- state-driven iteration (b200opt_gn_iterate): y_eval, the residual ping-pong parity and the frozen flag come from
  // the device state; the cross-CTA reduction, the mirror of the triangle and the publication to the peers are folded into
  // this kernel (cooperative launch: the CTAs meet at a grid-wide counter before each reduces its slice of the output).
  const GnState* state;  // NULL: y / r_old / r_out above are used as given
  float* rbuf0;          // residual buffers; step k writes rbuf[k & 1] and compares against rbuf[(k & 1) ^ 1]
  float* rbuf1;
  float* out;            // fold: [G | g | rr | f-stat]; NULL: per-CTA partials only (gram_reduce_kernel follows)
  unsigned* sync;        // fold: two grid-wide counters (arrived, reduced), zero between launches
  PeerXchg x;            // fold: publication of `out` into every peer's exchange buffer
  int debug;            // ablation switches for profiling only (build-time constant in the launcher): 1 no MMA, 2 no g loop, 4 no fused r
};
constexpr int WS_STRIDE = N * N + N + 4;

__global__ void __launch_bounds__(THREADS, 1) gram_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  if (SMEM_ALIGN_SLACK) smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  else if (smem_u32(smem_raw) & 1023u) __trap();
  unsigned char* tiles = smem;                                             // STAGES x 64 KiB, 1024-byte aligned
  float* rs = reinterpret_cast<float*>(smem + (size_t)STAGES * TILE_BYTES);  // STAGES x KROWS: residual tile (or d tile)
  float* ros = rs + STAGES * KROWS;                                         // STAGES x KROWS: previous residual tile
  float* gred = ros + STAGES * KROWS;                                       // [8][N] + 8: g-warp partials of g and rr
  uint64_t* full = reinterpret_cast<uint64_t*>(gred + GRED_FLOATS);        // TMA -> consumers
  uint64_t* empty = full + STAGES;                                          // consumers -> TMA
  uint64_t* accum_done = empty + STAGES;                                    // last MMA retired -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_done + 1);

  const float* y_ptr = p.y;
  const float* r_old_ptr = p.r_old;
  float* r_out_ptr = p.r_out;
  if (p.state) {
    if (gn_frozen(p.state)) return;  // uniform over the grid: the solve has ended, this iteration is a no-op
    const int par = p.state->steps & 1;
    y_ptr = p.state->y_eval;
    r_out_ptr = par ? p.rbuf1 : p.rbuf0;
    r_old_ptr = par ? p.rbuf0 : p.rbuf1;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = (p.num_tiles > (int)blockIdx.x) ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);        // one arrive.expect_tx by the producer (+ the TMA transaction bytes)
      mbar_init(&empty[s], 1 + 8);   // tcgen05.commit of the MMA warp + one arrive per g-warp
    }
    mbar_init(accum_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation is a warp-wide operation; this warp also frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // n < 256: the 32-column groups that lie entirely beyond the last column are never loaded; they are zeroed here, once, in
  // every stage (a partially filled group is zero-filled by the TMA's out-of-bounds handling)
  const int live_groups = (p.n_cols + BOX_COLS - 1) / BOX_COLS;
  if (live_groups < GROUPS) {
    const int dead_words = (GROUPS - live_groups) * GROUP_BYTES / 16;
    for (int s = 0; s < STAGES; ++s) {
      uint4* dead = reinterpret_cast<uint4*>(tiles + (size_t)s * TILE_BYTES + (size_t)live_groups * GROUP_BYTES);
      for (int i = threadIdx.x; i < dead_words; i += THREADS) dead[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros -> visible to the tensor core's async-proxy reads
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer (one elected lane) ================================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_tiles; ++it, s = (s + 1 == STAGES) ? 0 : s + 1, ph ^= (s == 0)) {
        mbar_wait(&empty[s], ph ^ 1);
        const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
        const int64_t row0 = tile * KROWS;
        const int64_t rows_left = p.m - row0;
        const uint32_t r_bytes = (uint32_t)((rows_left < KROWS ? rows_left : KROWS) * 4);
        const bool with_old = y_ptr != nullptr && r_old_ptr != nullptr;
        mbar_expect_tx(&full[s], (uint32_t)live_groups * GROUP_BYTES + r_bytes * (with_old ? 2u : 1u));  // OOB rows / columns are zero-filled and still counted
        unsigned char* dst = tiles + (size_t)s * TILE_BYTES;
#pragma unroll
        for (int g = 0; g < GROUPS; ++g)
          if (g < live_groups) tma_load_2d(dst + g * GROUP_BYTES, &tmap, g * BOX_COLS, (int)row0, &full[s]);
        bulk_load_1d(rs + s * KROWS, p.r + row0, r_bytes, &full[s]);
        if (with_old) bulk_load_1d(ros + s * KROWS, r_old_ptr + row0, r_bytes, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (one elected lane) ================================
    if (lane == 0) {
      constexpr uint32_t idesc_full = make_idesc(256);
      constexpr uint32_t idesc_half = make_idesc(128);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_tiles; ++it, s = (s + 1 == STAGES) ? 0 : s + 1, ph ^= (s == 0)) {
        mbar_wait(&full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = smem_u32(tiles + (size_t)s * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < ((p.debug & 1) ? 0 : KROWS / 8); ++k) {
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          // rows 0-127 of G x all 256 columns -> TMEM columns [0, 256)
          umma_tf32(tmem_base + 0, make_desc(base + k * 1024), make_desc(base + k * 1024), idesc_full, acc);
          // rows 128-255 of G x columns 128-255 -> TMEM columns [256, 384)
          umma_tf32(tmem_base + 256, make_desc(base + 4 * GROUP_BYTES + k * 1024), make_desc(base + 4 * GROUP_BYTES + k * 1024),
                    idesc_half, acc);
        }
        umma_commit(&empty[s]);                     // frees the stage once these MMAs have read it
        if (it == my_tiles - 1) umma_commit(accum_done);
      }
    }
  } else {
    // ================================ g = J^T r, rr = r^T r on CUDA cores ================================
    // ONE pass over the tile per g-warp: a lane keeps 32 columns of one row in registers, forms its share of the row's
    // dot product with y (fused mode), gets the row's residual with a 3-step butterfly over the 8 lanes of the row, and
    // immediately accumulates J[row][c] * r[row] into its 32 column partials of g.  Shared memory is read exactly once
    // (512 wavefronts per 64 KiB tile) - it is the contended resource: the tensor core streams its operands from it too.
    const int b = threadIdx.x - 64;                 // 0..255
    const bool fused = y_ptr != nullptr;
    constexpr int LPR = 8;                          // lanes per row
    constexpr int RPP = 32 / LPR;                   // rows per warp per pass (4)
    constexpr int PASSES = KROWS / (8 * RPP);       // passes per tile (2 for 64-row tiles)
    constexpr int NGRP = GROUPS / 2;                // column groups per lane (4): {gsel, gsel + 2, gsel + 4, gsel + 6}
    const int gw = warp - 2, sub = lane % LPR, rlane = lane / LPR, my_atom = sub & 3, gsel = sub >> 2;
    // the two 16-byte halves of a 32-byte atom are read in opposite order by the two half-rows, so that the 8 lanes of
    // a quarter-warp (= one row) hit 8 distinct 16-byte bank groups; y is loaded in the matching order
    const int flip = gsel;
    float yreg[NGRP * 8];
#pragma unroll
    for (int j = 0; j < NGRP; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) yreg[j * 8 + e] = fused ? y_ptr[(gsel + 2 * j) * 32 + my_atom * 8 + (e ^ (flip << 2))] : 0.f;
    float acc_g[NGRP * 8];
#pragma unroll
    for (int c = 0; c < NGRP * 8; ++c) acc_g[c] = 0.f;
    float acc_rr = 0.f, worst = 0.f;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < my_tiles; ++it) {
      mbar_wait(&full[s], ph);
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      const int64_t rows_left = p.m - tile * KROWS;
      const int rows = rows_left < KROWS ? (int)rows_left : KROWS;
      const unsigned char* tbase = tiles + (size_t)s * TILE_BYTES;
      const float* rt = rs + s * KROWS;
      if (!(p.debug & 2)) {
#pragma unroll
        for (int pass = 0; pass < PASSES; ++pass) {
          const int row = pass * 8 * RPP + gw * RPP + rlane;
          const unsigned char* ap = tbase + gsel * GROUP_BYTES + row * 128 + ((my_atom ^ (row & 3)) << 5);
          float4 v[NGRP * 2];
#pragma unroll
          for (int j = 0; j < NGRP; ++j) {
            v[2 * j] = *reinterpret_cast<const float4*>(ap + j * 2 * GROUP_BYTES + (flip << 4));
            v[2 * j + 1] = *reinterpret_cast<const float4*>(ap + j * 2 * GROUP_BYTES + ((flip ^ 1) << 4));
          }
          float rv = rt[row];                        // plain mode: the residual; fused mode: the data value d
          if (fused) {
            float dot[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < NGRP * 2; ++q) {
              dot[0] = fmaf(v[q].x, yreg[4 * q + 0], dot[0]);
              dot[1] = fmaf(v[q].y, yreg[4 * q + 1], dot[1]);
              dot[2] = fmaf(v[q].z, yreg[4 * q + 2], dot[2]);
              dot[3] = fmaf(v[q].w, yreg[4 * q + 3], dot[3]);
            }
            // the partial of a lane is summed in an order that does not depend on `flip`; the butterfly leaves the
            // same bits in all 8 lanes of the row
            float lo = (dot[0] + dot[1]) + (dot[2] + dot[3]);
            lo += __shfl_xor_sync(kFull, lo, 1);
            lo += __shfl_xor_sync(kFull, lo, 2);
            lo += __shfl_xor_sync(kFull, lo, 4);
            rv = lo - rv;
          }
          if (row >= rows) rv = 0.f;                 // ragged last tile: J rows are zero-filled, r / d rows are stale
          if (sub == 0 && row < rows) {
            if (fused) {
              r_out_ptr[tile * KROWS + row] = rv;
              if (r_old_ptr) worst = nan_max(worst, fabsf(rv - ros[s * KROWS + row]) / (p.atol + p.rtol * fabsf(rv)));
            }
            acc_rr = fmaf(rv, rv, acc_rr);
          }
#pragma unroll
          for (int q = 0; q < NGRP * 2; ++q) {
            acc_g[4 * q + 0] = fmaf(v[q].x, rv, acc_g[4 * q + 0]);
            acc_g[4 * q + 1] = fmaf(v[q].y, rv, acc_g[4 * q + 1]);
            acc_g[4 * q + 2] = fmaf(v[q].z, rv, acc_g[4 * q + 2]);
            acc_g[4 * q + 3] = fmaf(v[q].w, rv, acc_g[4 * q + 3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    if (fused && p.fmax_bits && sub == 0)
      atomicMax(p.fmax_bits, __float_as_uint(worst != worst ? __int_as_float(0x7fc00000) : worst));
    // ---- g: rows of a warp (xor 8, 16), then the 8 warps through shared memory in a fixed order (deterministic)
#pragma unroll
    for (int c = 0; c < NGRP * 8; ++c) {
      acc_g[c] += __shfl_xor_sync(kFull, acc_g[c], 8);
      acc_g[c] += __shfl_xor_sync(kFull, acc_g[c], 16);
    }
    if (rlane == 0) {
#pragma unroll
      for (int j = 0; j < NGRP; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) gred[gw * N + (gsel + 2 * j) * 32 + my_atom * 8 + (e ^ (flip << 2))] = acc_g[j * 8 + e];
    }
    acc_rr = warp_sum(acc_rr);
    if (lane == 0) gred[8 * N + gw] = acc_rr;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // ================================ epilogue: TMEM -> per-CTA partial in the workspace ================================
    float* ws = p.ws + (size_t)blockIdx.x * WS_STRIDE;
    {
      float tg = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tg += gred[w * N + b];
      ws[N * N + b] = tg;
      if (b == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += gred[8 * N + w];
        ws[N * N + N] = t;
      }
    }
    if (my_tiles > 0) {
      mbar_wait(accum_done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int q = warp & 3;                       // TMEM lane quarter this warp may access
      const int half = (warp - 2) >> 2;             // which half of the 384 accumulator columns
      const int row_lo = q * 32 + lane;             // accumulator row (TMEM lane)
#pragma unroll 1
      for (int c0 = half * 192; c0 < half * 192 + 192; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
              "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
              "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // columns [0,256): G[row_lo][c];  columns [256,384): G[128 + row_lo][128 + (c - 256)]
        const int grow = c0 < 256 ? row_lo : 128 + row_lo;
        const int gcol = c0 < 256 ? c0 : 128 + (c0 - 256);
        float4* dst = reinterpret_cast<float4*>(ws + (size_t)grow * N + gcol);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                               __uint_as_float(v[4 * j + 3]));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (p.out) {
      // ================================ folded reduction + publication ================================
      // Every CTA's partial is in the workspace: meet at the grid-wide counter (all CTAs are co-resident: cooperative
      // launch, one CTA per SM), then CTA c adds slice c of the output over the partials in CTA order - the same fixed
      // order as gram_reduce_kernel, so the two paths are bit-identical - mirrors the triangle that was not computed and,
      // in a row-sharded run, stores each value straight into this rank's slot of every peer's exchange buffer (P2P
      // stores over NVLink / NVSwitch).  The last CTA to finish raises this rank's flag on every peer.
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (b == 0) {
        // one poller per CTA, with back-off: 148 threads hammering one L2 line slow down the CTAs that are still streaming J
        // (measured: +170 us per pass with a 64 ns poll interval)
        atomicAdd(p.sync, 1u);
        unsigned ns = 256;
        for (;;) {
          unsigned seen;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.sync) : "memory");
          if (seen >= gridDim.x) break;
          __nanosleep(ns);
          if (ns < 4096) ns *= 2;
        }
        __threadfence();
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      constexpr int TOTAL = N * N + N + 2;  // G | g | rr | f-stat
      const int per = (TOTAL + (int)gridDim.x - 1) / (int)gridDim.x;
      const int lo = (int)blockIdx.x * per;
      const int hi = lo + per < TOTAL ? lo + per : TOTAL;
      unsigned seq = p.x.seq;
      int parity = p.x.parity;
      if (p.x.bases && p.x.seq_counter) {
        seq = *p.x.seq_counter + 1u;
        parity = (int)(seq & 1u);
      }
      for (int idx = lo + b; idx < hi; idx += 256) {
        float sv;
        if (idx < N * N + N + 1) {
          int src = idx;
          if (idx < N * N) {
            const int a = idx / N, c = idx % N;
            if (a >= 128 && c < 128) src = c * N + a;  // lower-left block: the transposed entry of the upper-right block
          }
          // the partials are added in CTA order (the order of gram_reduce_kernel); 16 loads are in flight per thread - with
          // 8 warps per SM the reduction is bound by memory-level parallelism, not by the adds
          sv = 0.f;
          const float* w = p.ws + src;
          const int parts = (int)gridDim.x;
          int c = 0;
          for (; c + 16 <= parts; c += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = __ldcg(w + (size_t)(c + u) * WS_STRIDE);
#pragma unroll
            for (int u = 0; u < 16; ++u) sv += v[u];
          }
          for (; c < parts; ++c) sv += __ldcg(w + (size_t)c * WS_STRIDE);
          if (idx < N * N && idx / N == idx % N && idx / N >= p.n_cols) sv = 1.f;  // padded unknowns: unit diagonal, they stay zero
          p.out[idx] = sv;
        } else {
          sv = __ldcg(p.out + idx);  // the f-statistic (atomicMax target of every CTA; complete after the grid-wide meeting)
        }
        if (p.x.bases) {
          const size_t off = ((size_t)(parity * p.x.world + p.x.rank) * XCHG_STRIDE + idx) * sizeof(float);
          for (int q = 0; q < p.x.world; ++q) *reinterpret_cast<float*>(p.x.bases[q] + off) = sv;
        }
      }
      if (p.x.bases) __threadfence_system();  // this thread's peer stores are visible system-wide before the CTA is counted
      else __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (b == 0) {
        const unsigned prev = atomicAdd(p.sync + 1, 1u);
        if (prev == gridDim.x - 1) {  // last CTA: every other CTA has fenced and counted
          p.sync[0] = 0;
          p.sync[1] = 0;
          if (p.x.bases) {
            __threadfence_system();
            for (int q = 0; q < p.x.world; ++q) {
              volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(p.x.bases[q] + xchg_flags_offset(p.x.world)) + parity * p.x.world + p.x.rank;
              *flag = seq;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// Deterministic reduction of the per-CTA partials, mirroring the triangle that was not computed.
// out = [G (N*N, full symmetric) | g (N) | rr]   (slot N*N+N+1, the f-statistic, was written by gram_kernel's atomicMax)
// With an exchange (row-sharded problem on several GPUs) the same kernel also PUBLISHES the block: every value is stored
// straight into the slot of this rank in every peer's exchange buffer (P2P stores over NVLink / NVSwitch), and the last CTA to
// finish raises this rank's flag on every peer.  The collective is therefore part of the reduction kernel; the consumer
// (gn_solve_kernel) waits on the flags.  No NCCL call, no extra launch, no host involvement.
__global__ void gram_reduce_kernel(const float* ws, int parts, int live_parts, float* out, PeerXchg x, const GnState* state, int n_cols) {
  if (state && gn_frozen(state)) return;  // state-driven iteration: the solve has ended, nothing was computed, nothing is published
  if (x.bases && x.seq_counter) {
    x.seq = *x.seq_counter + 1u;
    x.parity = (int)(x.seq & 1u);
  }
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = N * N + N + 1;
  float s = 0.f;
  if (idx < total) {
    int src = idx;
    if (idx < N * N) {
      const int a = idx / N, b = idx % N;
      if (a >= 128 && b < 128) src = b * N + a;  // lower-left block: read the transposed entry of the upper-right block
    }
    const bool is_gram = idx < N * N;
    const int count = is_gram ? live_parts : parts;  // CTAs without tiles never wrote their Gram block
    for (int c = 0; c < count; ++c) s += ws[(size_t)c * WS_STRIDE + src];
    if (is_gram && idx / N == idx % N && idx / N >= n_cols) s = 1.f;  // padded unknowns (n < 256): unit diagonal, they stay zero
    out[idx] = s;
  } else if (idx == total) {
    s = out[idx];  // the f-statistic slot rides along
  }
  if (x.bases == nullptr) return;
  if (idx <= total) {
    const size_t off = ((size_t)(x.parity * x.world + x.rank) * XCHG_STRIDE + idx) * sizeof(float);
    for (int q = 0; q < x.world; ++q) *reinterpret_cast<float*>(x.bases[q] + off) = s;
  }
  __threadfence_system();  // this thread's peer stores are visible system-wide before the block is counted
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* done = reinterpret_cast<unsigned*>(x.bases[x.rank] + xchg_flags_offset(x.world)) + 2 * x.world;
    const unsigned prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {  // last CTA: every other CTA has fenced and counted
      *done = 0;
      __threadfence_system();
      for (int q = 0; q < x.world; ++q) {
        volatile unsigned* flag = reinterpret_cast<volatile unsigned*>(x.bases[q] + xchg_flags_offset(x.world)) + x.parity * x.world + x.rank;
        *flag = x.seq;
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

}  // namespace gram
}  // namespace b200opt

using namespace b200opt;

extern "C" {

int64_t b200opt_gn_gram_workspace_size(int32_t n) {
  if (n < 4 || n > gram::N) return -1;  // the workspace does not depend on n: everything is laid out 256 wide
  // per-CTA partials for up to 256 CTAs + an fp64 n x n scratch for the solve kernel + 4 KiB (grid-wide counters)
  return (int64_t)gn_workspace_sync_offset() + 4096;
}

// `fold`: state-driven iteration (b200opt_gn_iterate) - reduction and publication inside gram_kernel (cooperative launch)
struct GnFold {
  const GnState* state;
  float *rbuf0, *rbuf1;
  bool fold;  // true: reduction + publication inside gram_kernel (cooperative launch); false: gram_reduce_kernel follows
};
static int32_t gn_gram_impl(const float* J, const float* r, const float* y, const float* r_old, float* r_out, uint32_t* fmax_bits,
                            float rtol, float atol, int64_t m_local, int32_t n, float* out, void* workspace, void* stream,
                            gram::PeerXchg x = gram::PeerXchg{nullptr, 0, 0, 0, 0, nullptr}, const GnFold* fold = nullptr);

int32_t b200opt_gn_gram(const float* J, const float* r, int64_t m_local, int32_t n, float* out, void* workspace, void* stream) {
  return gn_gram_impl(J, r, nullptr, nullptr, nullptr, nullptr, 0.f, 0.f, m_local, n, out, workspace, stream);
}

int32_t b200opt_gn_gram_linear(const float* J, const float* y, const float* d, const float* r_old, float* r_out,
                               uint32_t* fmax_bits, float rtol, float atol, int64_t m_local, int32_t n, float* out,
                               void* workspace, void* stream) {
  if (!y || !d || !r_out) {
    set_error("gn_gram_linear: NULL buffer");
    return B200OPT_EINVAL;
  }
  if (r_old && (reinterpret_cast<uintptr_t>(r_old) & 15)) {
    set_error("gn_gram_linear: r_old must be 16-byte aligned");
    return B200OPT_EINVAL;
  }
  return gn_gram_impl(J, d, y, r_old, r_out, fmax_bits, rtol, atol, m_local, n, out, workspace, stream);
}

int64_t b200opt_xchg_bytes(int32_t world, int32_t n) {
  if (n < 4 || n > gram::N || world < 1 || world > 64) return -1;
  return (int64_t)gram::xchg_bytes(world);
}

int32_t b200opt_xchg_alloc(int64_t bytes, void** ptr_out, void* ipc_handle_out) {
  if (!ptr_out || !ipc_handle_out || bytes <= 0) {
    set_error("xchg_alloc: bad arguments");
    return B200OPT_EINVAL;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI hands IPC handles around as 64 opaque bytes");
  void* p = nullptr;
  if (check_cuda(cudaMalloc(&p, (size_t)bytes), "xchg cudaMalloc")) return B200OPT_ECUDA;
  if (check_cuda(cudaMemset(p, 0, (size_t)bytes), "xchg memset")) return B200OPT_ECUDA;
  if (check_cuda(cudaDeviceSynchronize(), "xchg sync")) return B200OPT_ECUDA;
  cudaIpcMemHandle_t h;
  if (check_cuda(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle")) {
    cudaFree(p);
    return B200OPT_ECUDA;
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *ptr_out = p;
  return 0;
}

int32_t b200opt_xchg_open(const void* ipc_handle, void** ptr_out) {
  if (!ipc_handle || !ptr_out) {
    set_error("xchg_open: bad arguments");
    return B200OPT_EINVAL;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  return check_cuda(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

int32_t b200opt_xchg_close(void* ptr) { return ptr ? check_cuda(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle") : 0; }
int32_t b200opt_xchg_free(void* ptr) { return ptr ? check_cuda(cudaFree(ptr), "xchg cudaFree") : 0; }

int32_t b200opt_gn_gram_linear_xchg(const float* J, const float* y, const float* d, const float* r_old, float* r_out,
                                    uint32_t* fmax_bits, float rtol, float atol, int64_t m_local, int32_t n, float* out,
                                    void* workspace, const void* const* xchg_bases_dev, int32_t world, int32_t rank, int32_t parity,
                                    uint32_t seq, void* stream) {
  if (!y || !d || !r_out || !xchg_bases_dev || world < 1 || rank < 0 || rank >= world || (parity != 0 && parity != 1)) {
    set_error("gn_gram_linear_xchg: bad arguments");
    return B200OPT_EINVAL;
  }
  if (fmax_bits != reinterpret_cast<uint32_t*>(out + gram::N * gram::N + gram::N + 1)) {
    set_error("gn_gram_linear_xchg: the f-statistic must live in slot n*n+n+1 of `out` (it is published with the block)");
    return B200OPT_EINVAL;
  }
  gram::PeerXchg x{reinterpret_cast<unsigned char* const*>(const_cast<void* const*>(xchg_bases_dev)), world, rank, parity, seq, nullptr};
  return gn_gram_impl(J, d, y, r_old, r_out, fmax_bits, rtol, atol, m_local, n, out, workspace, stream, x);
}

}  // extern "C"

namespace b200opt {
// first half of b200opt_gn_iterate (gn_large.cu): fused residual + Gram pass driven by the device state, reduction and
// publication folded into the kernel
int32_t gn_iterate_gram(const float* J, const float* d, float* r0, float* r1, int64_t m_local, int32_t n, const void* state, float* out,
                        void* workspace, const void* const* xchg_bases_dev, void* xchg_local, int32_t world, int32_t rank, float rtol,
                        float atol, int32_t fold_reduce, void* stream) {
  gram::PeerXchg x{nullptr, 0, 0, 0, 0, nullptr};
  if (xchg_bases_dev) {
    x.bases = reinterpret_cast<unsigned char* const*>(const_cast<void* const*>(xchg_bases_dev));
    x.world = world;
    x.rank = rank;
    x.seq_counter = reinterpret_cast<const unsigned*>(static_cast<const unsigned char*>(xchg_local) + gram::xchg_flags_offset(world)) + 2 * world + 1;
  }
  GnFold fold{static_cast<const GnState*>(state), r0, r1, fold_reduce != 0};
  uint32_t* fmax_bits = reinterpret_cast<uint32_t*>(out + gram::N * gram::N + gram::N + 1);
  // y / r_old / r_out are taken from the state inside the kernel; `r0` stands in so that the argument checks see buffers
  return gn_gram_impl(J, d, static_cast<const GnState*>(state)->y_eval, r1, r0, fmax_bits, rtol, atol, m_local, n, out, workspace, stream, x, &fold);
}
}  // namespace b200opt

static int32_t gn_gram_impl(const float* J, const float* r, const float* y, const float* r_old, float* r_out, uint32_t* fmax_bits,
                            float rtol, float atol, int64_t m_local, int32_t n, float* out, void* workspace, void* stream, gram::PeerXchg x,
                            const GnFold* fold) {
  if (n < 4 || n > gram::N || (n & 3)) {
    set_error("gn_gram: need 4 <= n <= %d and n %% 4 == 0 (got %d); vectors and `out` keep the %d-wide layout", gram::N, n, gram::N);
    return B200OPT_EINVAL;
  }
  if (m_local < 0 || m_local > ((int64_t)1 << 31) - 64) {
    set_error("gn_gram: m_local out of range");
    return B200OPT_EINVAL;
  }
  if (!out || !workspace || (m_local > 0 && (!J || !r))) {
    set_error("gn_gram: NULL buffer");
    return B200OPT_EINVAL;
  }
  if ((reinterpret_cast<uintptr_t>(J) & 15) || (reinterpret_cast<uintptr_t>(r) & 15) || (m_local & 3)) {
    set_error("gn_gram: J and r must be 16-byte aligned and m_local a multiple of 4");
    return B200OPT_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ws = static_cast<float*>(workspace);
  const int num_tiles = (int)((m_local + gram::KROWS - 1) / gram::KROWS);
  int grid = device_sm_count();
  if (grid > 256) grid = 256;
  if (num_tiles < grid) grid = num_tiles > 0 ? num_tiles : 1;
  if (m_local == 0) {
    if (x.bases || fold) {
      set_error("gn_gram: a rank of a peer exchange must own at least one row");
      return B200OPT_EINVAL;
    }
    if (check_cuda(cudaMemsetAsync(out, 0, sizeof(float) * (gram::N * gram::N + gram::N + 1), st), "memset")) return B200OPT_ECUDA;
    return 0;
  }
  gram::EncodeTiledFn encode = gram::get_encode();
  if (!encode) {
    set_error("gn_gram: cuTensorMapEncodeTiled is not available from this driver");
    return B200OPT_ECUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)m_local};
  const cuuint64_t strides[1] = {(cuuint64_t)n * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)gram::BOX_COLS, (cuuint32_t)gram::KROWS};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(J), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    set_error("gn_gram: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    return B200OPT_ECUDA;
  }
  static LaunchCache cache;
  if (cache.configure(gram::gram_kernel, gram::THREADS, gram::SMEM_BYTES) < 0) return B200OPT_ECUDA;
  constexpr int debug = 0;  // ablation switch of profiles/README.md (1: TMA only, 2: TMA + MMA); a build-time constant in release
  gram::Params p{r, y, r_old, r_out, fmax_bits, rtol, atol, m_local, ws, num_tiles, n, nullptr, nullptr, nullptr, nullptr, nullptr,
                 gram::PeerXchg{nullptr, 0, 0, 0, 0, nullptr}, debug};
  if (fold) {
    p.state = fold->state;
    p.rbuf0 = fold->rbuf0;
    p.rbuf1 = fold->rbuf1;
  }
  if (fold && fold->fold) {
    // grid-wide meeting inside the kernel: cooperative launch guarantees that all CTAs are resident at the same time
    p.out = out;
    p.sync = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + gn_workspace_sync_offset());
    p.x = x;
    void* args[] = {&tmap, &p};
    return check_cuda(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gram::gram_kernel), dim3(grid), dim3(gram::THREADS), args,
                                                  gram::SMEM_BYTES, st),
                      "gram_kernel cooperative launch");
  }
  gram::gram_kernel<<<grid, gram::THREADS, gram::SMEM_BYTES, st>>>(tmap, p);
  if (check_cuda(cudaGetLastError(), "gram_kernel launch")) return B200OPT_ECUDA;
  const int total = gram::N * gram::N + gram::N + 1;
  gram::gram_reduce_kernel<<<(total + 1 + 255) / 256, 256, 0, st>>>(ws, grid, grid, out, x, fold ? fold->state : nullptr, n);
  return check_cuda(cudaGetLastError(), "gram_reduce_kernel launch");
}


