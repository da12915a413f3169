// Microbenchmark: what a tcgen05.mma (kind::f16, bf16, M=128, N=128, K=16) costs as a function of WHERE its operands come
// from and of what else uses shared memory at the same time - the operand-delivery side of the loss sweeps:
//   SS  K-major A and B from shared memory            (score tile  S = R . C^T:   8 KB of smem reads per MMA)
//   TS  A from TMEM, B MN-major from shared memory    (second MMA  acc += G . C:  4 KB per MMA, transposed read)
//   TSk A from TMEM, B K-major from shared memory     (for comparison)
// alone, mixed 9 : 8 as in a sweep tile, and with a bulk-copy stream into shared memory running beside them (the TMA ring).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_operand_bench mma_operand_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../matrix-factorization-torch_b200/csrc/ptx.cuh"
using namespace xb;

constexpr int BLOCK = 128 * 128;   // one [128 rows x 64 bf16] SWIZZLE_128B block

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// mode 0: SS   1: TS (B MN-major)   2: TS (B K-major)   3: 9 SS + 8 TS(MN) per round   4: 9 SS + 8 TS(K) per round
template <int MODE>
__global__ void bench(int rounds, int copy_kb_per_round, int ld_warps, const uint8_t* src, long long* out, int random_data, int heavy) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, cbar, dummy[3];
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operand data: zeros, or random bf16 values of magnitude ~1 (the switching activity of real operands: the chip's
  // power management, not the issue logic, may set the MMA rate)
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const uint32_t lo = 0x3f00u | (h & 0x80ffu), hi = 0x3f00u | ((h >> 16) & 0x80ffu);
    reinterpret_cast<uint32_t*>(smem)[i] = random_data ? (lo | (hi << 16)) : 0u;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&cbar, 1); for (int i = 0; i < 3; ++i) mbar_init(&dummy[i], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tbase);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (random_data && warp >= 2 && warp < 6) {   // random bf16 pairs in the TMEM columns the TS MMAs read as A
    uint32_t v[16];
    for (int c = 0; c < 8; ++c) {
      for (int j = 0; j < 16; ++j) {
        uint32_t h = (threadIdx.x * 131u + c * 16u + j) * 2654435761u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        v[j] = (0x3f00u | (h & 0x80ffu)) | ((0x3f00u | ((h >> 16) & 0x80ffu)) << 16);
      }
      tmem_st16(tb + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + 128 + c * 16, v);
    }
    tmem_st_wait();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint8_t* sA = smem;               // 32 KB "row tile"
  uint8_t* sB = smem + 32768;       // 32 KB "column tile"
  uint8_t* sD = smem + 65536;       // 32 KB landing zone of the copy stream
  if (warp == 0) {
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_mn = umma_idesc_bf16(128, 128, 0, 1);
    const uint32_t a_lo = umma_desc_lo(smem_u32(sA), 16), b_lo = umma_desc_lo(smem_u32(sB), 16);
    const uint32_t bmn_lo = umma_desc_lo(smem_u32(sB), BLOCK);
    const uint32_t aug_a = umma_desc_lo(smem_u32(sD), 16), aug_b = umma_desc_lo(smem_u32(sD + 4096), 16);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < rounds; ++r) {
        if (MODE == 0 || MODE >= 3) {
          if (MODE == 7) { umma_ss_lo(tb, aug_a, aug_b, idesc_s, 1u, UMMA_DESC_HI_SW32); continue; }
          for (int kb = 0; kb < 2; ++kb)
            for (int k = 0; k < 4; ++k) umma_ss_lo(tb, a_lo + kb * (BLOCK >> 4) + 2 * k, b_lo + kb * (BLOCK >> 4) + 2 * k, idesc_s, 1u);
          if (MODE == 5 || MODE == 6) umma_ss_lo(tb, aug_a, aug_b, idesc_s, 1u, UMMA_DESC_HI_SW32);   // the norm block: 32-byte rows
          else umma_ss_lo(tb, a_lo, b_lo, idesc_s, 1u);
          if (MODE == 6) umma_commit(&dummy[0]);
        }
        if (MODE == 1 || MODE == 3 || MODE == 5 || MODE == 6) {
          for (int kk = 0; kk < 8; ++kk) umma_ts_lo(tb + 384, tb + 128 + kk * 16, bmn_lo + kk * 128, idesc_mn, 1u);
          if (MODE == 6) { umma_commit(&dummy[1]); umma_commit(&dummy[2]); }
        }
        if (MODE == 2 || MODE == 4) {
          for (int kk = 0; kk < 8; ++kk) umma_ts_lo(tb + 384, tb + 128 + kk * 16, b_lo + (kk >> 2) * (BLOCK >> 4) + 2 * (kk & 3), idesc_s, 1u);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp == 1) {
    // the copy stream: copy_kb_per_round KB of bulk copies global -> shared per MMA round
    if (lane == 0 && copy_kb_per_round > 0) {
      uint32_t ph = 0;
      for (int r = 0; r < rounds; ++r) {
        const uint32_t bytes = static_cast<uint32_t>(copy_kb_per_round) * 1024u;
        mbar_arrive_expect_tx(&cbar, bytes);
        for (uint32_t o = 0; o < bytes; o += 16384) bulk_g2s(sD + (o & 16383u), src + (static_cast<size_t>(blockIdx.x) * 64 + (r & 63)) * 32768 + o, 16384, &cbar);
        mbar_wait(&cbar, ph);
        ph ^= 1;
      }
    }
  } else if (warp - 2 < ld_warps) {
    // epilogue-like TMEM traffic: tcgen05.ld / tcgen05.st of 16-column units on this warp's lane quadrant
    uint32_t v[16];
    const uint32_t taddr = tb + ((static_cast<uint32_t>(warp & 3) * 32u) << 16);
    uint32_t acc = 0;
    float2 rs = make_float2(0.f, 0.f);
    const int units = heavy ? rounds * 2 : rounds * 4;      // heavy: 16 warps x 2 units x 16 columns = one 128 x 128 tile per round
    for (int i = 0; i < units; ++i) {
      tmem_ld16(taddr + (i & 7) * 16, v);
      tmem_ld_wait16(v);
      uint32_t pk[8];
      if (heavy) {
        // the exponential-loss unit of the sweeps: FFMA2 -> 2^x (6 pairs MUFU, 2 pairs polynomial) -> FADD2 -> bf16 pack
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const float2 x = ffma2(make_float2(__uint_as_float(v[c]) * 1e-30f, __uint_as_float(v[c + 1]) * 1e-30f), make_float2(0.7f, 0.7f),
                                 make_float2(-1.5f, -1.5f));
          const float2 e = (c >> 1) < 2 ? ex2_poly2(x) : make_float2(ex2f(x.x), ex2f(x.y));
          rs = fadd2(rs, e);
          pk[c >> 1] = pack_bf16x2(e.x, e.y);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) { pk[c] = v[2 * c] + v[2 * c + 1]; acc += pk[c]; }
      }
      tmem_st8(taddr + 256 + (i & 7) * 8, pk);
    }
    acc += __float_as_uint(rs.x + rs.y);
    tmem_st_wait();
    if (acc == 12345u) out[7] = 1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tb);
}


// The issue loop of the sweeps, batch by batch: per tile [wait on an (already complete) mbarrier, tcgen05.fence::after_thread_sync,
// elect, 9 score MMAs, commit] then [the same with the 8 second MMAs].  `flags` switches the pieces on one at a time.
//   1: tcgen05.fence::after_thread_sync per batch   2: elect.sync per batch (else once)   4: tcgen05.commit per batch
//   8: mbarrier try_wait (on a completed phase) per batch   16: __syncwarp per batch
__global__ void batch_bench(int rounds, int flags, int variant, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, done_bar, dummy[2];
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done_bar, 1); mbar_init(&dummy[0], 1); mbar_init(&dummy[1], 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tbase);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (threadIdx.x == 0) mbar_arrive(&done_bar);     // phase 0 of done_bar is complete from now on
  __syncthreads();
  if (warp == 0) {
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0), idesc_mn = umma_idesc_bf16(128, 128, 0, 1);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 16), b_lo = umma_desc_lo(smem_u32(smem + 32768), 16);
    const uint32_t bmn_lo = umma_desc_lo(smem_u32(smem + 32768), BLOCK);
    const bool once = elect_one();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      // variant 0: the sweeps' order (the score MMA of round r+1 overwrites the buffer the second MMA of round r read)
      //         1: second MMA reads the buffer two rounds away from the next score MMA's   2: G tiles in their own TMEM
      //         columns (2 score buffers of 128 + 2 G regions of 64 + accumulator)        3: as 0, accumulate flag always set
      const uint32_t d_s = variant == 2 ? tb + (r & 1) * 128 : tb + (r % 3) * 128;
      const uint32_t a_g = variant == 2 ? tb + 256 + (r & 1) * 64 : tb + ((r + (variant == 1 ? 2 : 1)) % 3) * 128;
      for (int half = 0; half < 2; ++half) {
        if (flags & 8) mbar_wait(&done_bar, 0);
        if (flags & 1) tc_fence_after();
        const bool me = (flags & 2) ? elect_one() : once;
        if (me) {
          if (half == 0) {
            for (int kb = 0; kb < 2; ++kb)
              for (int k = 0; k < 4; ++k) umma_ss_lo(d_s, a_lo + kb * (BLOCK >> 4) + 2 * k, b_lo + kb * (BLOCK >> 4) + 2 * k, idesc_s, (variant == 3 || (kb | k)) ? 1u : 0u);
            umma_ss_lo(d_s, a_lo, b_lo, idesc_s, 1u);
          } else {
            for (int kk = 0; kk < 8; ++kk) umma_ts_lo(tb + 384, a_g + kk * (variant == 2 ? 8 : 16), bmn_lo + kk * 128, idesc_mn, 1u);
          }
          if (flags & 4) umma_commit(&dummy[half]);
        }
        if (flags & 16) __syncwarp();
      }
    }
    if (once) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tb);
}
void run_batch(const char* name, int flags, long long* d, int variant = 0) {
  cudaFuncSetAttribute(batch_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int rounds = 4000;
  batch_bench<<<148, 64, 96 * 1024>>>(rounds, flags, variant, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("batch loop, %-60s: %s  %.0f cycles per tile (17 MMAs)\n", name, cudaGetErrorString(e), (double)h / rounds);
}

template <int MODE>
void run(const char* name, int per_round, int copy_kb, int ld_warps, const uint8_t* src, long long* d, int random_data, int heavy = 0) {
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int rounds = 4000;
  bench<MODE><<<148, 32 * (2 + 16), 96 * 1024>>>(rounds, copy_kb, ld_warps, src, d, random_data, heavy);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s %s %-44s copy %2d KB/round, %2d TMEM warps: %s  %.1f cycles per MMA, %.0f per round of %d\n", random_data ? "random" : "zeros ", heavy ? "epilogue math" : "light warps  ", name, copy_kb, ld_warps,
         cudaGetErrorString(e), (double)h / rounds / per_round, (double)h / rounds, per_round);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  uint8_t* src;
  cudaMalloc(&src, static_cast<size_t>(148) * 64 * 32768 + 65536);
  cudaMemset(src, 0, static_cast<size_t>(148) * 64 * 32768 + 65536);
  run_batch("plain, sweep order (S(r+1) overwrites the buffer GC(r) read)", 0, d, 0);
  run_batch("plain, GC reads a buffer the next S does not touch", 0, d, 1);
  run_batch("plain, G tiles in their own TMEM columns", 0, d, 2);
  run_batch("plain, sweep order, accumulate flag always set", 0, d, 3);
  run_batch("all features, sweep order", 31, d, 0);
  run_batch("all features, GC reads a buffer the next S does not touch", 31, d, 1);
  run_batch("all features, G tiles in their own TMEM columns", 31, d, 2);
  return 0;
}
