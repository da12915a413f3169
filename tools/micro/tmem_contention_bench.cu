// Microbenchmark: tcgen05.ld throughput while the tensor core is busy.  Warp 16 issues back-to-back tcgen05.mma
// (M=128, N=128, K=16, bf16, SS) into TMEM columns [0,128); 16 other warps stream tcgen05.ld.x16 over columns
// [256,512).  Reports cycles per MMA and TMEM read bytes per clock, alone and together.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_contention_bench tmem_contention_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../matrix-factorization-torch_b200/csrc/ptx.cuh"
using namespace xb;

template <bool DO_MMA, bool DO_LD, bool TS>
__global__ void bench(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 16) tmem_alloc<512>(&tbase);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (warp == 16) {
    if (DO_MMA) {
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 16), b_lo = umma_desc_lo(smem_u32(smem + 16384), 16);
      const long long t0 = clock64();
      if (elect_one()) {
        for (int k = 0; k < iters; ++k) {
          if (TS) umma_ts_lo(tb, tb + 128 + (k & 7) * 8, b_lo + (k & 3) * 2, idesc, 1u);
          else umma_ss_lo(tb, a_lo + (k & 3) * 2, b_lo + (k & 3) * 2, idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
  } else if (DO_LD) {
    const uint32_t base = tb + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    uint32_t acc = 0;
    const int n = iters;   // one x16 load per MMA-equivalent step
    const long long t0 = clock64();
    uint32_t va[16], vb[16];
    tmem_ld16(base, va);
    for (int i = 0; i < n; i += 2) {
      tmem_ld_wait16(va);
      tmem_ld16(base + (((i + 1) * 16) & 255), vb);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= va[j];
      tmem_ld_wait16(vb);
      tmem_ld16(base + (((i + 2) * 16) & 255), va);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= vb[j];
    }
    tmem_ld_wait16(va);
    const long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[1 + warp] = t1 - t0;
    if (acc == 0x12345) out[40] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 16) tmem_dealloc<512>(tb);
}

template <bool DO_MMA, bool DO_LD, bool TS>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(bench<DO_MMA, DO_LD, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 4000;
  cudaMemset(d, 0, 64 * 8);
  bench<DO_MMA, DO_LD, TS><<<148, 17 * 32, 64 * 1024>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[20]; cudaMemcpy(h, d, 20 * 8, cudaMemcpyDeviceToHost);
  long long ldmax = 0; for (int w = 0; w < 16; ++w) ldmax = h[1 + w] > ldmax ? h[1 + w] : ldmax;
  printf("%-34s %s", name, cudaGetErrorString(e));
  if (DO_MMA) printf("  MMA: %.1f cyc/MMA", (double)h[0] / iters);
  if (DO_LD) printf("  LD: 16 warps x %d x 2 KB in %lld cyc -> %.1f B/clk/SM (64 KB tile in %.0f cyc)", iters, ldmax,
                    16.0 * iters * 2048 / ldmax, 65536.0 / (16.0 * iters * 2048 / ldmax));
  printf("\n");
}
int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  run<true, false, false>("MMA (SS) alone", d);
  run<false, true, false>("tcgen05.ld alone", d);
  run<true, true, false>("MMA (SS) + tcgen05.ld", d);
  run<true, false, true>("MMA (TS: A from TMEM) alone", d);
  run<true, true, true>("MMA (TS) + tcgen05.ld", d);
  return 0;
}
