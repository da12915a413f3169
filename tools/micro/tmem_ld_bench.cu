// Microbenchmark: throughput of tcgen05.ld (TMEM -> registers) per SM, as a function of the number of warps
// and of how many loads are in flight before tcgen05.wait::ld.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../matrix-factorization-torch_b200/csrc/ptx.cuh"
using namespace xb;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}

template <int MODE>
__global__ void bench(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tbase);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)(((i * 2 + (warp >> 2)) * 32) & 511) & ~31u;
    if (MODE == 0) {            // one x32 load, wait
      uint32_t v[32]; tmem_ld32(base + (col & 480), v); tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j];
    } else if (MODE == 1) {     // two x32 loads in flight, wait
      uint32_t v[32], w[32]; tmem_ld32(base + (col & 480), v); tmem_ld32(base + ((col + 32) & 480), w); tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j];
    } else {                    // two x16 loads, wait
      uint32_t v[16], w[16]; tmem_ld16(base + (col & 480), v); tmem_ld16(base + (col & 480) + 16, w); tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= v[j] ^ w[j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

int main() {
  long long* d_c; uint32_t* d_s; cudaMalloc(&d_c, 148 * 8); cudaMalloc(&d_s, 148 * 1024 * 4);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16}) {
      if (mode == 0) bench<0><<<148, warps * 32>>>(iters, d_c, d_s);
      if (mode == 1) bench<1><<<148, warps * 32>>>(iters, d_c, d_s);
      if (mode == 2) bench<2><<<148, warps * 32>>>(iters, d_c, d_s);
      cudaError_t e = cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
      const double bytes_per_iter = (mode == 1 ? 2.0 : 1.0) * warps * 32 * 32 * 4;   // per CTA (= per SM)
      printf("mode %d warps %2d: %s  %.1f cycles/iter  -> %.1f B/clk/SM  (64 KB tile in %.0f cycles)\n", mode, warps,
             cudaGetErrorString(e), (double)c / iters, bytes_per_iter * iters / c, 65536.0 / (bytes_per_iter * iters / c));
    }
  return 0;
}
