// Microbenchmark: issue / execution cost of tcgen05.mma (kind::f16, bf16, M=128, K=16) from one thread, as a function
// of N and of where A comes from (shared memory or TMEM).  One CTA per SM.  Reports cycles per MMA instruction for
// (a) the issue loop alone (clock after the last issue) and (b) until the commit barrier fires (execution).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue_bench mma_issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../matrix-factorization-torch_b200/csrc/ptx.cuh"
using namespace xb;

template <int N, bool TS, int EPIW>
__global__ void bench(int iters, int per_commit, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // operands: garbage is fine (zero-filled smem)
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tbase);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tbase;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint32_t a_lo = umma_desc_lo(smem_u32(smem), 16), b_lo = umma_desc_lo(smem_u32(smem + 16384), 16);
    long long t_issue = 0, t_exec = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      const long long t0 = clock64();
      if (elect_one()) {
        for (int k = 0; k < per_commit; ++k) {
          if (TS) umma_ts_lo(tb, tb + 256 + (k & 7) * 8, b_lo + (k & 3) * 2, idesc, k ? 1u : 0u);
          else umma_ss_lo(tb, a_lo + (k & 3) * 2, b_lo + (k & 3) * 2, idesc, k ? 1u : 0u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      const long long t1 = clock64();
      mbar_wait(&bar, ph); ph ^= 1;
      const long long t2 = clock64();
      t_issue += t1 - t0; t_exec += t2 - t0;
    }
    if (lane == 0 && blockIdx.x == 0) { out[0] = t_issue; out[1] = t_exec; }
  } else if (EPIW > 0) {
    // competing math warps (FMA + MUFU), to see whether issue slows down under load
    float x = threadIdx.x * 1e-3f, y = 0.f;
    for (int i = 0; i < iters * per_commit * 8; ++i) { x = fmaf(x, 1.0001f, 0.5f); y += ex2f(-x); }
    if (y == 123.f) out[5] = 1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tb);
}

template <int N, bool TS, int EPIW>
void run(const char* name, long long* d) {
  cudaFuncSetAttribute(bench<N, TS, EPIW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int pc : {8, 9, 16, 64}) {
    const int iters = 200;
    bench<N, TS, EPIW><<<148, 32 * (1 + EPIW), 64 * 1024>>>(iters, pc, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-28s per_commit %2d: %s  issue %.1f cyc/MMA  issue+exec %.1f cyc/MMA\n", name, pc, cudaGetErrorString(e),
           (double)h[0] / iters / pc, (double)h[1] / iters / pc);
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  run<128, false, 0>("N=128 SS idle", d);
  run<256, false, 0>("N=256 SS idle", d);
  run<128, true, 0>("N=128 TS idle", d);
  run<128, false, 16>("N=128 SS +16 math warps", d);
  run<256, false, 16>("N=256 SS +16 math warps", d);
  run<128, true, 16>("N=128 TS +16 math warps", d);
  return 0;
}
