// Host-side launch shims of the warpgroup-per-tile sweep (sweep_wg.cuh), one translation unit per mode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "sweep_rt.cuh"
#include "sweep_wg.cuh"

namespace xb {
cudaError_t launch_wg_fwdq(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                           const WgParams& p, int grid, size_t smem, cudaStream_t st);
cudaError_t launch_wg_gradi(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                            const WgParams& p, int grid, size_t smem, cudaStream_t st);

cudaError_t launch_rt(const CUtensorMap& tmR, const CUtensorMap& tmC, const RtParams& p, int grid, size_t smem, cudaStream_t st);

template <typename K>
inline cudaError_t launch_wg_impl(K kernel, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                                  const WgParams& p, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kernel<<<grid, WG_THREADS, smem, st>>>(tmR, tmC, tmRa, tmCa, p);
  return cudaGetLastError();
}
}  // namespace xb
