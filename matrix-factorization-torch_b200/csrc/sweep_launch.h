// Host-side launch shims for the sweep kernel instantiations (one translation unit per family so the
// template instantiations compile in parallel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "sweep.cuh"

namespace xb {
cudaError_t launch_sweep_fwd(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                             const SweepParams& p, dim3 grid, size_t smem, cudaStream_t st);
cudaError_t launch_sweep_grad_qrow(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                                   const SweepParams& p, dim3 grid, size_t smem, cudaStream_t st);
cudaError_t launch_sweep_grad_qcol(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                                   const SweepParams& p, dim3 grid, size_t smem, cudaStream_t st);
cudaError_t launch_sweep_fwdq(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                             const SweepParams& p, dim3 grid, size_t smem, cudaStream_t st);
cudaError_t launch_sweep_topk(bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa, const SweepParams& p,
                              dim3 grid, size_t smem, cudaStream_t st);
cudaError_t launch_sweep_debug(const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa, const SweepParams& p, dim3 grid,
                               size_t smem, cudaStream_t st);

template <typename K>
inline cudaError_t launch_sweep_impl(K kernel, int mode, int lm, bool qrow, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                                     const SweepParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kernel<<<grid, sweep_threads(mode, lm, qrow), smem, st>>>(tmR, tmC, tmRa, tmCa, p);
  return cudaGetLastError();
}
}  // namespace xb
