#include "sweep_rt.cuh"
namespace xb {
cudaError_t launch_rt(const CUtensorMap& tmR, const CUtensorMap& tmC, const RtParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kernel = p.mask != nullptr ? rt_kernel<true> : rt_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kernel<<<grid, RT_THREADS, smem, st>>>(tmR, tmC, p);
  return cudaGetLastError();
}
}  // namespace xb
