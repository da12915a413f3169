#include "sweep_launch.h"
namespace xb {
cudaError_t launch_sweep_topk(bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa, const SweepParams& p,
                              dim3 grid, size_t smem, cudaStream_t st) {
  // LM = 1 selects the mining key order (p.topk_mining = 1 / 2), LM = 0 plain scores (retrieval)
  if (p.topk_mining != 0)
    return logq ? launch_sweep_impl(sweep_kernel<MODE_TOPK, 1, true, true>, MODE_TOPK, 1, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st)
                : launch_sweep_impl(sweep_kernel<MODE_TOPK, 1, true, false>, MODE_TOPK, 1, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st);
  return launch_sweep_impl(sweep_kernel<MODE_TOPK, 0, true, false>, MODE_TOPK, 0, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st);
}
cudaError_t launch_sweep_debug(const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa, const SweepParams& p, dim3 grid,
                               size_t smem, cudaStream_t st) {
  return launch_sweep_impl(sweep_kernel<MODE_DEBUG, 0, true, false>, MODE_DEBUG, 0, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st);
}
}  // namespace xb
