#include "sweep_launch.h"
namespace xb {
cudaError_t launch_sweep_topk(bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa, const SweepParams& p,
                              dim3 grid, size_t smem, cudaStream_t st) {
  // LM = 1 selects the mining key orders (MINE = p.topk_mining, compiled in), LM = 0 plain scores (retrieval)
#define XB_MINE_CASE(M)                                                                                                    \
  case M:                                                                                                                  \
    return logq ? launch_sweep_impl(sweep_kernel<MODE_TOPK, 1, true, true, M>, MODE_TOPK, 1, true, tmR, tmC, tmRa, tmCa, p, \
                                    grid, smem, st)                                                                        \
                : launch_sweep_impl(sweep_kernel<MODE_TOPK, 1, true, false, M>, MODE_TOPK, 1, true, tmR, tmC, tmRa, tmCa,   \
                                    p, grid, smem, st);
  switch (p.topk_mining) {
    XB_MINE_CASE(1)
    XB_MINE_CASE(2)
    XB_MINE_CASE(3)
    XB_MINE_CASE(4)
    default: break;
  }
#undef XB_MINE_CASE
  return launch_sweep_impl(sweep_kernel<MODE_TOPK, 0, true, false>, MODE_TOPK, 0, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st);
}
cudaError_t launch_sweep_debug(const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa, const SweepParams& p, dim3 grid,
                               size_t smem, cudaStream_t st) {
  return launch_sweep_impl(sweep_kernel<MODE_DEBUG, 0, true, false>, MODE_DEBUG, 0, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st);
}
}  // namespace xb
