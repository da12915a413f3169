#include "sweep_wg_launch.h"
namespace xb {
#define XB_WG_CASE(LMV)                                                                                   \
  case LMV:                                                                                               \
    return logq ? launch_wg_impl(wg_kernel<WG_GRADI, LMV, true>, tmR, tmC, tmRa, tmCa, p, grid, smem, st)       \
                : launch_wg_impl(wg_kernel<WG_GRADI, LMV, false>, tmR, tmC, tmRa, tmCa, p, grid, smem, st);
cudaError_t launch_wg_gradi(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                         const WgParams& p, int grid, size_t smem, cudaStream_t st) {
  switch (lm) {
    XB_WG_CASE(LM_CONTR)
    XB_WG_CASE(LM_INFONCE)
    XB_WG_CASE(LM_MINE)
    XB_WG_CASE(LM_HINGE)
    XB_WG_CASE(LM_LOGI)
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace xb
