#include "sweep_launch.h"
namespace xb {
#define XB_GRAD_CASE(LMV)                                                                                     \
  case LMV:                                                                                                   \
    return logq ? launch_sweep_impl(sweep_kernel<MODE_GRAD, LMV, false, true>, MODE_GRAD, LMV, false, tmR, tmC, tmRa, tmCa, p, grid, smem, st)     \
                : launch_sweep_impl(sweep_kernel<MODE_GRAD, LMV, false, false>, MODE_GRAD, LMV, false, tmR, tmC, tmRa, tmCa, p, grid, smem, st);
cudaError_t launch_sweep_grad_qcol(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                                   const SweepParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  switch (lm) {
    XB_GRAD_CASE(LM_CONTR)
    XB_GRAD_CASE(LM_INFONCE)
    XB_GRAD_CASE(LM_MINE)
    XB_GRAD_CASE(LM_HINGE)
    XB_GRAD_CASE(LM_LOGI)
    XB_GRAD_CASE(LM_ALL)
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace xb
