#include "sweep_launch.h"
namespace xb {
#define XB_FWD_CASE(LMV)                                                                                   \
  case LMV:                                                                                                \
    return logq ? launch_sweep_impl(sweep_kernel<MODE_FWD, LMV, true, true>, MODE_FWD, LMV, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st)  \
                : launch_sweep_impl(sweep_kernel<MODE_FWD, LMV, true, false>, MODE_FWD, LMV, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st);
cudaError_t launch_sweep_fwd(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                             const SweepParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  switch (lm) {
    XB_FWD_CASE(LM_CONTR)
    XB_FWD_CASE(LM_INFONCE)
    XB_FWD_CASE(LM_MINE)
    XB_FWD_CASE(LM_HINGE)
    XB_FWD_CASE(LM_LOGI)
    XB_FWD_CASE(LM_ALL)
    default: return cudaErrorInvalidValue;
  }
}
#define XB_FWDQ_CASE(LMV)                                                                                  \
  case LMV:                                                                                                \
    return logq ? launch_sweep_impl(sweep_kernel<MODE_FWDQ, LMV, true, true>, MODE_FWDQ, LMV, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st)  \
                : launch_sweep_impl(sweep_kernel<MODE_FWDQ, LMV, true, false>, MODE_FWDQ, LMV, true, tmR, tmC, tmRa, tmCa, p, grid, smem, st);
cudaError_t launch_sweep_fwdq(int lm, bool logq, const CUtensorMap& tmR, const CUtensorMap& tmC, const CUtensorMap& tmRa, const CUtensorMap& tmCa,
                              const SweepParams& p, dim3 grid, size_t smem, cudaStream_t st) {
  switch (lm) {
    XB_FWDQ_CASE(LM_CONTR)
    XB_FWDQ_CASE(LM_INFONCE)
    XB_FWDQ_CASE(LM_MINE)
    XB_FWDQ_CASE(LM_HINGE)
    XB_FWDQ_CASE(LM_LOGI)
    default: return cudaErrorInvalidValue;
  }
}
}  // namespace xb
