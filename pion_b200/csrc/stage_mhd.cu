// MHD instantiations of the stage kernel (HLL, HLLD, Roe-CV; FKJ98 on/off).
#include "stage_kernel.cuh"
namespace pion {
const char* launch_stage_mhd(int solver, int fkj, const StageArgs& a, cudaStream_t s) {
  if (solver == SOLVE_LF) {
    if (fkj) return launch_stage_t<EQ_MHD, SOLVE_LF, true>(a, s);
    else return launch_stage_t<EQ_MHD, SOLVE_LF, false>(a, s);
  }
  if (solver == SOLVE_RSLINEAR) {
    if (fkj) return launch_stage_t<EQ_MHD, SOLVE_RSLINEAR, true>(a, s);
    else return launch_stage_t<EQ_MHD, SOLVE_RSLINEAR, false>(a, s);
  }
  if (solver == SOLVE_ROE) {
    if (fkj) return launch_stage_t<EQ_MHD, SOLVE_ROE, true>(a, s);
    else return launch_stage_t<EQ_MHD, SOLVE_ROE, false>(a, s);
  } else if (solver == SOLVE_HLLD) {
    if (fkj) return launch_stage_t<EQ_MHD, SOLVE_HLLD, true>(a, s);
    else return launch_stage_t<EQ_MHD, SOLVE_HLLD, false>(a, s);
  } else {
    if (fkj) return launch_stage_t<EQ_MHD, SOLVE_HLL, true>(a, s);
    else return launch_stage_t<EQ_MHD, SOLVE_HLL, false>(a, s);
  }
}
}  // namespace pion
