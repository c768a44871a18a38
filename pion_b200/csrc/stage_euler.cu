// Euler instantiations of the stage kernel (HLL, Roe-CV; FKJ98 on/off).
#include "stage_kernel.cuh"
namespace pion {
const char* launch_stage_euler(int solver, int fkj, const StageArgs& a, cudaStream_t s) {
  if (solver == SOLVE_LF) {
    if (fkj) return launch_stage_t<EQ_EULER, SOLVE_LF, true>(a, s);
    else return launch_stage_t<EQ_EULER, SOLVE_LF, false>(a, s);
  }
  if (solver == SOLVE_RSLINEAR) {
    if (fkj) return launch_stage_t<EQ_EULER, SOLVE_RSLINEAR, true>(a, s);
    else return launch_stage_t<EQ_EULER, SOLVE_RSLINEAR, false>(a, s);
  }
  if (solver == SOLVE_RSEXACT) {
    if (fkj) return launch_stage_t<EQ_EULER, SOLVE_RSEXACT, true>(a, s);
    else return launch_stage_t<EQ_EULER, SOLVE_RSEXACT, false>(a, s);
  }
  if (solver == SOLVE_RSHYBRID) {
    if (fkj) return launch_stage_t<EQ_EULER, SOLVE_RSHYBRID, true>(a, s);
    else return launch_stage_t<EQ_EULER, SOLVE_RSHYBRID, false>(a, s);
  }
  if (solver == SOLVE_ROE) {
    if (fkj) return launch_stage_t<EQ_EULER, SOLVE_ROE, true>(a, s);
    else return launch_stage_t<EQ_EULER, SOLVE_ROE, false>(a, s);
  } else if (solver == SOLVE_FVS) {
    if (fkj) return launch_stage_t<EQ_EULER, SOLVE_FVS, true>(a, s);
    else return launch_stage_t<EQ_EULER, SOLVE_FVS, false>(a, s);
  } else if (solver == SOLVE_ROE_PV) {
    if (fkj) return launch_stage_t<EQ_EULER, SOLVE_ROE_PV, true>(a, s);
    else return launch_stage_t<EQ_EULER, SOLVE_ROE_PV, false>(a, s);
  } else {
    if (fkj) return launch_stage_t<EQ_EULER, SOLVE_HLL, true>(a, s);
    else return launch_stage_t<EQ_EULER, SOLVE_HLL, false>(a, s);
  }
}
}  // namespace pion
