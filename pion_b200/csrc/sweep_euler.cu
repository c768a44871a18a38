// euler instantiations of the flux-once sweep stage kernel (stage_sweep.cuh).
#include "stage_sweep_tma.cuh"
namespace pion {
const char* launch_sweep_euler(int solver, int fkj, const StageArgs& a, cudaStream_t s) {
  if (solver == SOLVE_ROE) {
    if (fkj) return launch_sweep_any<EQ_EULER, SOLVE_ROE, true>(a, s);
    else return launch_sweep_any<EQ_EULER, SOLVE_ROE, false>(a, s);
  } else if (solver == SOLVE_FVS) {
    if (fkj) return launch_sweep_any<EQ_EULER, SOLVE_FVS, true>(a, s);
    else return launch_sweep_any<EQ_EULER, SOLVE_FVS, false>(a, s);
  } else if (solver == SOLVE_ROE_PV) {
    if (fkj) return launch_sweep_any<EQ_EULER, SOLVE_ROE_PV, true>(a, s);
    else return launch_sweep_any<EQ_EULER, SOLVE_ROE_PV, false>(a, s);
  } else {
    if (fkj) return launch_sweep_any<EQ_EULER, SOLVE_HLL, true>(a, s);
    else return launch_sweep_any<EQ_EULER, SOLVE_HLL, false>(a, s);
  }
}
}  // namespace pion
