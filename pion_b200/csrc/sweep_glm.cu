// glm instantiations of the flux-once sweep stage kernel (stage_sweep.cuh).
#include "stage_sweep_tma.cuh"
namespace pion {
const char* launch_sweep_glm(int solver, int fkj, const StageArgs& a, cudaStream_t s) {
  if (solver == SOLVE_ROE) {
    if (fkj) return launch_sweep_any<EQ_GLM, SOLVE_ROE, true>(a, s);
    else return launch_sweep_any<EQ_GLM, SOLVE_ROE, false>(a, s);
  } else if (solver == SOLVE_HLLD) {
    if (fkj) return launch_sweep_any<EQ_GLM, SOLVE_HLLD, true>(a, s);
    else return launch_sweep_any<EQ_GLM, SOLVE_HLLD, false>(a, s);
  } else {
    if (fkj) return launch_sweep_any<EQ_GLM, SOLVE_HLL, true>(a, s);
    else return launch_sweep_any<EQ_GLM, SOLVE_HLL, false>(a, s);
  }
}
void sweep_tile_cells(int eq, int* cx, int* cy) { sweep_tile_cells_impl(eq, cx, cy); }
void sweep_tma_box(int eq, int order, int ntr, int* cw, int* rh, int* nb, int* tx, int* ty) { sweep_tma_box_impl(eq, order, ntr, cw, rh, nb, tx, ty); }
bool sweep_tma_fits(int eq, int ntr) { return sweep_tma_fits_impl(eq, ntr); }
}  // namespace pion
