/* TEST INFRASTRUCTURE ONLY (oracle build shim) -- not product code.
 * Abort-on-call SUNDIALS stubs; see sundials/sundials_types.h. */
#include <stdio.h>
#include <stdlib.h>
#include "sundials/sundials_types.h"
static void die(const char *f) {
  fprintf(stderr, "oracle shim: SUNDIALS call %s is out of scope (no CVODE in this build)\n", f);
  abort();
}
N_Vector N_VNew_Serial(sunindextype n) { (void)n; die("N_VNew_Serial"); return 0; }
void N_VDestroy_Serial(N_Vector v) { (void)v; }
void N_VDestroy(N_Vector v) { (void)v; }
SUNMatrix SUNDenseMatrix(sunindextype M, sunindextype N) { (void)M; (void)N; die("SUNDenseMatrix"); return 0; }
SUNLinearSolver SUNLinSol_Dense(N_Vector y, SUNMatrix A) { (void)y; (void)A; die("SUNLinSol_Dense"); return 0; }
SUNLinearSolver SUNDenseLinearSolver(N_Vector y, SUNMatrix A) { (void)y; (void)A; die("SUNDenseLinearSolver"); return 0; }
void *CVodeCreate(int lmm) { (void)lmm; die("CVodeCreate"); return 0; }
int CVodeInit(void *m, CVRhsFn f, realtype t0, N_Vector y0) { (void)m; (void)f; (void)t0; (void)y0; die("CVodeInit"); return 1; }
int CVodeReInit(void *m, realtype t0, N_Vector y0) { (void)m; (void)t0; (void)y0; die("CVodeReInit"); return 1; }
int CVodeSVtolerances(void *m, realtype r, N_Vector a) { (void)m; (void)r; (void)a; die("CVodeSVtolerances"); return 1; }
int CVodeSetLinearSolver(void *m, SUNLinearSolver L, SUNMatrix A) { (void)m; (void)L; (void)A; die("CVodeSetLinearSolver"); return 1; }
int CVodeSetJacFn(void *m, CVLsJacFn j) { (void)m; (void)j; die("CVodeSetJacFn"); return 1; }
int CVodeSetUserData(void *m, void *u) { (void)m; (void)u; die("CVodeSetUserData"); return 1; }
int CVodeSetMaxNumSteps(void *m, long int n) { (void)m; (void)n; die("CVodeSetMaxNumSteps"); return 1; }
int CVode(void *m, realtype t, N_Vector y, realtype *tr, int it) { (void)m; (void)t; (void)y; (void)tr; (void)it; die("CVode"); return 1; }
void CVodeFree(void **m) { (void)m; }
