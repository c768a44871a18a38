/* TEST INFRASTRUCTURE ONLY (oracle build shim) -- not product code.
 * Abort-on-call stand-ins for the SUNDIALS/CVODE API (built with -DCVODE5).
 * The reference's grid/setup_fixed_grid.cpp hard-includes the MPv3/5/6/7/8
 * chemistry headers, which include microphysics/cvode_integrator.h:60-85.
 * None of the in-scope configurations instantiates those classes; any call
 * into this shim aborts.  */
#ifndef PION_ORACLE_SUNDIALS_SHIM_H
#define PION_ORACLE_SUNDIALS_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef double realtype;
typedef int booleantype;
typedef long int sunindextype;
struct pion_shim_nvector { long int length; double *data; };
typedef struct pion_shim_nvector *N_Vector;
struct pion_shim_sunmatrix { long int M, N; double *data; };
typedef struct pion_shim_sunmatrix *SUNMatrix;
typedef void *SUNLinearSolver;
#define NV_Ith_S(v, i) ((v)->data[i])
#define NV_DATA_S(v) ((v)->data)
#define NV_LENGTH_S(v) ((v)->length)
#define SM_ELEMENT_D(A, i, j) ((A)->data[(j) * (A)->M + (i)])
#define CV_SUCCESS 0
#define CV_BDF 2
#define CV_NORMAL 1
typedef int (*CVRhsFn)(realtype, N_Vector, N_Vector, void *);
typedef int (*CVLsJacFn)(realtype, N_Vector, N_Vector, SUNMatrix, void *, N_Vector, N_Vector, N_Vector);
N_Vector N_VNew_Serial(sunindextype n);
void N_VDestroy_Serial(N_Vector v);
void N_VDestroy(N_Vector v);
SUNMatrix SUNDenseMatrix(sunindextype M, sunindextype N);
SUNLinearSolver SUNLinSol_Dense(N_Vector y, SUNMatrix A);
SUNLinearSolver SUNDenseLinearSolver(N_Vector y, SUNMatrix A);
void *CVodeCreate(int lmm);
int CVodeInit(void *mem, CVRhsFn f, realtype t0, N_Vector y0);
int CVodeReInit(void *mem, realtype t0, N_Vector y0);
int CVodeSVtolerances(void *mem, realtype reltol, N_Vector abstol);
int CVodeSetLinearSolver(void *mem, SUNLinearSolver LS, SUNMatrix A);
int CVodeSetJacFn(void *mem, CVLsJacFn jac);
int CVodeSetUserData(void *mem, void *user_data);
int CVodeSetMaxNumSteps(void *mem, long int mxsteps);
int CVode(void *mem, realtype tout, N_Vector yout, realtype *tret, int itask);
void CVodeFree(void **mem);
#ifdef __cplusplus
}
#endif
#endif
