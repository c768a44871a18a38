/* TEST INFRASTRUCTURE ONLY (oracle build shim) -- see gsl_interp.h. */
#ifndef PION_ORACLE_GSL_SPLINE_SHIM_H
#define PION_ORACLE_GSL_SPLINE_SHIM_H
#include "gsl/gsl_interp.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  size_t n;
  double *x, *y, *c; /* knots, values, second-derivative coefficients */
} gsl_spline;
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t size);
int gsl_spline_eval_e(const gsl_spline *s, double x, gsl_interp_accel *a, double *y);
void gsl_spline_free(gsl_spline *s);
#ifdef __cplusplus
}
#endif
#endif
