/* TEST INFRASTRUCTURE ONLY (oracle build shim) -- not product code.
 * Minimal stand-in for <gsl/gsl_interp.h>: the reference includes it from
 * source/tools/interpolate.h:18.  GSL is not installed in this image, so
 * oracle/shims/gsl_shim.c restates the published natural-cubic-spline
 * algorithm (gsl_interp_cspline) behind the same four entry points.  */
#ifndef PION_ORACLE_GSL_INTERP_SHIM_H
#define PION_ORACLE_GSL_INTERP_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { int kind; } gsl_interp_type;
typedef struct { size_t cache; } gsl_interp_accel;
extern const gsl_interp_type *gsl_interp_cspline;
#ifdef __cplusplus
}
#endif
#endif
