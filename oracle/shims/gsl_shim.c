/* TEST INFRASTRUCTURE ONLY (oracle build shim) -- not product code.
 *
 * Functional stand-in for the four GSL entry points used by the reference's
 * source/tools/interpolate.cpp:59-118 (gsl_spline_alloc/init/eval_e with
 * gsl_interp_cspline).  GSL (version unpinned by the reference: distro
 * libgsl-dev, extra_libraries/install_all_libs.sh:76) is absent here, so this
 * restates its published algorithm: natural cubic spline, second derivatives
 * from the symmetric tridiagonal system, evaluation by binary search +
 * cubic polynomial in (x - x_i).  Spline values are only consumed at SETUP
 * (cooling tables); tests feed the *same tables* to oracle and GPU, so this
 * boundary is "parity unpinned" against real GSL but cannot leak into parity.
 */
#include <stdlib.h>
#include <string.h>
#include "gsl/gsl_spline.h"

static const gsl_interp_type cspline_type = {1};
const gsl_interp_type *gsl_interp_cspline = &cspline_type;

gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size) {
  (void)T;
  gsl_spline *s = (gsl_spline *)calloc(1, sizeof(gsl_spline));
  s->n = size;
  s->x = (double *)calloc(size, sizeof(double));
  s->y = (double *)calloc(size, sizeof(double));
  s->c = (double *)calloc(size, sizeof(double));
  return s;
}

void gsl_spline_free(gsl_spline *s) {
  if (!s) return;
  free(s->x); free(s->y); free(s->c); free(s);
}

/* natural spline: c[0]=c[n-1]=0; for i=1..n-2
 *   h[i-1] c[i-1] + 2(h[i-1]+h[i]) c[i] + h[i] c[i+1]
 *      = 3( (y[i+1]-y[i])/h[i] - (y[i]-y[i-1])/h[i-1] )
 * where c = y''/2. */
int gsl_spline_init(gsl_spline *s, const double *xa, const double *ya, size_t n) {
  if (!s || n != s->n || n < 3) return 1;
  memcpy(s->x, xa, n * sizeof(double));
  memcpy(s->y, ya, n * sizeof(double));
  size_t m = n - 2;
  double *diag = (double *)malloc(m * sizeof(double));
  double *off = (double *)malloc(m * sizeof(double));
  double *g = (double *)malloc(m * sizeof(double));
  for (size_t i = 0; i < m; i++) {
    double h_i = xa[i + 1] - xa[i];
    double h_ip1 = xa[i + 2] - xa[i + 1];
    double yd_i = ya[i + 1] - ya[i];
    double yd_ip1 = ya[i + 2] - ya[i + 1];
    off[i] = h_ip1;
    diag[i] = 2.0 * (h_ip1 + h_i);
    g[i] = 3.0 * (yd_ip1 / h_ip1 - yd_i / h_i);
  }
  /* Thomas algorithm on the symmetric tridiagonal system */
  for (size_t i = 1; i < m; i++) {
    double w = off[i - 1] / diag[i - 1];
    diag[i] -= w * off[i - 1];
    g[i] -= w * g[i - 1];
  }
  s->c[0] = 0.0;
  s->c[n - 1] = 0.0;
  if (m > 0) {
    s->c[m] = g[m - 1] / diag[m - 1];
    for (size_t i = m - 1; i-- > 0;)
      s->c[i + 1] = (g[i] - off[i] * s->c[i + 2]) / diag[i];
  }
  free(diag); free(off); free(g);
  return 0;
}

int gsl_spline_eval_e(const gsl_spline *s, double x, gsl_interp_accel *a, double *y) {
  (void)a;
  size_t n = s->n;
  if (x < s->x[0] || x > s->x[n - 1]) { *y = 0.0; return 1; /* GSL_EDOM */ }
  size_t lo = 0, hi = n - 1;
  while (hi > lo + 1) {
    size_t mid = (lo + hi) / 2;
    if (s->x[mid] > x) hi = mid; else lo = mid;
  }
  double dx = s->x[lo + 1] - s->x[lo];
  double dy = s->y[lo + 1] - s->y[lo];
  double c_i = s->c[lo], c_ip1 = s->c[lo + 1];
  double b_i = dy / dx - dx * (c_ip1 + 2.0 * c_i) / 3.0;
  double d_i = (c_ip1 - c_i) / (3.0 * dx);
  double delx = x - s->x[lo];
  *y = s->y[lo] + delx * (b_i + delx * (c_i + delx * d_i));
  return 0;
}
