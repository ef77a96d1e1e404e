/* mini-GSL shim, TEST INFRASTRUCTURE: the slice of <gsl/gsl_spline.h> src/convert_pt.c uses
 * (gsl_interp_cspline = natural cubic spline; alloc / init / eval / free, accelerator objects).
 * Restated from the published algorithm (GSL 2.x interpolation/cspline.c: tridiagonal system for
 * the second derivatives with c[0] = c[n-1] = 0).  Header-only so that a plain `gcc convert_pt.c`
 * links without a GSL library. */
#ifndef RTRG_SHIM_GSL_SPLINE_H
#define RTRG_SHIM_GSL_SPLINE_H
#include <stdlib.h>
#include <string.h>

typedef struct { int cspline; } gsl_interp_type;
static const gsl_interp_type rtrg_shim_cspline = {1};
static const gsl_interp_type *gsl_interp_cspline = &rtrg_shim_cspline;
typedef struct { size_t cache; } gsl_interp_accel;
typedef struct {
  size_t size;
  double *x, *y, *c; /* c: half the second derivative at the nodes */
} gsl_spline;

static inline gsl_interp_accel *gsl_interp_accel_alloc(void) { return (gsl_interp_accel *)calloc(1, sizeof(gsl_interp_accel)); }
static inline void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }
static inline gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size) {
  gsl_spline *s = (gsl_spline *)calloc(1, sizeof(gsl_spline));
  (void)T;
  s->size = size;
  s->x = (double *)malloc(size * sizeof(double));
  s->y = (double *)malloc(size * sizeof(double));
  s->c = (double *)calloc(size, sizeof(double));
  return s;
}
static inline void gsl_spline_free(gsl_spline *s) {
  if (!s) return;
  free(s->x), free(s->y), free(s->c), free(s);
}
static inline int gsl_spline_init(gsl_spline *s, const double xa[], const double ya[], size_t size) {
  size_t i, n = size;
  double *g, *diag, *off;
  if (size != s->size) return 4; /* GSL_EINVAL */
  for (i = 1; i < n; i++)
    if (!(xa[i] > xa[i - 1])) return 4; /* x values must be strictly increasing */
  memcpy(s->x, xa, n * sizeof(double));
  memcpy(s->y, ya, n * sizeof(double));
  memset(s->c, 0, n * sizeof(double));
  if (n < 3) return 0;
  g = (double *)malloc(n * sizeof(double)), diag = (double *)malloc(n * sizeof(double)), off = (double *)malloc(n * sizeof(double));
  for (i = 0; i + 2 < n; i++) { /* rows for the interior nodes 1..n-2 */
    const double h0 = xa[i + 1] - xa[i], h1 = xa[i + 2] - xa[i + 1];
    off[i] = h1;
    diag[i] = 2.0 * (h0 + h1);
    g[i] = 3.0 * ((ya[i + 2] - ya[i + 1]) / h1 - (ya[i + 1] - ya[i]) / h0);
  }
  for (i = 1; i + 2 < n; i++) { /* Thomas algorithm (symmetric tridiagonal) */
    const double w = off[i - 1] / diag[i - 1];
    diag[i] -= w * off[i - 1];
    g[i] -= w * g[i - 1];
  }
  for (i = n - 2; i-- > 0;) {
    const double next = (i + 3 < n) ? s->c[i + 2] : 0.0;
    s->c[i + 1] = (g[i] - off[i] * next) / diag[i];
  }
  free(g), free(diag), free(off);
  return 0;
}
static inline double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a) {
  size_t lo = 0, hi = s->size - 1;
  (void)a;
  while (hi > lo + 1) {
    const size_t mid = (lo + hi) / 2;
    if (s->x[mid] > x) hi = mid; else lo = mid;
  }
  {
    const double h = s->x[lo + 1] - s->x[lo], d = x - s->x[lo];
    const double b = (s->y[lo + 1] - s->y[lo]) / h - h * (s->c[lo + 1] + 2.0 * s->c[lo]) / 3.0;
    const double dd = (s->c[lo + 1] - s->c[lo]) / (3.0 * h);
    return s->y[lo] + d * (b + d * (s->c[lo] + d * dd));
  }
}
#endif
