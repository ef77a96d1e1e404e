/* Mini-GSL shim -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product).
 * GSL is an un-vendored, unpinned dependency of the reference (CMakeLists.txt:9,
 * src/Makefile:5) and is absent from this image; this header declares exactly the
 * symbols src/redTime.cc and src/AU_cosmological_parameters.h use so that the
 * UNMODIFIED reference sources compile.  Algorithms restated in ../gsl_shim.cc. */
#ifndef SHIM_GSL_INTEGRATION_H
#define SHIM_GSL_INTEGRATION_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { double (*function)(double x, void *params); void *params; } gsl_function;
typedef struct {
  size_t limit, size, nrmax, i, maximum_level;
  double *alist, *blist, *rlist, *elist;
  size_t *order, *level;
} gsl_integration_workspace;
gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
/* adaptive Gauss-Kronrod; only key 6 (61-point) is used by the reference
 * (AU_cosmological_parameters.h:757,865,957) and implemented here. */
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *workspace,
                        double *result, double *abserr);
/* instrumentation (shim-only): counters for tests / BASELINE work counts */
extern long shim_qag_calls, shim_qag_intervals, shim_qag_fevals;
#ifdef __cplusplus
}
#endif
#endif
