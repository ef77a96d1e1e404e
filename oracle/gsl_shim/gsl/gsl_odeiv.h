/* Mini-GSL shim -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product).
 * GSL is an un-vendored, unpinned dependency of the reference (CMakeLists.txt:9,
 * src/Makefile:5) and is absent from this image; this header declares exactly the
 * symbols src/redTime.cc and src/AU_cosmological_parameters.h use so that the
 * UNMODIFIED reference sources compile.  Algorithms restated in ../gsl_shim.cc. */
#ifndef SHIM_GSL_ODEIV_H
#define SHIM_GSL_ODEIV_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
  int (*function)(double t, const double y[], double dydt[], void *params);
  int (*jacobian)(double t, const double y[], double *dfdy, double dfdt[], void *params);
  size_t dimension;
  void *params;
} gsl_odeiv_system;

typedef struct gsl_odeiv_step_type_s {
  const char *name;
  int can_use_dydt_in;
  int n_stages;      /* shim: 6 (rkf45) or 13 (rk8pd) */
  unsigned int order;
} gsl_odeiv_step_type;
typedef struct { const gsl_odeiv_step_type *type; size_t dimension; double *work; } gsl_odeiv_step;
typedef struct { double eps_abs, eps_rel, a_y, a_dydt; } gsl_odeiv_control;
typedef struct {
  size_t dimension;
  double *y0, *yerr, *dydt_in, *dydt_out;
  double last_step;
  unsigned long count, failed_steps;
} gsl_odeiv_evolve;

extern const gsl_odeiv_step_type *gsl_odeiv_step_rkf45;  /* redTime.cc:1591 */
extern const gsl_odeiv_step_type *gsl_odeiv_step_rk8pd;  /* AU_cosmological_parameters.h:172 */

gsl_odeiv_step *gsl_odeiv_step_alloc(const gsl_odeiv_step_type *T, size_t dim);
void gsl_odeiv_step_free(gsl_odeiv_step *s);
gsl_odeiv_control *gsl_odeiv_control_y_new(double eps_abs, double eps_rel);
void gsl_odeiv_control_free(gsl_odeiv_control *c);
gsl_odeiv_evolve *gsl_odeiv_evolve_alloc(size_t dim);
void gsl_odeiv_evolve_free(gsl_odeiv_evolve *e);
int gsl_odeiv_evolve_apply(gsl_odeiv_evolve *e, gsl_odeiv_control *con, gsl_odeiv_step *step,
                           const gsl_odeiv_system *dydt, double *t, double t1, double *h, double y[]);

/* instrumentation (shim-only).  shim_ode_trace != NULL: one line per attempted step of
 * systems with dimension > 2 (the main Time-RG system) is appended to that FILE. */
extern long shim_ode_attempts[2], shim_ode_rejects[2], shim_ode_rhs[2]; /* [0]=rkf45 [1]=rk8pd */
#ifdef __cplusplus
}
#endif
#endif
