// Mini-GSL shim -- ORACLE / TEST INFRASTRUCTURE ONLY.
//
// The reference (michaelbuehlmann/redTime) links against GSL (un-vendored, version
// unpinned: CMakeLists.txt:9 `find_package(GSL REQUIRED)`, src/Makefile:5 `-lgsl
// -lgslcblas`).  GSL is not installed in this image and there is no network, so the
// oracle restates the published algorithms of the handful of GSL routines the hot
// path uses (GSL 2.x semantics, legacy gsl_odeiv v1 API):
//
//   gsl_odeiv_step_rkf45 / rk8pd, gsl_odeiv_control_y_new, gsl_odeiv_evolve_apply
//       -> redTime.cc:1591-1616, AU_cosmological_parameters.h:172-188
//   gsl_integration_qag (key 6 = 61-point Gauss-Kronrod, QUADPACK dqage)
//       -> AU_cosmological_parameters.h:865,957
//   gsl_fft_{real,halfcomplex,complex}_radix2_*   -> redTime.cc:360-392
//   gsl_sf_lngamma_complex_e (Lanczos g=7 + reflection) -> redTime.cc:310-313
//
// Pinning: the UNMODIFIED reference sources compiled against this shim must reproduce
// the reference's own golden output examples/1_redTime/example_redTime_result.dat
// (tests/test_oracle_golden.py).  Nothing here is ever linked into the product library.
#include <cfloat>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gsl/gsl_errno.h"
#include "gsl/gsl_fft_complex.h"
#include "gsl/gsl_fft_halfcomplex.h"
#include "gsl/gsl_fft_real.h"
#include "gsl/gsl_integration.h"
#include "gsl/gsl_odeiv.h"
#include "gsl/gsl_sf_gamma.h"

#include "gk61_table.h"

// ---------------------------------------------------------------------------------
// ODE steppers (explicit embedded Runge-Kutta, tableau driven)
// ---------------------------------------------------------------------------------
namespace {

struct Tableau {
  int stages;
  const double *c;        // nodes c[1..stages-1] (c[0] = 0)
  const double *const *a; // a[s] = row for stage s (s >= 1), length s
  const double *b;        // propagated solution weights, length stages
  const double *e;        // error weights (yerr = h * sum e_i k_i), length stages
};

// --- Fehlberg 4(5); the 5th-order solution is propagated (GSL rkf45.c) ------------
const double f_c[6] = {0.0, 1.0 / 4.0, 3.0 / 8.0, 12.0 / 13.0, 1.0, 1.0 / 2.0};
const double f_a1[] = {1.0 / 4.0};
const double f_a2[] = {3.0 / 32.0, 9.0 / 32.0};
const double f_a3[] = {1932.0 / 2197.0, -7200.0 / 2197.0, 7296.0 / 2197.0};
const double f_a4[] = {8341.0 / 4104.0, -32832.0 / 4104.0, 29440.0 / 4104.0, -845.0 / 4104.0};
const double f_a5[] = {-6080.0 / 20520.0, 41040.0 / 20520.0, -28352.0 / 20520.0,
                       9295.0 / 20520.0, -5643.0 / 20520.0};
const double *const f_a[6] = {nullptr, f_a1, f_a2, f_a3, f_a4, f_a5};
const double f_b[6] = {902880.0 / 7618050.0, 0.0, 3953664.0 / 7618050.0, 3855735.0 / 7618050.0,
                       -1371249.0 / 7618050.0, 277020.0 / 7618050.0};
const double f_e[6] = {1.0 / 360.0, 0.0, -128.0 / 4275.0, -2197.0 / 75240.0, 1.0 / 50.0,
                       2.0 / 55.0};
const Tableau TAB_RKF45 = {6, f_c, f_a, f_b, f_e};

// --- Prince-Dormand 8(7)13M; the 8th-order solution is propagated (GSL rk8pd.c) ---
const double p_c[13] = {0.0,
                        1.0 / 18.0,
                        1.0 / 12.0,
                        1.0 / 8.0,
                        5.0 / 16.0,
                        3.0 / 8.0,
                        59.0 / 400.0,
                        93.0 / 200.0,
                        5490023248.0 / 9719169821.0,
                        13.0 / 20.0,
                        1201146811.0 / 1299019798.0,
                        1.0,
                        1.0};
const double p_a1[] = {1.0 / 18.0};
const double p_a2[] = {1.0 / 48.0, 1.0 / 16.0};
const double p_a3[] = {1.0 / 32.0, 0.0, 3.0 / 32.0};
const double p_a4[] = {5.0 / 16.0, 0.0, -75.0 / 64.0, 75.0 / 64.0};
const double p_a5[] = {3.0 / 80.0, 0.0, 0.0, 3.0 / 16.0, 3.0 / 20.0};
const double p_a6[] = {29443841.0 / 614563906.0, 0.0, 0.0, 77736538.0 / 692538347.0,
                       -28693883.0 / 1125000000.0, 23124283.0 / 1800000000.0};
const double p_a7[] = {16016141.0 / 946692911.0, 0.0, 0.0, 61564180.0 / 158732637.0,
                       22789713.0 / 633445777.0, 545815736.0 / 2771057229.0,
                       -180193667.0 / 1043307555.0};
const double p_a8[] = {39632708.0 / 573591083.0, 0.0, 0.0, -433636366.0 / 683701615.0,
                       -421739975.0 / 2616292301.0, 100302831.0 / 723423059.0,
                       790204164.0 / 839813087.0, 800635310.0 / 3783071287.0};
const double p_a9[] = {246121993.0 / 1340847787.0, 0.0, 0.0, -37695042795.0 / 15268766246.0,
                       -309121744.0 / 1061227803.0, -12992083.0 / 490766935.0,
                       6005943493.0 / 2108947869.0, 393006217.0 / 1396673457.0,
                       123872331.0 / 1001029789.0};
const double p_a10[] = {-1028468189.0 / 846180014.0, 0.0, 0.0, 8478235783.0 / 508512852.0,
                        1311729495.0 / 1432422823.0, -10304129995.0 / 1701304382.0,
                        -48777925059.0 / 3047939560.0, 15336726248.0 / 1032824649.0,
                        -45442868181.0 / 3398467696.0, 3065993473.0 / 597172653.0};
const double p_a11[] = {185892177.0 / 718116043.0, 0.0, 0.0, -3185094517.0 / 667107341.0,
                        -477755414.0 / 1098053517.0, -703635378.0 / 230739211.0,
                        5731566787.0 / 1027545527.0, 5232866602.0 / 850066563.0,
                        -4093664535.0 / 808688257.0, 3962137247.0 / 1805957418.0,
                        65686358.0 / 487910083.0};
const double p_a12[] = {403863854.0 / 491063109.0, 0.0, 0.0, -5068492393.0 / 434740067.0,
                        -411421997.0 / 543043805.0, 652783627.0 / 914296604.0,
                        11173962825.0 / 925320556.0, -13158990841.0 / 6184727034.0,
                        3936647629.0 / 1978049680.0, -160528059.0 / 685178525.0,
                        248638103.0 / 1413531060.0, 0.0};
const double *const p_a[13] = {nullptr, p_a1, p_a2, p_a3, p_a4, p_a5, p_a6,
                               p_a7,    p_a8, p_a9, p_a10, p_a11, p_a12};
const double p_b8[13] = {14005451.0 / 335480064.0, 0.0, 0.0, 0.0, 0.0,
                         -59238493.0 / 1068277825.0, 181606767.0 / 758867731.0,
                         561292985.0 / 797845732.0, -1041891430.0 / 1371343529.0,
                         760417239.0 / 1151165299.0, 118820643.0 / 751138087.0,
                         -528747749.0 / 2220607170.0, 1.0 / 4.0};
const double p_b7[13] = {13451932.0 / 455176623.0, 0.0, 0.0, 0.0, 0.0,
                         -808719846.0 / 976000145.0, 1757004468.0 / 5645159321.0,
                         656045339.0 / 265891186.0, -3867574721.0 / 1518517206.0,
                         465885868.0 / 322736535.0, 53011238.0 / 667516719.0, 2.0 / 45.0, 0.0};
double p_e[13]; // = b7 - b8 (GSL: yerr = h*(ksum7 - ksum8)), filled at load
struct InitPE {
  InitPE() {
    for (int i = 0; i < 13; i++) p_e[i] = p_b7[i] - p_b8[i];
  }
} init_pe;
const Tableau TAB_RK8PD = {13, p_c, p_a, p_b8, p_e};

const gsl_odeiv_step_type TYPE_RKF45 = {"rkf45", 1, 6, 5};
const gsl_odeiv_step_type TYPE_RK8PD = {"rk8pd", 1, 13, 8};

} // namespace

extern "C" {

const gsl_odeiv_step_type *gsl_odeiv_step_rkf45 = &TYPE_RKF45;
const gsl_odeiv_step_type *gsl_odeiv_step_rk8pd = &TYPE_RK8PD;
long shim_ode_attempts[2] = {0, 0}, shim_ode_rejects[2] = {0, 0}, shim_ode_rhs[2] = {0, 0};

gsl_odeiv_step *gsl_odeiv_step_alloc(const gsl_odeiv_step_type *T, size_t dim) {
  gsl_odeiv_step *s = (gsl_odeiv_step *)malloc(sizeof(gsl_odeiv_step));
  s->type = T;
  s->dimension = dim;
  s->work = (double *)malloc(sizeof(double) * dim * (size_t)(T->n_stages + 2));
  return s;
}
void gsl_odeiv_step_free(gsl_odeiv_step *s) {
  if (!s) return;
  free(s->work);
  free(s);
}
gsl_odeiv_control *gsl_odeiv_control_y_new(double eps_abs, double eps_rel) {
  gsl_odeiv_control *c = (gsl_odeiv_control *)malloc(sizeof(gsl_odeiv_control));
  c->eps_abs = eps_abs;
  c->eps_rel = eps_rel;
  c->a_y = 1.0;
  c->a_dydt = 0.0;
  return c;
}
void gsl_odeiv_control_free(gsl_odeiv_control *c) { free(c); }
gsl_odeiv_evolve *gsl_odeiv_evolve_alloc(size_t dim) {
  gsl_odeiv_evolve *e = (gsl_odeiv_evolve *)malloc(sizeof(gsl_odeiv_evolve));
  e->dimension = dim;
  e->y0 = (double *)malloc(sizeof(double) * dim);
  e->yerr = (double *)malloc(sizeof(double) * dim);
  e->dydt_in = (double *)malloc(sizeof(double) * dim);
  e->dydt_out = (double *)malloc(sizeof(double) * dim);
  e->last_step = 0;
  e->count = 0;
  e->failed_steps = 0;
  return e;
}
void gsl_odeiv_evolve_free(gsl_odeiv_evolve *e) {
  if (!e) return;
  free(e->y0);
  free(e->yerr);
  free(e->dydt_in);
  free(e->dydt_out);
  free(e);
}

} // extern "C"

namespace {

// One explicit RK step: y (in/out), yerr (out), dydt_in (k1), dydt_out = f(t+h, y_new).
int step_apply(const gsl_odeiv_step *step, const Tableau &T, double t, double h, double y[],
               double yerr[], const double dydt_in[], double dydt_out[],
               const gsl_odeiv_system *sys, int which) {
  const size_t dim = step->dimension;
  double *k = step->work;                         // stages * dim
  double *ytmp = step->work + (size_t)T.stages * dim;
  double *y0 = ytmp + dim;
  memcpy(y0, y, sizeof(double) * dim);
  memcpy(k, dydt_in, sizeof(double) * dim);
  for (int s = 1; s < T.stages; s++) {
    const double *row = T.a[s];
    for (size_t i = 0; i < dim; i++) {
      double acc = 0.0;
      for (int j = 0; j < s; j++)
        if (row[j] != 0.0) acc += row[j] * k[(size_t)j * dim + i];
      ytmp[i] = y0[i] + h * acc;
    }
    int st = sys->function(t + T.c[s] * h, ytmp, k + (size_t)s * dim, sys->params);
    shim_ode_rhs[which]++;
    if (st != GSL_SUCCESS) return st;
  }
  for (size_t i = 0; i < dim; i++) {
    double acc = 0.0, err = 0.0;
    for (int j = 0; j < T.stages; j++) {
      const double kj = k[(size_t)j * dim + i];
      if (T.b[j] != 0.0) acc += T.b[j] * kj;
      if (T.e[j] != 0.0) err += T.e[j] * kj;
    }
    y[i] = y0[i] + h * acc;
    yerr[i] = h * err;
  }
  int st = sys->function(t + h, y, dydt_out, sys->params);
  shim_ode_rhs[which]++;
  return st;
}

// GSL cstd.c: std_control_hadjust.  returns -1 (DEC), 0 (NIL), +1 (INC)
int control_hadjust(const gsl_odeiv_control *c, size_t dim, unsigned int ord, const double y[],
                    const double yerr[], const double yp[], double *h, double *rmax_out) {
  const double S = 0.9;
  const double h_old = *h;
  double rmax = DBL_MIN;
  for (size_t i = 0; i < dim; i++) {
    const double D0 =
        c->eps_rel * (c->a_y * fabs(y[i]) + c->a_dydt * fabs(h_old * yp[i])) + c->eps_abs;
    const double r = fabs(yerr[i]) / fabs(D0);
    if (r > rmax) rmax = r;
  }
  if (rmax_out) *rmax_out = rmax;
  if (rmax > 1.1) {
    double r = S / pow(rmax, 1.0 / ord);
    if (r < 0.2) r = 0.2;
    *h = r * h_old;
    return -1;
  } else if (rmax < 0.5) {
    double r = S / pow(rmax, 1.0 / (ord + 1.0));
    if (r > 5.0) r = 5.0;
    if (r < 1.0) r = 1.0;
    *h = r * h_old;
    return 1;
  }
  return 0;
}

} // namespace

extern "C" int gsl_odeiv_evolve_apply(gsl_odeiv_evolve *e, gsl_odeiv_control *con,
                                      gsl_odeiv_step *step, const gsl_odeiv_system *dydt,
                                      double *t, double t1, double *h, double y[]) {
  const double t0 = *t;
  double h0 = *h;
  int final_step = 0;
  const double dt = t1 - t0;
  const size_t dim = e->dimension;
  const int which = (step->type == &TYPE_RKF45) ? 0 : 1;
  const Tableau &T = which == 0 ? TAB_RKF45 : TAB_RK8PD;
  static FILE *trace = nullptr;
  static int trace_init = 0;
  if (!trace_init) {
    trace_init = 1;
    const char *p = getenv("SHIM_ODE_TRACE");
    if (p && *p) trace = fopen(p, "w");
  }

  if ((dt < 0.0 && h0 > 0.0) || (dt > 0.0 && h0 < 0.0)) return GSL_EINVAL;

  memcpy(e->y0, y, sizeof(double) * dim);
  {
    int st = dydt->function(t0, y, e->dydt_in, dydt->params);
    shim_ode_rhs[which]++;
    if (st) return st;
  }

  for (;;) {
    if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) {
      h0 = dt;
      final_step = 1;
    } else {
      final_step = 0;
    }
    int st = step_apply(step, T, t0, h0, y, e->yerr, e->dydt_in, e->dydt_out, dydt, which);
    if (st != GSL_SUCCESS) {
      *h = h0;
      *t = t0;
      return st;
    }
    e->count++;
    e->last_step = h0;
    shim_ode_attempts[which]++;
    *t = final_step ? t1 : t0 + h0;

    const double h_old = h0;
    double rmax = 0;
    const int adj =
        control_hadjust(con, dim, step->type->order, y, e->yerr, e->dydt_out, &h0, &rmax);
    if (trace && which == 0)
      fprintf(trace, "t0=%.15g h=%.15g rmax=%.6e adj=%d final=%d\n", t0, h_old, rmax, adj,
              final_step);
    if (adj == -1) {
      volatile double t_curr = *t;
      volatile double t_next = (*t) + h0;
      if (fabs(h0) < fabs(h_old) && t_next != t_curr) {
        memcpy(y, e->y0, sizeof(double) * dim);
        e->failed_steps++;
        shim_ode_rejects[which]++;
        continue; // retry with the smaller h0, re-using dydt_in
      } else {
        h0 = h_old;
      }
    }
    break;
  }
  *h = h0;
  return GSL_SUCCESS;
}

// ---------------------------------------------------------------------------------
// QAG, 61-point Gauss-Kronrod (QUADPACK dqage / GSL qag.c + qk.c + qpsrt.c)
// ---------------------------------------------------------------------------------
extern "C" {
long shim_qag_calls = 0, shim_qag_intervals = 0, shim_qag_fevals = 0;

gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n) {
  gsl_integration_workspace *w =
      (gsl_integration_workspace *)malloc(sizeof(gsl_integration_workspace));
  w->limit = n;
  w->size = 0;
  w->nrmax = 0;
  w->i = 0;
  w->maximum_level = 0;
  w->alist = (double *)malloc(n * sizeof(double));
  w->blist = (double *)malloc(n * sizeof(double));
  w->rlist = (double *)malloc(n * sizeof(double));
  w->elist = (double *)malloc(n * sizeof(double));
  w->order = (size_t *)malloc(n * sizeof(size_t));
  w->level = (size_t *)malloc(n * sizeof(size_t));
  return w;
}
void gsl_integration_workspace_free(gsl_integration_workspace *w) {
  if (!w) return;
  free(w->alist);
  free(w->blist);
  free(w->rlist);
  free(w->elist);
  free(w->order);
  free(w->level);
  free(w);
}
}

namespace {

double rescale_error(double err, const double result_abs, const double result_asc) {
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    double scale = pow((200 * err / result_asc), 1.5);
    err = (scale < 1) ? result_asc * scale : result_asc;
  }
  if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
    double min_err = 50 * DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  return err;
}

void qk61(const gsl_function *f, double a, double b, double *result, double *abserr,
          double *resabs, double *resasc) {
  const int n = 31;
  const double *xgk = gk61_xgk, *wgk = gk61_wgk, *wg = gk61_wg;
  double fv1[31], fv2[31];
  const double center = 0.5 * (a + b);
  const double half_length = 0.5 * (b - a);
  const double abs_half_length = fabs(half_length);
  const double f_center = f->function(center, f->params);
  double result_gauss = 0;
  double result_kronrod = f_center * wgk[n - 1];
  double result_abs = fabs(result_kronrod);
  double result_asc = 0;
  if (n % 2 == 0) result_gauss = f_center * wg[n / 2 - 1];
  for (int j = 0; j < (n - 1) / 2; j++) {
    const int jtw = j * 2 + 1;
    const double abscissa = half_length * xgk[jtw];
    const double fval1 = f->function(center - abscissa, f->params);
    const double fval2 = f->function(center + abscissa, f->params);
    const double fsum = fval1 + fval2;
    fv1[jtw] = fval1;
    fv2[jtw] = fval2;
    result_gauss += wg[j] * fsum;
    result_kronrod += wgk[jtw] * fsum;
    result_abs += wgk[jtw] * (fabs(fval1) + fabs(fval2));
  }
  for (int j = 0; j < n / 2; j++) {
    const int jtwm1 = j * 2;
    const double abscissa = half_length * xgk[jtwm1];
    const double fval1 = f->function(center - abscissa, f->params);
    const double fval2 = f->function(center + abscissa, f->params);
    fv1[jtwm1] = fval1;
    fv2[jtwm1] = fval2;
    result_kronrod += wgk[jtwm1] * (fval1 + fval2);
    result_abs += wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
  }
  shim_qag_fevals += 61;
  const double mean = result_kronrod * 0.5;
  result_asc = wgk[n - 1] * fabs(f_center - mean);
  for (int j = 0; j < n - 1; j++)
    result_asc += wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
  double err = (result_kronrod - result_gauss) * half_length;
  result_kronrod *= half_length;
  result_abs *= abs_half_length;
  result_asc *= abs_half_length;
  *result = result_kronrod;
  *resabs = result_abs;
  *resasc = result_asc;
  *abserr = rescale_error(err, result_abs, result_asc);
}

// QUADPACK dqpsrt as in GSL qpsrt.c: keeps order[] sorted by decreasing error estimate.
void qpsrt(gsl_integration_workspace *w) {
  const size_t last = w->size - 1;
  const size_t limit = w->limit;
  double *elist = w->elist;
  size_t *order = w->order;
  size_t i_nrmax = w->nrmax;
  size_t i_maxerr = order[i_nrmax];

  if (last < 2) {
    order[0] = 0;
    order[1] = 1;
    w->i = i_maxerr;
    return;
  }
  const double errmax = elist[i_maxerr];
  while (i_nrmax > 0 && errmax > elist[order[i_nrmax - 1]]) {
    order[i_nrmax] = order[i_nrmax - 1];
    i_nrmax--;
  }
  int top;
  if (last < (limit / 2 + 2))
    top = (int)last;
  else
    top = (int)(limit - last + 1);
  int i = (int)i_nrmax + 1;
  while (i < top && errmax < elist[order[i]]) {
    order[i - 1] = order[i];
    i++;
  }
  order[i - 1] = i_maxerr;
  const double errmin = elist[last];
  int k = top - 1;
  while (k > i - 2 && errmin >= elist[order[k]]) {
    order[k + 1] = order[k];
    k--;
  }
  order[k + 1] = last;
  i_maxerr = order[i_nrmax];
  w->i = i_maxerr;
  w->nrmax = i_nrmax;
}

int subinterval_too_small(double a1, double a2, double b2) {
  const double e = DBL_EPSILON, u = DBL_MIN;
  double tmp = (1 + 100 * e) * (fabs(a2) + 1000 * u);
  return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

} // namespace

extern "C" int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs,
                                   double epsrel, size_t limit, int key,
                                   gsl_integration_workspace *w, double *result,
                                   double *abserr) {
  if (key != 6) {
    fprintf(stderr, "gsl shim: qag key %d not implemented (reference only uses 6)\n", key);
    abort();
  }
  shim_qag_calls++;
  // initialise
  w->size = 0;
  w->nrmax = 0;
  w->i = 0;
  w->alist[0] = a;
  w->blist[0] = b;
  w->rlist[0] = 0;
  w->elist[0] = 0;
  w->order[0] = 0;
  w->level[0] = 0;
  w->maximum_level = 0;
  *result = 0;
  *abserr = 0;
  if (limit > w->limit) return GSL_EINVAL;
  if (epsabs <= 0 && (epsrel < 50 * DBL_EPSILON || epsrel < 0.5e-28)) return GSL_EINVAL;

  double result0, abserr0, resabs0, resasc0;
  qk61(f, a, b, &result0, &abserr0, &resabs0, &resasc0);
  w->size = 1;
  w->rlist[0] = result0;
  w->elist[0] = abserr0;
  shim_qag_intervals++;

  double tolerance = fmax(epsabs, epsrel * fabs(result0));
  volatile double round_off = 50 * DBL_EPSILON * resabs0;
  if (abserr0 <= round_off && abserr0 > tolerance) {
    *result = result0;
    *abserr = abserr0;
    return GSL_EROUND;
  } else if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) {
    *result = result0;
    *abserr = abserr0;
    return GSL_SUCCESS;
  } else if (limit == 1) {
    *result = result0;
    *abserr = abserr0;
    return GSL_EMAXITER;
  }

  double area = result0, errsum = abserr0;
  size_t iteration = 1;
  int roundoff_type1 = 0, roundoff_type2 = 0, error_type = 0;
  do {
    const size_t imax = w->i;
    const double a_i = w->alist[imax], b_i = w->blist[imax];
    const double r_i = w->rlist[imax], e_i = w->elist[imax];
    const double a1 = a_i, b1 = 0.5 * (a_i + b_i), a2 = b1, b2 = b_i;
    double area1, area2, error1, error2, resabs1, resabs2, resasc1, resasc2;
    qk61(f, a1, b1, &area1, &error1, &resabs1, &resasc1);
    qk61(f, a2, b2, &area2, &error2, &resabs2, &resasc2);
    shim_qag_intervals += 2;
    const double area12 = area1 + area2, error12 = error1 + error2;
    errsum += (error12 - e_i);
    area += area12 - r_i;
    if (resasc1 != error1 && resasc2 != error2) {
      double delta = r_i - area12;
      if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) roundoff_type1++;
      if (iteration >= 10 && error12 > e_i) roundoff_type2++;
    }
    tolerance = fmax(epsabs, epsrel * fabs(area));
    if (errsum > tolerance) {
      if (roundoff_type1 >= 6 || roundoff_type2 >= 20) error_type = 2;
      if (subinterval_too_small(a1, a2, b2)) error_type = 3;
    }
    // update(): the half with the larger error stays in slot imax, the other is appended
    const size_t i_new = w->size;
    const size_t new_level = w->level[imax] + 1;
    if (error2 > error1) {
      w->alist[imax] = a2;
      w->rlist[imax] = area2;
      w->elist[imax] = error2;
      w->level[imax] = new_level;
      w->alist[i_new] = a1;
      w->blist[i_new] = b1;
      w->rlist[i_new] = area1;
      w->elist[i_new] = error1;
      w->level[i_new] = new_level;
    } else {
      w->blist[imax] = b1;
      w->rlist[imax] = area1;
      w->elist[imax] = error1;
      w->level[imax] = new_level;
      w->alist[i_new] = a2;
      w->blist[i_new] = b2;
      w->rlist[i_new] = area2;
      w->elist[i_new] = error2;
      w->level[i_new] = new_level;
    }
    w->size++;
    if (new_level > w->maximum_level) w->maximum_level = new_level;
    qpsrt(w);
    iteration++;
  } while (iteration < limit && !error_type && errsum > tolerance);

  double sum = 0;
  for (size_t k = 0; k < w->size; k++) sum += w->rlist[k];
  *result = sum;
  *abserr = errsum;
  if (errsum <= tolerance) return GSL_SUCCESS;
  if (error_type == 2) return GSL_EROUND;
  if (error_type == 3) return GSL_ESING;
  if (iteration == limit) return GSL_EMAXITER;
  return GSL_EFAILED;
}

// ---------------------------------------------------------------------------------
// radix-2 FFTs (GSL sign/layout conventions; twiddles from sincos, not a recurrence)
// ---------------------------------------------------------------------------------
namespace {

typedef std::complex<double> cplx;

bool is_pow2(size_t n) { return n && !(n & (n - 1)); }

// in-place complex FFT, sign = -1 forward / +1 backward, unnormalised
void fft_c(cplx *x, size_t n, int sign) {
  for (size_t i = 1, j = 0; i < n; i++) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(x[i], x[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const double ang = sign * 2.0 * M_PI / (double)len;
    const size_t half = len >> 1;
    std::vector<cplx> w(half);
    for (size_t k = 0; k < half; k++) w[k] = cplx(cos(ang * (double)k), sin(ang * (double)k));
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < half; k++) {
        cplx u = x[i + k], v = x[i + k + half] * w[k];
        x[i + k] = u + v;
        x[i + k + half] = u - v;
      }
  }
}

} // namespace

extern "C" {

int gsl_fft_complex_radix2_forward(gsl_complex_packed_array data, size_t stride, size_t n) {
  if (stride != 1 || !is_pow2(n)) return GSL_EINVAL;
  fft_c(reinterpret_cast<cplx *>(data), n, -1);
  return GSL_SUCCESS;
}
int gsl_fft_complex_radix2_backward(gsl_complex_packed_array data, size_t stride, size_t n) {
  if (stride != 1 || !is_pow2(n)) return GSL_EINVAL;
  fft_c(reinterpret_cast<cplx *>(data), n, +1);
  return GSL_SUCCESS;
}
int gsl_fft_complex_radix2_inverse(gsl_complex_packed_array data, size_t stride, size_t n) {
  int st = gsl_fft_complex_radix2_backward(data, stride, n);
  if (st) return st;
  const double norm = 1.0 / (double)n;
  for (size_t i = 0; i < 2 * n; i++) data[i] *= norm;
  return GSL_SUCCESS;
}

// real -> halfcomplex: x[k] = Re z_k (k=0..n/2), x[n-k] = Im z_k (k=1..n/2-1)
int gsl_fft_real_radix2_transform(double data[], size_t stride, size_t n) {
  if (stride != 1 || !is_pow2(n)) return GSL_EINVAL;
  std::vector<cplx> z(n);
  for (size_t i = 0; i < n; i++) z[i] = cplx(data[i], 0.0);
  fft_c(z.data(), n, -1);
  data[0] = z[0].real();
  if (n > 1) data[n / 2] = z[n / 2].real();
  for (size_t k = 1; k < n / 2; k++) {
    data[k] = z[k].real();
    data[n - k] = z[k].imag();
  }
  return GSL_SUCCESS;
}

int gsl_fft_halfcomplex_radix2_backward(double data[], size_t stride, size_t n) {
  if (stride != 1 || !is_pow2(n)) return GSL_EINVAL;
  std::vector<cplx> z(n);
  z[0] = cplx(data[0], 0.0);
  if (n > 1) z[n / 2] = cplx(data[n / 2], 0.0);
  for (size_t k = 1; k < n / 2; k++) {
    z[k] = cplx(data[k], data[n - k]);
    z[n - k] = cplx(data[k], -data[n - k]);
  }
  fft_c(z.data(), n, +1);
  for (size_t i = 0; i < n; i++) data[i] = z[i].real();
  return GSL_SUCCESS;
}
int gsl_fft_halfcomplex_radix2_inverse(double data[], size_t stride, size_t n) {
  int st = gsl_fft_halfcomplex_radix2_backward(data, stride, n);
  if (st) return st;
  const double norm = 1.0 / (double)n;
  for (size_t i = 0; i < n; i++) data[i] *= norm;
  return GSL_SUCCESS;
}

} // extern "C"

// ---------------------------------------------------------------------------------
// complex log-Gamma: Lanczos (g = 7, 9 coefficients) with reflection for Re z <= 1/2
// (GSL specfunc/gamma.c lngamma_lanczos_complex + gsl_sf_complex_logsin_e)
// ---------------------------------------------------------------------------------
namespace {

const double lanczos_7_c[9] = {0.99999999999980993227684700473478,
                               676.520368121885098567009190444019,
                               -1259.13921672240287047156078755283,
                               771.3234287776530788486528258894,
                               -176.61502916214059906584551354,
                               12.507343278686904814458936853,
                               -0.13857109526572011689554707,
                               9.984369578019570859563e-6,
                               1.50563273514931155834e-7};
const double LogRootTwoPi = 0.9189385332046727418;
const double LnPi = 1.14472988584940017414342735135;

double angle_restrict_symm(double theta) {
  // to (-pi, pi]
  const double P1 = 4 * 7.8539812564849853515625e-01;
  const double P2 = 4 * 3.7748947079307981766760e-08;
  const double P3 = 4 * 2.6951514290790594840552e-15;
  const double TwoPi = 2 * (P1 + P2 + P3);
  const double y = (theta >= 0 ? 1.0 : -1.0) * 2 * floor(fabs(theta) / TwoPi);
  double r = ((theta - y * P1) - y * P2) - y * P3;
  if (r > M_PI)
    r = (((r - 2 * P1) - 2 * P2) - 2 * P3);
  else if (r < -M_PI)
    r = (((r + 2 * P1) + 2 * P2) + 2 * P3);
  return r;
}

void complex_log(double zr, double zi, double *lnr, double *theta) {
  const double ax = fabs(zr), ay = fabs(zi);
  const double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
  *lnr = log(mx) + 0.5 * log(1.0 + (mn / mx) * (mn / mx));
  *theta = atan2(zi, zr);
}

void lngamma_lanczos_complex(double zr, double zi, double *yr, double *yi) {
  zr -= 1.0;
  double Ag_r = lanczos_7_c[0], Ag_i = 0.0;
  for (int k = 1; k <= 8; k++) {
    const double R = zr + k, I = zi;
    const double a = lanczos_7_c[k] / (R * R + I * I);
    Ag_r += a * R;
    Ag_i -= a * I;
  }
  double log1_r, log1_i, logAg_r, logAg_i;
  complex_log(zr + 7.5, zi, &log1_r, &log1_i);
  complex_log(Ag_r, Ag_i, &logAg_r, &logAg_i);
  *yr = (zr + 0.5) * log1_r - zi * log1_i - (zr + 7.5) + LogRootTwoPi + logAg_r;
  *yi = angle_restrict_symm(zi * log1_r + (zr + 0.5) * log1_i - zi + logAg_i);
}

void complex_logsin(double zr, double zi, double *lszr, double *lszi) {
  if (zi > 60.0) {
    *lszr = -M_LN2 + zi;
    *lszi = 0.5 * M_PI - zr;
  } else if (zi < -60.0) {
    *lszr = -M_LN2 - zi;
    *lszi = -0.5 * M_PI + zr;
  } else {
    const double sr = sin(zr) * cosh(zi), si = cos(zr) * sinh(zi);
    complex_log(sr, si, lszr, lszi);
  }
  *lszi = angle_restrict_symm(*lszi);
}

} // namespace

extern "C" int gsl_sf_lngamma_complex_e(double zr, double zi, gsl_sf_result *lnr,
                                        gsl_sf_result *arg) {
  if (zr <= 0.5) {
    double a, b, lr, li;
    lngamma_lanczos_complex(1.0 - zr, -zi, &a, &b);
    complex_logsin(M_PI * zr, M_PI * zi, &lr, &li);
    lnr->val = LnPi - lr - a;
    arg->val = angle_restrict_symm(-li - b);
  } else {
    lngamma_lanczos_complex(zr, zi, &lnr->val, &arg->val);
  }
  lnr->err = 0;
  arg->err = 0;
  return GSL_SUCCESS;
}
