// Scalar building blocks of the Time-RG hot path, written once as __host__ __device__
// inline functions.  The CUDA kernels (kernels_*.cu) are thin parallel drivers around
// these; tests/harness/math_harness.cc compiles the same header with g++ to check the arithmetic against
// the oracle on machines without a GPU (test infrastructure only -- the shipped library
// has no CPU execution path).
//
// Reference citations: "rt:" = src/redTime.cc, "hdr:" = src/AU_cosmological_parameters.h,
// "tab:" = src/AU_tabfun.h, "itp:" = src/AU_interp.h.
#pragma once
#include <cfloat>
#include <cmath>

#ifdef __CUDACC__
#define RT_HD __host__ __device__ __forceinline__
#define RT_UNROLL _Pragma("unroll")
#else
#define RT_HD inline
#define RT_UNROLL
#endif

namespace rtrg {

// number of state components per wavenumber (rt:150): 3 ln P, 14 I, 24 Q
enum { N_UP = 3, N_UI = 14, N_UQ = 24, N_U = 41 };

// ------------------------------------------------------------------------------------
// per-cosmology scalars (device resident, one struct per cosmology)
// ------------------------------------------------------------------------------------
struct Cosmo {
  // params_redTime.dat (hdr:325-333)
  double ns, s8, h, Om, Ob, On, TK, w0, wa;
  // derived as in the constructor (hdr:343-349)
  double Og, fnu, fcb, On_hot, anu, Or, OL;
  double z_in, a_in;
  int sw_nl, sw_1l, sw_pl, sw_pr;
  int n_out;
  // tables (offsets into the pooled device arrays, in doubles)
  // tables: offsets (in doubles) into the single device-resident input pool.  The transfer
  // columns are uploaded raw and turned into log tables in place by k_prep_inputs.
  int nT;             // rows of the z=0 transfer table
  long long offT;     // k_T       -> ln k            (hdr:812-823)
  long long offLT;    // Tc_T      -> ln(T_cb/T_cb[0])
  long long offTb;    // Tb_T      (raw, consumed by the transform)
  int n_z, n_kb;      // beta table a-nodes x k-nodes
  long long offA;     // a nodes = 1/(1+z_interp)
  long long offKb;    // k nodes of the interpolation files
  long long offB;     // beta = f_nu T_nu / T_c [n_z][n_kb]  (hdr:556-623): formed while staging, or
                      // (page-locked caller buffers, sent as they are) T_nu here until
  long long offTc;    // k_prep_inputs divides by the raw T_c stored at offTc; -1 when staged
  // reduce_beta: instead of the full beta table the host sends what the run consumes:
  long long offRow1;  // beta(a = 1, k_b[*]) [n_kb] (sigma_8 / sigma_v integrands); -1: full table
  long long offBred;  // beta pre-reduced in k at the nk grid and n_lnk+1 growth wavenumbers [n_z][nkk]
  // results of the device-side initialisation
  double Norm;        // sigma_8 normalisation (hdr:874)
  double sigv2_0;     // sigma_v^2(z=0)       (hdr:961)
  int status;
  int pad_;
};

static const double C_RHO_GAM = 4.46911743913795e-07;  // hdr:64
static const double C_NU_HOT = 0.681321952980717;      // hdr:65
static const double H0H = 0.00033356754857714242474;   // rt:69, H0 / (h/Mpc)

// fill the derived members (hdr:343-349)
RT_HD void cosmo_derive(Cosmo &c) {
  c.Og = C_RHO_GAM * (c.TK * c.TK * c.TK * c.TK) / (c.h * c.h);
  c.fnu = c.On / c.Om;
  c.fcb = 1.0 - c.fnu;
  c.On_hot = C_NU_HOT * c.Og;
  c.anu = C_NU_HOT * c.Og / (c.fnu * c.Om + 1e-15);
  c.Or = c.Og + c.On_hot * (c.anu > 1.0);
  c.OL = 1.0 - c.Om - c.Or;
  c.a_in = 1.0 / (1.0 + c.z_in);
}

// ---- background: member versions (hdr:395-497), used by the RHS and the outputs -------
RT_HD double bg_E(const Cosmo &c, double a) {
  return pow(a, -3.0 * (1.0 + c.w0 + c.wa)) * exp(-3.0 * c.wa * (1.0 - a));
}
RT_HD double bg_dEda(const Cosmo &c, double a) {
  return 3.0 * bg_E(c, a) * (c.wa - (1.0 + c.w0 + c.wa) / a);
}
RT_HD double bg_Y(const Cosmo &c, double a) {
  if (a >= c.anu) return c.fnu / c.fcb;
  return C_NU_HOT * c.Og / (c.fcb * c.Om * a);
}
RT_HD double bg_dYda(const Cosmo &c, double a) {
  if (a >= c.anu) return 0;
  return -C_NU_HOT * c.Og / (c.fcb * c.Om * a * a);
}
RT_HD double bg_H2(const Cosmo &c, double a) {
  return c.fcb * c.Om * (1.0 + bg_Y(c, a)) / pow(a, 3.0) + c.OL * bg_E(c, a) + c.Og / pow(a, 4.0);
}
RT_HD double bg_dlnH(const Cosmo &c, double a) {
  return 0.5 * a / bg_H2(c, a) *
         (c.fcb * c.Om * (-3.0 * (1.0 + bg_Y(c, a)) + a * bg_dYda(c, a)) / pow(a, 4.0) +
          c.OL * bg_dEda(c, a) - 4.0 * c.Og / pow(a, 5.0));
}

// ---- background: the static p[]-based twins used inside the growth ODE (hdr:438-493) --
struct BgStatic {
  double Og, fn, fc, anu, Om, On, OL, w0, wa;
};
RT_HD BgStatic bg_static(const Cosmo &c) {
  BgStatic s;
  const double t = c.TK * c.TK / c.h;
  s.Og = C_RHO_GAM * (t * t);  // pow(p[7]*p[7]/p[2], 2)
  s.fn = c.On / c.Om;
  s.fc = 1.0 - s.fn;
  s.anu = C_NU_HOT * s.Og / (c.On + 1e-15);
  s.Om = c.Om;
  s.On = c.On;
  s.OL = c.OL;
  s.w0 = c.w0;
  s.wa = c.wa;
  return s;
}
RT_HD double bgs_E(const BgStatic &s, double a) {
  return pow(a, -3.0 * (1.0 + s.w0 + s.wa)) * exp(-3.0 * s.wa * (1.0 - a));
}
RT_HD double bgs_Y(const BgStatic &s, double a) {
  if (a >= s.anu) return s.fn / s.fc;
  return C_NU_HOT * s.Og / (s.fc * s.Om * a);
}
RT_HD double bgs_dYda(const BgStatic &s, double a) {
  if (a >= s.anu) return 0;
  return -C_NU_HOT * s.Og / ((s.Om - s.On) * a * a);
}
RT_HD double bgs_H2(const BgStatic &s, double a) {
  return (s.Om - s.On) * (1.0 + bgs_Y(s, a)) / pow(a, 3.0) + s.OL * bgs_E(s, a) +
         s.Og / pow(a, 4.0);
}
RT_HD double bgs_dlnH(const BgStatic &s, double a) {
  const double E = bgs_E(s, a), dEda = 3.0 * E * (s.wa - (1.0 + s.w0 + s.wa) / a);
  return 0.5 * a / bgs_H2(s, a) *
         (s.fc * s.Om * (-3.0 * (1.0 + bgs_Y(s, a)) + a * bgs_dYda(s, a)) / pow(a, 4.0) +
          s.OL * dEda - 4.0 * s.Og / pow(a, 5.0));
}

// ------------------------------------------------------------------------------------
// tabulated_function semantics (tab:250-328, 437-501) with O(log n) interval search that
// returns the same interval index as the reference's linear scans.
// ------------------------------------------------------------------------------------
// n = 0; while (x[n+1] < xq && n < size-2) n++;   (tab:473-501)
RT_HD int tab_find(const double *x, int size, double xq) {
  int lb = 1, ub = size - 1;
  while (lb < ub) {
    const int mid = (lb + ub) >> 1;
    if (x[mid] < xq) lb = mid + 1; else ub = mid;
  }
  return lb - 1;
}
RT_HD double lin2(double x0, double x1, double f0, double f1, double xq) {
  return f0 + (f1 - f0) / (x1 - x0) * (xq - x0);  // tab:437-441
}
// Lagrange cubic through 4 points, same operation order as tab:444-471 / itp:38-65
RT_HD double cub4(const double *x, double f0, double f1, double f2, double f3, double xq) {
  return (xq - x[1]) * (xq - x[2]) * (xq - x[3]) / (x[0] - x[1]) / (x[0] - x[2]) / (x[0] - x[3]) * f0 +
         (xq - x[0]) * (xq - x[2]) * (xq - x[3]) / (x[1] - x[0]) / (x[1] - x[2]) / (x[1] - x[3]) * f1 +
         (xq - x[0]) * (xq - x[1]) * (xq - x[3]) / (x[2] - x[0]) / (x[2] - x[1]) / (x[2] - x[3]) * f2 +
         (xq - x[0]) * (xq - x[1]) * (xq - x[2]) / (x[3] - x[0]) / (x[3] - x[1]) / (x[3] - x[2]) * f3;
}
// interpolation weights on 4 consecutive nodes n0..n0+3: f(xq) = sum_m w[m] f[n0+m]
struct Stencil {
  int n0;
  double w[4];
};
RT_HD double stencil_apply(const Stencil &s, const double *f, int stride = 1) {
  double r = 0;
  for (int m = 0; m < 4; m++)
    if (s.w[m] != 0.0) r += s.w[m] * f[(long long)(s.n0 + m) * stride];
  return r;
}
// 1-D table f(x) (tab:250-260)
RT_HD double tab1d(const double *x, const double *f, int size, double xq) {
  const int n1 = tab_find(x, size, xq);
  if (n1 <= 0) return lin2(x[0], x[1], f[0], f[1], xq);
  if (n1 >= size - 2) return lin2(x[size - 2], x[size - 1], f[size - 2], f[size - 1], xq);
  return cub4(x + n1 - 1, f[n1 - 1], f[n1], f[n1 + 1], f[n1 + 2], xq);
}
// 2-D table f(x,y), storage f[ny + Y*nx] (tab:262-328,435): interpolate in x on the four
// y-rows ny-1..ny+2, then in y.  Rows outside the table are never used by the result
// (the reference reads them out of bounds); they are skipped here.
RT_HD double tab2d(const double *xs, int X, const double *ys, int Y, const double *f, double xq,
                   double yq) {
  const int nx = tab_find(xs, X, xq), ny = tab_find(ys, Y, yq);
  const bool xcub = (nx > 0 && nx < X - 2);
  const bool ycub = (ny > 0 && ny < Y - 2);
  double fy[4] = {0, 0, 0, 0};
  for (int r = 0; r < 4; r++) {
    const int iy = ny - 1 + r;
    if (!ycub && (r == 0 || r == 3)) continue;
    if (iy < 0 || iy >= Y) continue;
    if (xcub)
      fy[r] = cub4(xs + nx - 1, f[iy + (long long)Y * (nx - 1)], f[iy + (long long)Y * nx],
                   f[iy + (long long)Y * (nx + 1)], f[iy + (long long)Y * (nx + 2)], xq);
    else
      fy[r] = lin2(xs[nx], xs[nx + 1], f[iy + (long long)Y * nx], f[iy + (long long)Y * (nx + 1)], xq);
  }
  if (ycub) return cub4(ys + ny - 1, fy[0], fy[1], fy[2], fy[3], yq);
  return lin2(ys[ny], ys[ny + 1], fy[1], fy[2], yq);
}
// x-direction rule of the 2-D table applied to one pre-reduced row g[0..X)
RT_HD double tab_row_x(const double *xs, int X, const double *g, double xq, long long gs = 1) {
  const int nx = tab_find(xs, X, xq);
  if (nx > 0 && nx < X - 2)
    return cub4(xs + nx - 1, g[(nx - 1) * gs], g[nx * gs], g[(nx + 1) * gs], g[(nx + 2) * gs], xq);
  return lin2(xs[nx], xs[nx + 1], g[nx * gs], g[(nx + 1) * gs], xq);
}
// The same rule split in two: the part that depends on the abscissa only (shared by every row
// that is looked up at the same xq) and its application to one row.  The weights are cub4's own
// quotient chains, so apply() returns what tab_row_x returns.
struct RowX {
  int nx, cub;
  double w[4];   // cubic: weights on nodes nx-1..nx+2
  double x0, x1, xq;  // linear: the two nodes
};
// nx = tab_find(xs, X, xq), found by the caller
RT_HD RowX tab_row_x_prepare_at(const double *xs, int X, double xq, int nx) {
  RowX r;
  r.nx = nx;
  r.cub = (r.nx > 0 && r.nx < X - 2);
  r.xq = xq;
  r.x0 = xs[r.nx];
  r.x1 = xs[r.nx + 1];
  r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0.0;
  if (r.cub) {
    const double *x = xs + r.nx - 1;
    r.w[0] = (xq - x[1]) * (xq - x[2]) * (xq - x[3]) / (x[0] - x[1]) / (x[0] - x[2]) / (x[0] - x[3]);
    r.w[1] = (xq - x[0]) * (xq - x[2]) * (xq - x[3]) / (x[1] - x[0]) / (x[1] - x[2]) / (x[1] - x[3]);
    r.w[2] = (xq - x[0]) * (xq - x[1]) * (xq - x[3]) / (x[2] - x[0]) / (x[2] - x[1]) / (x[2] - x[3]);
    r.w[3] = (xq - x[0]) * (xq - x[1]) * (xq - x[2]) / (x[3] - x[0]) / (x[3] - x[1]) / (x[3] - x[2]);
  }
  return r;
}
RT_HD RowX tab_row_x_prepare(const double *xs, int X, double xq) {
  return tab_row_x_prepare_at(xs, X, xq, tab_find(xs, X, xq));
}
RT_HD double tab_row_x_apply(const RowX &r, const double *g, long long gs = 1) {
  if (r.cub)
    return r.w[0] * g[(r.nx - 1) * gs] + r.w[1] * g[r.nx * gs] + r.w[2] * g[(r.nx + 1) * gs] +
           r.w[3] * g[(r.nx + 2) * gs];
  return lin2(r.x0, r.x1, g[r.nx * gs], g[(r.nx + 1) * gs], r.xq);
}
// y-direction weights of the 2-D table (rows ny-1..ny+2; linear at the two edge intervals)
RT_HD Stencil tab_stencil_y(const double *ys, int Y, double yq) {
  Stencil s;
  const int ny = tab_find(ys, Y, yq);
  s.w[0] = s.w[1] = s.w[2] = s.w[3] = 0;
  if (ny > 0 && ny < Y - 2) {
    const double *p = ys + ny - 1;
    s.n0 = ny - 1;
    s.w[0] = (yq - p[1]) * (yq - p[2]) * (yq - p[3]) / (p[0] - p[1]) / (p[0] - p[2]) / (p[0] - p[3]);
    s.w[1] = (yq - p[0]) * (yq - p[2]) * (yq - p[3]) / (p[1] - p[0]) / (p[1] - p[2]) / (p[1] - p[3]);
    s.w[2] = (yq - p[0]) * (yq - p[1]) * (yq - p[3]) / (p[2] - p[0]) / (p[2] - p[1]) / (p[2] - p[3]);
    s.w[3] = (yq - p[0]) * (yq - p[1]) * (yq - p[2]) / (p[3] - p[0]) / (p[3] - p[1]) / (p[3] - p[2]);
  } else {
    s.n0 = ny;  // weights on rows ny, ny+1 stored in w[0], w[1]
    const double t = (yq - ys[ny]) / (ys[ny + 1] - ys[ny]);
    s.w[0] = 1.0 - t;
    s.w[1] = t;
  }
  return s;
}

// ------------------------------------------------------------------------------------
// Beta_P(a,k) = f_nu T_nu/T_c (hdr:513-637): table in (a, linear k), clamped
// ------------------------------------------------------------------------------------
struct BetaTab {
  int n_z, n_kb;
  const double *a;     // [n_z] ascending
  const double *k;     // [n_kb]
  const double *beta;  // [n_z][n_kb]  (k fastest, as tab: f[ny + Y*nx]); nullptr when reduced
  const double *row1;  // reduced upload: the table interpolated in a at a = 1, [n_kb]
  double fn, kmin, kmax;
};
// returns NaN for a > 1.001 (the reference aborts, hdr:528-531)
RT_HD double beta_P(const BetaTab &t, double a, double k) {
  if (t.n_z == 0) return 0;
  if (t.fn < 1e-10) return 0;
  if (a > 1.001) return NAN;
  if (a > 1.0) a = 1.0;
  if (k < t.kmin) k = t.kmin;
  if (k > t.kmax) k = t.kmax;
  if (!t.beta) {
    // reduced upload: only a = 1 is available in general k (tab2d interpolates in a first, column
    // by column, so the pre-interpolated row followed by the k rule gives the same value)
    if (a != 1.0) return NAN;
    const int ny = tab_find(t.k, t.n_kb, k);
    if (ny > 0 && ny < t.n_kb - 2) return cub4(t.k + ny - 1, t.row1[ny - 1], t.row1[ny], t.row1[ny + 1], t.row1[ny + 2], k);
    return lin2(t.k[ny], t.k[ny + 1], t.row1[ny], t.row1[ny + 1], k);
  }
  return tab2d(t.a, t.n_z, t.k, t.n_kb, t.beta, a, k);
}
// the same look-up through a row pre-reduced in k: brow[j] = sum_r wy_r beta[j][ny-1+r]
RT_HD double beta_row(const BetaTab &t, const double *brow, double a, long long bs = 1) {
  if (t.n_z == 0) return 0;
  if (t.fn < 1e-10) return 0;
  if (a > 1.001) return NAN;
  if (a > 1.0) a = 1.0;
  return tab_row_x(t.a, t.n_z, brow, a, bs);
}

// ------------------------------------------------------------------------------------
// explicit embedded Runge-Kutta tableaux (GSL rkf45.c / rk8pd.c)
// ------------------------------------------------------------------------------------
struct RKF45 {
  // nodes, rows, 5th-order weights (propagated), error weights.  Written as switches so that
  // device code holds no local constant arrays.
  static RT_HD double c(int s) {
    switch (s) {
      case 1: return 1.0 / 4.0;
      case 2: return 3.0 / 8.0;
      case 3: return 12.0 / 13.0;
      case 4: return 1.0;
      case 5: return 1.0 / 2.0;
      default: return 0.0;
    }
  }
  static RT_HD double a(int s, int j) {
    switch (8 * s + j) {
      case 8 * 1 + 0: return 1.0 / 4.0;
      case 8 * 2 + 0: return 3.0 / 32.0;
      case 8 * 2 + 1: return 9.0 / 32.0;
      case 8 * 3 + 0: return 1932.0 / 2197.0;
      case 8 * 3 + 1: return -7200.0 / 2197.0;
      case 8 * 3 + 2: return 7296.0 / 2197.0;
      case 8 * 4 + 0: return 8341.0 / 4104.0;
      case 8 * 4 + 1: return -32832.0 / 4104.0;
      case 8 * 4 + 2: return 29440.0 / 4104.0;
      case 8 * 4 + 3: return -845.0 / 4104.0;
      case 8 * 5 + 0: return -6080.0 / 20520.0;
      case 8 * 5 + 1: return 41040.0 / 20520.0;
      case 8 * 5 + 2: return -28352.0 / 20520.0;
      case 8 * 5 + 3: return 9295.0 / 20520.0;
      case 8 * 5 + 4: return -5643.0 / 20520.0;
      default: return 0.0;
    }
  }
  static RT_HD double b(int j) {
    switch (j) {
      case 0: return 902880.0 / 7618050.0;
      case 2: return 3953664.0 / 7618050.0;
      case 3: return 3855735.0 / 7618050.0;
      case 4: return -1371249.0 / 7618050.0;
      case 5: return 277020.0 / 7618050.0;
      default: return 0.0;
    }
  }
  static RT_HD double e(int j) {
    switch (j) {
      case 0: return 1.0 / 360.0;
      case 2: return -128.0 / 4275.0;
      case 3: return -2197.0 / 75240.0;
      case 4: return 1.0 / 50.0;
      case 5: return 2.0 / 55.0;
      default: return 0.0;
    }
  }
};

// GSL std_control_hadjust with a_y = 1, a_dydt = 0 (control_y_new): decision from rmax.
// returns -1 (decrease), 0 (keep), +1 (increase); *h is updated.
RT_HD int gsl_hadjust(double rmax, int ord, double *h) {
  const double S = 0.9, h_old = *h;
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

// Prince-Dormand 8(7) tableau
#define RT_PD_ROWS                                                                              \
  {0},                                                                                          \
  {1.0 / 18.0},                                                                                 \
  {1.0 / 48.0, 1.0 / 16.0},                                                                     \
  {1.0 / 32.0, 0.0, 3.0 / 32.0},                                                                \
  {5.0 / 16.0, 0.0, -75.0 / 64.0, 75.0 / 64.0},                                                 \
  {3.0 / 80.0, 0.0, 0.0, 3.0 / 16.0, 3.0 / 20.0},                                               \
  {29443841.0 / 614563906.0, 0.0, 0.0, 77736538.0 / 692538347.0, -28693883.0 / 1125000000.0,    \
   23124283.0 / 1800000000.0},                                                                  \
  {16016141.0 / 946692911.0, 0.0, 0.0, 61564180.0 / 158732637.0, 22789713.0 / 633445777.0,      \
   545815736.0 / 2771057229.0, -180193667.0 / 1043307555.0},                                    \
  {39632708.0 / 573591083.0, 0.0, 0.0, -433636366.0 / 683701615.0,                              \
   -421739975.0 / 2616292301.0, 100302831.0 / 723423059.0, 790204164.0 / 839813087.0,           \
   800635310.0 / 3783071287.0},                                                                 \
  {246121993.0 / 1340847787.0, 0.0, 0.0, -37695042795.0 / 15268766246.0,                        \
   -309121744.0 / 1061227803.0, -12992083.0 / 490766935.0, 6005943493.0 / 2108947869.0,         \
   393006217.0 / 1396673457.0, 123872331.0 / 1001029789.0},                                     \
  {-1028468189.0 / 846180014.0, 0.0, 0.0, 8478235783.0 / 508512852.0,                           \
   1311729495.0 / 1432422823.0, -10304129995.0 / 1701304382.0, -48777925059.0 / 3047939560.0,   \
   15336726248.0 / 1032824649.0, -45442868181.0 / 3398467696.0, 3065993473.0 / 597172653.0},    \
  {185892177.0 / 718116043.0, 0.0, 0.0, -3185094517.0 / 667107341.0,                            \
   -477755414.0 / 1098053517.0, -703635378.0 / 230739211.0, 5731566787.0 / 1027545527.0,        \
   5232866602.0 / 850066563.0, -4093664535.0 / 808688257.0, 3962137247.0 / 1805957418.0,        \
   65686358.0 / 487910083.0},                                                                   \
  {403863854.0 / 491063109.0, 0.0, 0.0, -5068492393.0 / 434740067.0,                            \
   -411421997.0 / 543043805.0, 652783627.0 / 914296604.0, 11173962825.0 / 925320556.0,          \
   -13158990841.0 / 6184727034.0, 3936647629.0 / 1978049680.0, -160528059.0 / 685178525.0,      \
   248638103.0 / 1413531060.0, 0.0}
#define RT_PD_C                                                                                 \
  0.0, 1.0 / 18.0, 1.0 / 12.0, 1.0 / 8.0, 5.0 / 16.0, 3.0 / 8.0, 59.0 / 400.0, 93.0 / 200.0,    \
      5490023248.0 / 9719169821.0, 13.0 / 20.0, 1201146811.0 / 1299019798.0, 1.0, 1.0
#define RT_PD_B8                                                                                \
  14005451.0 / 335480064.0, 0.0, 0.0, 0.0, 0.0, -59238493.0 / 1068277825.0,                     \
      181606767.0 / 758867731.0, 561292985.0 / 797845732.0, -1041891430.0 / 1371343529.0,       \
      760417239.0 / 1151165299.0, 118820643.0 / 751138087.0, -528747749.0 / 2220607170.0,       \
      1.0 / 4.0
#define RT_PD_B7                                                                                \
  13451932.0 / 455176623.0, 0.0, 0.0, 0.0, 0.0, -808719846.0 / 976000145.0,                     \
      1757004468.0 / 5645159321.0, 656045339.0 / 265891186.0, -3867574721.0 / 1518517206.0,     \
      465885868.0 / 322736535.0, 53011238.0 / 667516719.0, 2.0 / 45.0, 0.0

// ------------------------------------------------------------------------------------
// growth factor ODE (hdr:133-190): y = {D, dD/da}, integrated with RK8PD + GSL control
// (eps_abs = 0, eps_rel = 1e-6), restarting every leg with dt = 1e-6 a_begin.
// ------------------------------------------------------------------------------------
struct GrowthCtx {
  BgStatic bg;
  BetaTab bt;
  const double *brow;  // beta row pre-reduced at this wavenumber (clamped k)
  long long bstride;   // stride between successive a-nodes of brow
  RT_HD void operator()(double a, const double y[2], double f[2]) const;
};
// One evaluation costs one pow + one exp (the dark-energy factor E, shared by H^2 and
// dlnH/dlna) and a handful of multiplications: the integer powers of a are products, not
// pow() calls -- this function runs ~12k times per table entry (13 stages x ~940 attempts).
RT_HD void growth_rhs(const GrowthCtx &g, double a, const double y[2], double f[2]) {
  const BgStatic &s = g.bg;
  const double a2 = a * a, a3 = a2 * a, a4 = a2 * a2, a5 = a4 * a;
  const double E = bgs_E(s, a), dEda = 3.0 * E * (s.wa - (1.0 + s.w0 + s.wa) / a);
  const double Y = bgs_Y(s, a), dYda = bgs_dYda(s, a);
  const double H2 = (s.Om - s.On) * (1.0 + Y) / a3 + s.OL * E + s.Og / a4;  // hdr:473-476
  const double dlnH = 0.5 * a / H2 *
                      (s.fc * s.Om * (-3.0 * (1.0 + Y) + a * dYda) / a4 + s.OL * dEda - 4.0 * s.Og / a5);
  const double F0 = 1.5 * s.Om / (a5 * H2);
  const double F1 = (3.0 + dlnH) / a;
  const double beta = (a < 1e-3) ? s.fn : beta_row(g.bt, g.brow, fmin(a, 1.0), g.bstride);
  f[0] = y[1];
  f[1] = -F1 * y[1] + F0 * (s.fc + beta) * y[0];  // F_MG = 0 (hdr:151-153)
}

RT_HD void GrowthCtx::operator()(double a, const double y[2], double f[2]) const { growth_rhs(*this, a, y, f); }

struct PDTableau {
  double A[13][12], C[13], B8[13], B7[13];
};
inline PDTableau make_pd_tableau() {
  const double A[13][12] = {RT_PD_ROWS};
  const double C[13] = {RT_PD_C};
  const double B8[13] = {RT_PD_B8};
  const double B7[13] = {RT_PD_B7};
  PDTableau T;
  for (int i = 0; i < 13; i++) {
    for (int j = 0; j < 12; j++) T.A[i][j] = A[i][j];
    T.C[i] = C[i];
    T.B8[i] = B8[i];
    T.B7[i] = B7[i];
  }
  return T;
}

// one leg a_begin -> a_end (hdr:170-190); returns number of attempted steps
template <class Rhs>
RT_HD int growth_integrate(const PDTableau &PD, const Rhs &g, double a_begin, double a_end,
                           double y[2]) {
  const double(*A)[12] = PD.A;
  const double *C = PD.C, *B8 = PD.B8, *B7 = PD.B7;
  double t = a_begin;
  const double t1 = a_end;
  double h = 1e-6 * t;
  int attempts = 0;
  while ((t1 - t) * h > 0) {
    // --- gsl_odeiv_evolve_apply ---
    const double t0 = t, dt = t1 - t0;
    const double y0[2] = {y[0], y[1]};
    double k[13][2];
    g(t0, y0, k[0]);
    double h0 = h;
    for (;;) {
      bool final_step = false;
      if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) {
        h0 = dt;
        final_step = true;
      }
      for (int s = 1; s < 13; s++) {
        double acc0 = 0, acc1 = 0;
        for (int j = 0; j < s; j++) {
          const double a_sj = A[s][j];
          if (a_sj != 0.0) {
            acc0 += a_sj * k[j][0];
            acc1 += a_sj * k[j][1];
          }
        }
        const double yt[2] = {y0[0] + h0 * acc0, y0[1] + h0 * acc1};
        g(t0 + C[s] * h0, yt, k[s]);
      }
      double s8[2] = {0, 0}, s7[2] = {0, 0};
      for (int j = 0; j < 13; j++) {
        if (B8[j] != 0.0) {
          s8[0] += B8[j] * k[j][0];
          s8[1] += B8[j] * k[j][1];
        }
        if (B7[j] != 0.0) {
          s7[0] += B7[j] * k[j][0];
          s7[1] += B7[j] * k[j][1];
        }
      }
      const double yn[2] = {y0[0] + h0 * s8[0], y0[1] + h0 * s8[1]};
      const double ye[2] = {h0 * (s7[0] - s8[0]), h0 * (s7[1] - s8[1])};
      attempts++;
      const double tn = final_step ? t1 : t0 + h0;
      // control_y_new(0, 1e-6), order 8
      double rmax = DBL_MIN;
      for (int i = 0; i < 2; i++) {
        const double D0 = 1e-6 * fabs(yn[i]) + 0.0;
        const double r = fabs(ye[i]) / fabs(D0);
        if (r > rmax) rmax = r;
      }
      const double h_old = h0;
      const int adj = gsl_hadjust(rmax, 8, &h0);
      if (adj == -1) {
        const double t_next = tn + h0;
        if (fabs(h0) < fabs(h_old) && t_next != tn) continue;  // reject, retry smaller
        h0 = h_old;
      }
      y[0] = yn[0];
      y[1] = yn[1];
      t = tn;
      h = h0;
      break;
    }
    if (attempts > 2000000) break;
  }
  return attempts;
}

// ------------------------------------------------------------------------------------
// growth tables (hdr:639-738): G(ln a, ln k) = D/a and dD/da on (n_lna+1) x (n_lnk+1)
// ------------------------------------------------------------------------------------
struct GrowthTab {
  int n_lna, n_lnk;      // number of intervals; tables have +1 nodes
  const double *lna;     // [n_lna+1]
  const double *lnk;     // [n_lnk+1]
  const double *G;       // [(n_lna+1)][(n_lnk+1)]  (k fastest)
  const double *dD;      // same layout
  const double *Dnorm;   // [n_lnk+1] = G(ln a = 0, lnk_j)
};
static const double GROWTH_A_MIN = 1e-3, GROWTH_A_MAX = 1.1;     // hdr:644
static const double GROWTH_K_MIN = 1.5e-4, GROWTH_K_MAX = 9.0;   // hdr:651
// D and dD/da at (z,k); returns false when z is outside the table (reference aborts)
RT_HD bool growth_D_dD(const GrowthTab &t, double z, double k, double *D, double *dDda) {
  const double a = 1.0 / (z + 1.0);
  if (a > GROWTH_A_MAX || a < GROWTH_A_MIN) return false;
  if (k > GROWTH_K_MAX) k = GROWTH_K_MAX;
  if (k < GROWTH_K_MIN) k = GROWTH_K_MIN;
  const double lna0 = log(a), lnk0 = log(k);
  const double D0 = tab1d(t.lnk, t.Dnorm, t.n_lnk + 1, lnk0);
  *D = tab2d(t.lna, t.n_lna + 1, t.lnk, t.n_lnk + 1, t.G, lna0, lnk0) * a / D0;
  *dDda = tab2d(t.lna, t.n_lna + 1, t.lnk, t.n_lnk + 1, t.dD, lna0, lnk0) / D0;
  return true;
}
// the same through rows pre-reduced at a fixed wavenumber: Grow[i], dDrow[i], i<=n_lna
RT_HD bool growth_D_dD_row(const double *lna, int n_lna, const double *Grow, const double *dDrow,
                           long long rs, double D0, double z, double *D, double *dDda) {
  const double a = 1.0 / (z + 1.0);
  if (a > GROWTH_A_MAX || a < GROWTH_A_MIN) return false;
  const double lna0 = log(a);
  *D = tab_row_x(lna, n_lna + 1, Grow, lna0, rs) * a / D0;
  *dDda = tab_row_x(lna, n_lna + 1, dDrow, lna0, rs) / D0;
  return true;
}

// ------------------------------------------------------------------------------------
// linear power spectrum (hdr:790-930)
// ------------------------------------------------------------------------------------
struct LinCtx {
  const Cosmo *c;
  BetaTab bt;
  GrowthTab gt;
  const double *lnkT, *lnT;  // log transfer table (hdr:812-823)
  int nT;
};
RT_HD double transfer_cb(const LinCtx &L, double k) { return exp(tab1d(L.lnkT, L.lnT, L.nT, log(k))); }

// integrand of the sigma_8 normalisation (hdr:204-217)
RT_HD double sigma8_integrand(const LinCtx &L, double lnkR) {
  const double R = 8.0;
  const double kR = exp(lnkR), kR2 = kR * kR, kR3 = kR2 * kR, k = kR / R;
  const double T = transfer_cb(L, k);
  const double F = 1.0 - L.c->On / L.c->Om + beta_P(L.bt, 1.0, k);
  double W = 1.0 - 0.1 * kR * kR;
  if (kR > 1e-2) W = 3.0 * (sin(kR) / kR3 - cos(kR) / kR2);
  return W * W * T * T * F * F * pow(k, L.c->ns + 3.0) / (2.0 * M_PI * M_PI);
}
// Plin(z,k) (hdr:881-890); Norm must be set
RT_HD double plin(const LinCtx &L, double z, double k) {
  const double T = transfer_cb(L, k);
  const double F = 1.0 - L.c->On / L.c->Om + beta_P(L.bt, 1.0 / (1.0 + z), k);
  double D = NAN, dD = NAN;
  growth_D_dD(L.gt, z, k, &D, &dD);
  return L.c->Norm * pow(k, L.c->ns) * T * T * F * F * D * D;
}
RT_HD double plin_cb(const LinCtx &L, double z, double k) {
  const double fn = L.c->On / L.c->Om, fc = 1.0 - fn;
  if (fn <= 1e-10) return plin(L, z, k);
  const double a = 1.0 / (1.0 + z), Rr = 1.0 / (fc + beta_P(L.bt, a, k));
  return plin(L, z, k) * Rr * Rr;
}
RT_HD double plin_nu(const LinCtx &L, double z, double k) {
  const double fn = L.c->On / L.c->Om, fc = 1.0 - fn;
  if (fn <= 1e-10) return 0;
  const double a = 1.0 / (1.0 + z), B = beta_P(L.bt, a, k), F = fc + B, Rr = B / fn / F;
  return plin(L, z, k) * Rr * Rr;
}
// integrand of sigma_v^2(z=0) (hdr:219-223)
RT_HD double sigmav_integrand(const LinCtx &L, double lnk) { return exp(lnk) * plin(L, 0.0, exp(lnk)); }

// ------------------------------------------------------------------------------------
// QAG with the 61-point Gauss-Kronrod rule (QUADPACK dqage as in GSL).
// The 61 integrand values of one interval are produced by the caller (in parallel on the
// device); the combination below is sequential and in QUADPACK's summation order.
// ------------------------------------------------------------------------------------
struct Qk61Out {
  double result, abserr, resabs, resasc;
};
// abscissa of sample s in [0,61): s=0 centre, s=1+2j / 2+2j = centre -/+ half*xgk[j]
RT_HD double qk61_abscissa(const double *xgk, double a, double b, int s) {
  const double center = 0.5 * (a + b), half_length = 0.5 * (b - a);
  if (s == 0) return center;
  const int j = (s - 1) >> 1;
  const double absc = half_length * xgk[j];
  return ((s - 1) & 1) ? center + absc : center - absc;
}
RT_HD Qk61Out qk61_combine(const double *xgk, const double *wgk, const double *wg, double a,
                           double b, const double *fv /*[61] in qk61_abscissa order*/) {
  (void)xgk;
  const int n = 31;
  const double half_length = 0.5 * (b - a), abs_half_length = fabs(half_length);
  const double f_center = fv[0];
  double result_gauss = 0, result_kronrod = f_center * wgk[n - 1];
  double result_abs = fabs(result_kronrod), result_asc = 0;
  for (int j = 0; j < (n - 1) / 2; j++) {
    const int jtw = j * 2 + 1;
    const double fval1 = fv[1 + 2 * jtw], fval2 = fv[2 + 2 * jtw], fsum = fval1 + fval2;
    result_gauss += wg[j] * fsum;
    result_kronrod += wgk[jtw] * fsum;
    result_abs += wgk[jtw] * (fabs(fval1) + fabs(fval2));
  }
  for (int j = 0; j < n / 2; j++) {
    const int jtwm1 = j * 2;
    const double fval1 = fv[1 + 2 * jtwm1], fval2 = fv[2 + 2 * jtwm1];
    result_kronrod += wgk[jtwm1] * (fval1 + fval2);
    result_abs += wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
  }
  const double mean = result_kronrod * 0.5;
  result_asc = wgk[n - 1] * fabs(f_center - mean);
  for (int j = 0; j < n - 1; j++)
    result_asc += wgk[j] * (fabs(fv[1 + 2 * j] - mean) + fabs(fv[2 + 2 * j] - mean));
  double err = (result_kronrod - result_gauss) * half_length;
  result_kronrod *= half_length;
  result_abs *= abs_half_length;
  result_asc *= abs_half_length;
  // rescale_error
  err = fabs(err);
  if (result_asc != 0 && err != 0) {
    const double scale = pow((200 * err / result_asc), 1.5);
    err = (scale < 1) ? result_asc * scale : result_asc;
  }
  if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
    const double min_err = 50 * DBL_EPSILON * result_abs;
    if (min_err > err) err = min_err;
  }
  Qk61Out o = {result_kronrod, err, result_abs, result_asc};
  return o;
}

// Adaptive driver state; intervals kept in QUADPACK order.  QAG_CAP bounds the interval
// count (the reference passes limit = 1000; its integrands need ~15 intervals).
enum { QAG_CAP = 192 };
struct QagState {
  double alist[QAG_CAP], blist[QAG_CAP], rlist[QAG_CAP], elist[QAG_CAP];
  short order[QAG_CAP];
  int size, nrmax, imax;
  double area, errsum, tolerance;
  int iteration, roundoff1, roundoff2, error_type, done;
  double epsabs, epsrel;
};
RT_HD void qag_qpsrt(QagState &w, int limit) {
  const int last = w.size - 1;
  int i_nrmax = w.nrmax;
  int i_maxerr = w.order[i_nrmax];
  if (last < 2) {
    w.order[0] = 0;
    w.order[1] = 1;
    w.imax = i_maxerr;
    return;
  }
  const double errmax = w.elist[i_maxerr];
  while (i_nrmax > 0 && errmax > w.elist[w.order[i_nrmax - 1]]) {
    w.order[i_nrmax] = w.order[i_nrmax - 1];
    i_nrmax--;
  }
  const int top = (last < (limit / 2 + 2)) ? last : limit - last + 1;
  int i = i_nrmax + 1;
  while (i < top && errmax < w.elist[w.order[i]]) {
    w.order[i - 1] = w.order[i];
    i++;
  }
  w.order[i - 1] = (short)i_maxerr;
  const double errmin = w.elist[last];
  int k = top - 1;
  while (k > i - 2 && errmin >= w.elist[w.order[k]]) {
    w.order[k + 1] = w.order[k];
    k--;
  }
  w.order[k + 1] = (short)last;
  w.imax = w.order[i_nrmax];
  w.nrmax = i_nrmax;
}
// start: result of the whole interval.  returns true when finished.
RT_HD bool qag_begin(QagState &w, double a, double b, double epsabs, double epsrel,
                     const Qk61Out &q0) {
  w.size = 1;
  w.nrmax = 0;
  w.imax = 0;
  w.alist[0] = a;
  w.blist[0] = b;
  w.rlist[0] = q0.result;
  w.elist[0] = q0.abserr;
  w.order[0] = 0;
  w.epsabs = epsabs;
  w.epsrel = epsrel;
  w.area = q0.result;
  w.errsum = q0.abserr;
  w.iteration = 1;
  w.roundoff1 = w.roundoff2 = w.error_type = 0;
  w.tolerance = fmax(epsabs, epsrel * fabs(q0.result));
  const double round_off = 50 * DBL_EPSILON * q0.resabs;
  w.done = 0;
  if (q0.abserr <= round_off && q0.abserr > w.tolerance) w.done = 2;  // GSL_EROUND
  else if ((q0.abserr <= w.tolerance && q0.abserr != q0.resasc) || q0.abserr == 0.0) w.done = 1;
  return w.done != 0;
}
// the interval to bisect next
RT_HD void qag_next(const QagState &w, double *a1, double *b1, double *a2, double *b2) {
  const double a_i = w.alist[w.imax], b_i = w.blist[w.imax];
  *a1 = a_i;
  *b1 = 0.5 * (a_i + b_i);
  *a2 = *b1;
  *b2 = b_i;
}
// feed the two halves; returns true when finished
RT_HD bool qag_update(QagState &w, const Qk61Out &q1, const Qk61Out &q2, int limit) {
  const int imax = w.imax;
  const double a_i = w.alist[imax], b_i = w.blist[imax], r_i = w.rlist[imax], e_i = w.elist[imax];
  const double a1 = a_i, b1 = 0.5 * (a_i + b_i), a2 = b1, b2 = b_i;
  const double area12 = q1.result + q2.result, error12 = q1.abserr + q2.abserr;
  w.errsum += (error12 - e_i);
  w.area += area12 - r_i;
  if (q1.resasc != q1.abserr && q2.resasc != q2.abserr) {
    const double delta = r_i - area12;
    if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) w.roundoff1++;
    if (w.iteration >= 10 && error12 > e_i) w.roundoff2++;
  }
  w.tolerance = fmax(w.epsabs, w.epsrel * fabs(w.area));
  if (w.errsum > w.tolerance) {
    if (w.roundoff1 >= 6 || w.roundoff2 >= 20) w.error_type = 2;
    const double tmp = (1 + 100 * DBL_EPSILON) * (fabs(a2) + 1000 * DBL_MIN);
    if (fabs(a1) <= tmp && fabs(b2) <= tmp) w.error_type = 3;
  }
  const int i_new = w.size;
  if (q2.abserr > q1.abserr) {
    w.alist[imax] = a2;
    w.rlist[imax] = q2.result;
    w.elist[imax] = q2.abserr;
    w.alist[i_new] = a1;
    w.blist[i_new] = b1;
    w.rlist[i_new] = q1.result;
    w.elist[i_new] = q1.abserr;
  } else {
    w.blist[imax] = b1;
    w.rlist[imax] = q1.result;
    w.elist[imax] = q1.abserr;
    w.alist[i_new] = a2;
    w.blist[i_new] = b2;
    w.rlist[i_new] = q2.result;
    w.elist[i_new] = q2.abserr;
  }
  w.size++;
  qag_qpsrt(w, limit);
  w.iteration++;
  if (!(w.iteration < limit && !w.error_type && w.errsum > w.tolerance)) w.done = 1;
  if (w.size >= QAG_CAP - 1 && !w.done) w.done = 3;  // cap reached
  return w.done != 0;
}
RT_HD double qag_result(const QagState &w) {
  double s = 0;
  for (int k = 0; k < w.size; k++) s += w.rlist[k];
  return s;
}

// ------------------------------------------------------------------------------------
// Time-RG right-hand side for one wavenumber (rt:1383-1547)
// ------------------------------------------------------------------------------------
// position of the 64-slot index 32a+16c+8d+4b+2e+f inside the 14 unique components
// (rt:147-157, 236-259): -1 = identically zero
RT_HD int i64_slot(int J) {
  switch (J) {
    case 8: case 16: return 0;
    case 9: case 18: return 1;
    case 10: case 17: return 2;
    case 11: case 19: return 3;
    case 12: case 20: return 4;
    case 13: case 22: return 5;
    case 14: case 21: return 6;
    case 15: case 23: return 7;
    case 56: return 8;
    case 57: case 58: return 9;
    case 59: return 10;
    case 60: return 11;
    case 61: case 62: return 12;
    case 63: return 13;
    default: return -1;
  }
}
RT_HD int nAI(int a, int c, int d, int b, int e, int f) { return 32 * a + 16 * c + 8 * d + 4 * b + 2 * e + f; }
// the unique components (rt:151-157): a = c = (j >= 8), d = 1, and (b,e,f) the binary digits of
// j for j < 8, of {0,1,3,4,5,7}[j-8] otherwise.  Pure arithmetic: no local tables in device code.
RT_HD void unique_abcdef(int j, int *a, int *c, int *d, int *b, int *e, int *f) {
  const int hi = (j >= 8);
  const int bef = hi ? ((0x754310 >> (4 * (j - 8))) & 7) : j;
  *a = hi; *c = hi; *d = 1; *b = (bef >> 2) & 1; *e = (bef >> 1) & 1; *f = bef & 1;
}

// Omega(i,j) (rt:1383-1411) without a local table
#define RT_OM(i, j) ((i) == 0 ? ((j) == 0 ? 1.0 : -1.0) : ((j) == 0 ? Om10 : Om11))
// ln P_ab (3) and I_acd,bef (14): y17 = y[0..17), A14 = unique A_acd,bef, dy17 out (rt:1449-1513)
RT_HD void trg_rhs_PI(double eeta, double k, double Om10, double Om11, int nonlinear, const double *y,
                      const double *A14, double *dy) {
  const double P[3] = {exp(y[0]), exp(y[1]), exp(y[2])};
  double dP[3] = {0, 0, 0};
  RT_UNROLL
  for (int c = 0; c < 2; c++) {
    dP[0] -= RT_OM(0, c) * P[c] + RT_OM(0, c) * P[c];
    dP[1] -= RT_OM(0, c) * P[c + 1] + RT_OM(1, c) * P[c];
    dP[2] -= RT_OM(1, c) * P[c + 1] + RT_OM(1, c) * P[c + 1];
    if (nonlinear) {
      RT_UNROLL
      for (int d = 0; d < 2; d++) {
        RT_UNROLL
        for (int q = 0; q < 3; q++) {
          const int a = (q > 0), b = (q > 1);  // (a,b) = 00, 10, 11
          const int s0 = i64_slot(nAI(a, c, d, b, c, d)), s1 = i64_slot(nAI(b, c, d, a, c, d));
          const double I0 = s0 >= 0 ? y[N_UP + s0] : 0.0, I1 = s1 >= 0 ? y[N_UP + s1] : 0.0;
          dP[q] += eeta * 4.0 * M_PI / k * (I0 + I1);
        }
      }
    }
  }
  dy[0] = dP[0] / P[0];
  dy[1] = dP[1] / P[1];
  dy[2] = dP[2] / P[2];
  if (dy[2] < -10.0) dy[2] = -10.0;  // rt:1488-1491
  if (dy[2] > 10.0) dy[2] = 10.0;
  RT_UNROLL
  for (int j = N_UP; j < N_UP + N_UI; j++) dy[j] = 0;
  if (!nonlinear) return;
  RT_UNROLL
  for (int j = 0; j < N_UI; j++) {
    int a, c, d, b, e, f;
    unique_abcdef(j, &a, &c, &d, &b, &e, &f);
    double v = 2.0 * eeta * A14[j];
    RT_UNROLL
    for (int g = 0; g < 2; g++) {
      const int s1 = i64_slot(nAI(a, c, d, g, e, f)), s2 = i64_slot(nAI(a, c, d, b, g, f)),
                s3 = i64_slot(nAI(a, c, d, b, e, g));
      const double I1 = s1 >= 0 ? y[N_UP + s1] : 0.0, I2 = s2 >= 0 ? y[N_UP + s2] : 0.0,
                   I3 = s3 >= 0 ? y[N_UP + s3] : 0.0;
      v += -RT_OM(b, g) * I1 - RT_OM(e, g) * I2 - RT_OM(f, g) * I3;
    }
    dy[N_UP + j] = v;
  }
}
// one multipole of Q^l_abc: Q, R, dQ of length 8 (rt:1516-1539)
RT_HD void trg_rhs_Q(double eeta, double Om10, double Om11, const double *Q, const double *R, double *dQ) {
  RT_UNROLL
  for (int a = 0; a < 2; a++)
    RT_UNROLL
    for (int b = 0; b < 2; b++)
      RT_UNROLL
      for (int c = 0; c < 2; c++) {
        const int j = 4 * a + 2 * b + c;
        double v = 2.0 * eeta * R[j];
        RT_UNROLL
        for (int d = 0; d < 2; d++)
          v += -RT_OM(a, d) * Q[4 * d + 2 * b + c] - RT_OM(b, d) * Q[4 * a + 2 * d + c] -
               RT_OM(c, d) * Q[4 * a + 2 * b + d];
        dQ[j] = v;
      }
}
// y[41] -> dy[41] for one k.  A14: the 14 unique A_{acd,bef}; R24: R^ell_{abc}.
// Om10 = Omega(1,0), Om11 = Omega(1,1) (rt:1395-1401).  evolve_Q: rt:1516.
RT_HD void trg_rhs_row(double eta, double k, double Om10, double Om11, int nonlinear, int evolve_Q,
                       const double *y, const double *A14, const double *R24, double *dy) {
  const double eeta = exp(eta);
  trg_rhs_PI(eeta, k, Om10, Om11, nonlinear, y, A14, dy);
  for (int j = N_UP + N_UI; j < N_U; j++) dy[j] = 0;
  if (nonlinear && evolve_Q)
    for (int l = 0; l < 3; l++)
      trg_rhs_Q(eeta, Om10, Om11, y + N_UP + N_UI + 8 * l, R24 + 8 * l, dy + N_UP + N_UI + 8 * l);
}

#undef RT_OM

// Omega(1,0) and Omega(1,1) (rt:1395-1401)
RT_HD void trg_omega(const Cosmo &c, double A, double beta, double *Om10, double *Om11) {
  *Om10 = -1.5 * c.Om * (c.fcb + beta) / (A * A * A * bg_H2(c, A));
  *Om11 = 3.0 + bg_dlnH(c, A);
}

// 1-loop rescaling of the z1l cache (rt:1316-1337): exponents of f for the unique A and R
RT_HD int a14_fpow(int j) {
  int a, c, d, b, e, f;
  unique_abcdef(j, &a, &c, &d, &b, &e, &f);
  return b + e + f + 1;
}
RT_HD int r24_fpow(int j) {
  const int abc = j % 8;
  return abc / 4 + (abc % 4) / 2 + abc % 2 + 1;
}
RT_HD double ipow(double x, int n) {
  double r = 1;
  for (int i = 0; i < n; i++) r *= x;
  return r;
}

// P_{B,j} combinations of Q (rt:269-298); Q = y + 17 (24 values), returns without pi*k
RT_HD double pbis_comb(const double *Q, int j_mu, int m_b) {
#define RT_QQ(l, a, b, c) Q[(l)*8 + 4 * (a) + 2 * (b) + (c)]
  double q = 0;
  if (j_mu == 2) {
    if (m_b == 2) q = -2.0 * RT_QQ(0, 0, 1, 0) + (4.0 / 3.0) * RT_QQ(1, 0, 1, 0);
    if (m_b == 1) q = (4.0 / 3.0) * RT_QQ(1, 0, 1, 1) + (6.0 / 5.0) * RT_QQ(2, 0, 1, 1);
  } else if (j_mu == 4) {
    if (m_b == 1)
      q = -2.0 * RT_QQ(0, 1, 1, 0) + (4.0 / 3.0) * RT_QQ(1, 1, 1, 0) - 2.0 * RT_QQ(0, 0, 1, 1) -
          2.0 * RT_QQ(2, 0, 1, 1);
    if (m_b == 0) q = (4.0 / 3.0) * RT_QQ(1, 1, 1, 1) + (6.0 / 5.0) * RT_QQ(2, 1, 1, 1);
  } else if (j_mu == 6) {
    if (m_b == 0) q = -2.0 * RT_QQ(0, 1, 1, 1) - 2.0 * RT_QQ(2, 1, 1, 1);
  }
#undef RT_QQ
  return q;
}

}  // namespace rtrg
