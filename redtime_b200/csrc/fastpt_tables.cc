// Host-side builder of the cosmology-independent quadrature weights (see fastpt_tables.h).
// Everything here runs once per (nk, kmin, kmax) grid at rtrg_create() time.
#include "fastpt_tables.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>

namespace rtrg {

typedef long double ld;
typedef std::complex<long double> cld;

static const ld PI_L = 3.14159265358979323846264338327950288L;
static const ld LN2_L = 0.693147180559945309417232121458176568L;
static const int NU_INT = -2;  // FAST-PT bias exponent nu (redTime.cc:71)

// ------------------------------------------------------------------------------------
// grid + windows (redTime.cc:90-138).  The floating-point expression order of the
// reference is kept so that the window edges fall on the same side of the same samples.
// ------------------------------------------------------------------------------------
GridSpec make_grid(int nk, double kmin, double kmax) {
  GridSpec g;
  g.nk = nk;
  g.np = 4 * nk;
  g.nshift = (g.np - nk) / 2;
  g.kmin = kmin;
  g.kmax = kmax;
  const double lnkmin = std::log(kmin), lnkmax = std::log(kmax);
  g.dlnk = (lnkmax - lnkmin) / (nk - 1);
  g.lnk_pad_min = lnkmin - g.dlnk * g.nshift;
  g.nloMR = g.nshift - nk / 2;
  g.jlo = g.np;
  for (int i = 0; i < g.np; i++)
    if (window_P(g, i) > 0) { g.jlo = i; break; }
  g.nsup = g.np - g.jlo;
  return g;
}

static inline double w_edge(double x) { return x - std::sin(2.0 * M_PI * x) / (2.0 * M_PI); }

double window_P(const GridSpec &g, int ipad) {
  // split of the padded interval in units of nk/16 (redTime.cc:102-103, "for np = 8*nk"
  // but compiled with np = 4*nk: the right taper lies beyond the array)
  const int s_padL = 7 + 16, s_tapL = 1 + 8, s_extL = 16 + 8, s_extR = 16 + 8, s_tapR = 1 + 8;
  const int nk = g.nk;
  const double dlnk = g.dlnk;
  const double Lo = g.lnk_pad_min + dlnk * nk * s_padL / 16;
  const double Li = Lo + dlnk * nk * s_tapL / 16;
  const double Ri = Li + dlnk * (nk * (16 + s_extL + s_extR) / 16 - 1);
  const double Ro = Ri + dlnk * nk * s_tapR / 16;
  const double lnk = g.lnk_pad_min + dlnk * ipad;
  if (lnk <= Lo) return 0;
  if (lnk < Li) return w_edge((lnk - Lo) / (Li - Lo));
  if (lnk < Ri) return 1;
  if (lnk < Ro) return w_edge((Ro - lnk) / (Ro - Ri));
  return 0;
}

double window_C(const GridSpec &g, int n) {
  const int np = g.np, nl = np / 8, nc = np / 2, nr = 7 * np / 8, Dn = 3 * np / 8;
  if (n <= nl || n >= nr) return 1;
  if (n < nc) return w_edge(double(nc - n) / Dn);
  if (n < nr) return w_edge(double(n - nc) / Dn);
  return 1;
}

KernSpec kern_spec(int n) {
  // J: (alpha_n, -alpha_n, ell_n), redTime.cc:731-732,786;  Jn0: redTime.cc:734-736,808
  static const int ell_J[7] = {0, 0, 1, 2, 2, 3, 4}, alpha_J[7] = {0, 2, 1, 0, 2, 1, 0};
  static const int ell_0[7] = {0, 2, 4, 0, 2, 4, 6}, alpha_0[7] = {0, 0, 0, 2, 2, 2, 2};
  KernSpec s;
  if (n < 7) {
    s.alpha = alpha_J[n];
    s.beta = -alpha_J[n];
    s.ell = ell_J[n];
  } else {
    s.alpha = alpha_0[n - 7];
    s.beta = 2;
    s.ell = ell_0[n - 7];
  }
  s.reg = (s.ell == 0 && s.alpha == 2 && s.beta == -2);  // redTime.cc:518
  return s;
}

// ------------------------------------------------------------------------------------
// complex log-Gamma in long double: upward recurrence to Re z >= 24, then Stirling
// ------------------------------------------------------------------------------------
static cld lgamma_c(cld z) {
  static const ld B[] = {1.0L / 12.0L,        -1.0L / 360.0L,       1.0L / 1260.0L,
                         -1.0L / 1680.0L,     1.0L / 1188.0L,       -691.0L / 360360.0L,
                         1.0L / 156.0L,       -3617.0L / 122400.0L, 43867.0L / 244188.0L,
                         -174611.0L / 125400.0L, 77683.0L / 5796.0L};
  cld shift(0, 0);
  while (z.real() < 24.0L) {
    shift += std::log(z);
    z += 1.0L;
  }
  cld zi = 1.0L / z, zi2 = zi * zi, term = zi, s(0, 0);
  for (int k = 0; k < 11; k++) {
    s += B[k] * term;
    term *= zi2;
  }
  return (z - 0.5L) * std::log(z) - z + 0.5L * std::log(2.0L * PI_L) + s - shift;
}

// Gamma((mu+kappa+1)/2) / Gamma((mu-kappa+1)/2), kappa = reK + i imK  (redTime.cc:306-319)
static cld gamma_ratio(ld mu, ld reK, ld imK) {
  cld top(0.5L * (mu + reK + 1.0L), 0.5L * imK), bot(0.5L * (mu - reK + 1.0L), -0.5L * imK);
  if (std::abs(bot) == 0.0L) return cld(0, 0);  // pole of the denominator
  return std::exp(lgamma_c(top) - lgamma_c(bot));
}

// f(rho) of McEwen et al., including the 2^{i Im rho} phase (redTime.cc:321-328)
static cld f_rho(ld reRho, ld imRho) {
  const ld pre = 0.5L * std::sqrt(PI_L) * std::pow(2.0L, reRho);
  cld g = gamma_ratio(0.5L, reRho - 0.5L, imRho);
  return pre * g * std::exp(cld(0, imRho * LN2_L));
}

struct Freq {
  int np;
  ld dlnk;
  ld tau(int m) const { return 2.0L * PI_L * (ld)m / (dlnk * (ld)np); }
};

// g_ell(alpha, m) (redTime.cc:344-355), m may be any non-negative frequency < np/2
static cld g_side(const Freq &F, int ell, int alpha, int m) {
  if (m == 0 && alpha == ell - NU_INT) return cld(0, 0);
  if (alpha == -2 && ell == 0) return f_rho((ld)NU_INT, F.tau(m));  // g_reg (redTime.cc:338)
  return gamma_ratio(0.5L + ell, 1.5L + NU_INT + alpha, F.tau(m));
}

// f(alpha+beta, h) (redTime.cc:331-336), h >= 0
static cld f_side(const Freq &F, int alpha, int beta, int h) {
  return f_rho(-4.0L - 2.0L * NU_INT - (ld)(alpha + beta), -F.tau(h));
}

static void fft_ld(cld *x, int n, int sign) {
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(x[i], x[j]);
  }
  std::vector<cld> w(n / 2);
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1;
    for (int k = 0; k < half; k++) {
      ld ang = sign * 2.0L * PI_L * (ld)k / (ld)len;
      w[k] = cld(std::cos(ang), std::sin(ang));
    }
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < half; k++) {
        cld u = x[i + k], v = x[i + k + half] * w[k];
        x[i + k] = u + v;
        x[i + k + half] = u - v;
      }
  }
}

// ------------------------------------------------------------------------------------
// T_n: the FAST-PT transform of redTime.cc:411-597 written as a bilinear form.
//   x_{2i} = sum_{m,n} F(m+n) GA_m GB_n  c^a_m c^b_n  e^{2 pi i (m+n) i/np},
//   c_m = sum_j a_j e^{-2 pi i j m/np}
// => T[u][v] = pre * sum_{m,n} F(m+n) GA_m GB_n e^{2 pi i (m u + n v)/np},  |m|,|n| < np/2,
// with Hermitian continuation to negative frequencies (the reference stores halfcomplex
// arrays) and these quirks kept: the m=0 coefficient of the non-regularised transform
// uses |g(0)| (redTime.cc:547-548), the h=0 factor is Re f(0) (:568, :481), the Nyquist
// coefficient is dropped (:550, :445).
// ------------------------------------------------------------------------------------
void build_T(const GridSpec &g, int n, std::vector<double> &T, std::vector<double> &kfac) {
  const KernSpec s = kern_spec(n);
  const int np = g.np, nh = np / 2;
  Freq F = {np, (ld)g.dlnk};

  std::vector<cld> GA(nh), GB(nh), FH(np);
  for (int m = 0; m < nh; m++) {
    cld ga = g_side(F, s.ell, s.alpha, m), gb = g_side(F, s.ell, s.beta, m);
    if (s.reg) {
      // redTime.cc:437-438,448-449: extra 2^{1.5+nu+alpha} 2^{i tau} on the alpha side;
      // m = 0 keeps the complex value (cos/sin of the argument, :439-442)
      ga *= std::pow(2.0L, 1.5L + NU_INT + s.alpha) * std::exp(cld(0, F.tau(m) * LN2_L));
    } else if (m == 0) {
      ga = cld(std::abs(ga), 0);
      gb = cld(std::abs(gb), 0);
    }
    const ld w = (ld)window_C(g, m);
    GA[m] = w * ga;
    GB[m] = w * gb;
  }
  for (int h = 0; h < np; h++) {
    cld f = f_side(F, s.alpha, s.beta, h);
    if (!s.reg) f *= std::exp(cld(0, LN2_L * F.tau(h)));  // redTime.cc:579-580
    if (h == 0) f = cld(f.real(), 0);
    FH[h] = f;
  }
  auto at = [&](const std::vector<cld> &G, int m) -> cld {  // m in (-nh, nh)
    return m >= 0 ? G[m] : std::conj(G[-m]);
  };
  auto fh = [&](int h) -> cld { return h >= 0 ? FH[h] : std::conj(FH[-h]); };

  std::vector<cld> K((size_t)np * np, cld(0, 0));
  for (int m = -(nh - 1); m <= nh - 1; m++) {
    const cld ga = at(GA, m);
    const int mi = (m + np) % np;
    for (int q = -(nh - 1); q <= nh - 1; q++) {
      const int qi = (q + np) % np;
      K[(size_t)mi * np + qi] = fh(m + q) * ga * at(GB, q);
    }
  }
  // 2-D transform with kernel e^{+2 pi i (m u + q v)/np}
  std::vector<cld> col(np);
  for (int m = 0; m < np; m++) fft_ld(&K[(size_t)m * np], np, +1);
  for (int v = 0; v < np; v++) {
    for (int m = 0; m < np; m++) col[m] = K[(size_t)m * np + v];
    fft_ld(col.data(), np, +1);
    for (int u = 0; u < np; u++) K[(size_t)u * np + v] = col[u];
  }
  const ld sl = (s.ell % 2 == 0) ? 1.0L : -1.0L;
  ld pre = sl / (2.0L * PI_L * PI_L * (ld)np * (ld)np);
  if (s.reg) pre *= std::sqrt(2.0L / PI_L);  // redTime.cc:503
  T.resize((size_t)np * np);
  for (size_t i = 0; i < (size_t)np * np; i++) T[i] = (double)(pre * K[i].real());

  // k-dependent prefactor: (2k)^{3+2nu+alpha+beta} (redTime.cc:592) or k^{...} (:506)
  kfac.resize(np);
  const double p = 3.0 + 2.0 * NU_INT + s.alpha + s.beta;
  for (int i = 0; i < np; i++) {
    const double ki = std::exp(g.lnk_pad_min + g.dlnk * i);
    kfac[i] = s.reg ? std::pow(ki, p) : std::pow(ki * 2, p);
  }
}

// ------------------------------------------------------------------------------------
// Z_n kernels of the P13-type terms (redTime.cc:599-687): closed forms, a 10-term
// expansion for r < 0.01 or r > 100, and the r == 1 limits.
// ------------------------------------------------------------------------------------
int zkern_index(int n) {
  static const int Zn[7] = {0, 1, -1, 3, -3, 5, -5};  // redTime.cc:738
  return Zn[n];
}

double Zreg(int n, double r) {
  if (n < 0) return Zreg(-n, 1.0 / r);
  const int NT = 10;
  const double eps = 1e-2;
  const bool small = r < eps, large = r > 1.0 / eps;
  const double L = std::log(std::fabs((1.0 + r) / (1.0 - r)));
  double Z = 0;
  switch (n) {
    case 0:
      return 1.0;
    case 1:
      if (small) {
        for (int m = 0; m < NT; m++) Z += 2.0 * std::pow(r, 2.0 * m + 1.0) * (1.0 - r) / (2.0 * m + 1.0);
      } else if (large) {
        for (int m = 0; m < NT; m++) Z += 2.0 * std::pow(r, -2.0 * m - 1.0) * (1.0 - r) / (2.0 * m + 1.0);
      } else if (r == 1) {
        Z = 0.0;
      } else {
        Z = (1.0 - r) * L;
      }
      return Z;
    case 3: {
      const double r3 = r * r * r;
      if (small) {
        Z = r * r;
        for (int m = 0; m < NT; m++) Z += (1.0 - r3) * std::pow(r, 2 * m + 1) / (2.0 * m + 1.0);
      } else if (large) {
        for (int m = 0; m < NT; m++)
          Z += std::pow(r, -2 * m) * ((2.0 * m + 3.0) / r - 2.0 * m - 1.0) / ((2.0 * m + 1.0) * (2.0 * m + 3.0));
      } else if (r == 1) {
        Z = 1.0;
      } else {
        Z = r * r + 0.5 * (1.0 - r3) * L;
      }
      return Z;
    }
    case 5: {
      const double r2 = r * r, r4 = r2 * r2, r5 = r * r * r * r2;
      if (small) {
        Z = r4 + r2 / 3.0;
        for (int m = 0; m < NT; m++) Z += (1.0 - r5) * std::pow(r, 2 * m + 1) / (2.0 * m + 1.0);
      } else if (large) {
        for (int m = 0; m < NT; m++)
          Z += std::pow(r, -2 * m) * ((2.0 * m + 5.0) / r - 2.0 * m - 1.0) / ((2.0 * m + 1.0) * (2.0 * m + 5.0));
      } else if (r == 1) {
        Z = 4.0 / 3.0;
      } else {
        Z = r4 + r2 / 3.0 + 0.5 * (1.0 - r5) * L;
      }
      return Z;
    }
    default:
      std::fprintf(stderr, "rtrg: Z kernel %d is not used by the Time-RG integrals\n", n);
      std::abort();
  }
}

// G_n[d + np - 1] = Z_n(r) r^3 with r = q_m/k_i = exp(-dlnk d), d = i - m (redTime.cc:698-713)
void build_G(const GridSpec &g, int n, std::vector<double> &G) {
  const int np = g.np, zi = zkern_index(n);
  G.assign(2 * np - 1, 0.0);
  for (int d = -(np - 1); d <= np - 1; d++) {
    if (d == 0) {
      G[d + np - 1] = Zreg(zi, 1.0);
    } else {
      const double si = g.dlnk * d, r = std::exp(-si), r2 = r * r, r3 = r * r2;
      G[d + np - 1] = Zreg(zi, r) * r3;
    }
  }
}

void build_extrap_stencil(const GridSpec &g, std::vector<int> &ex_n0, std::vector<double> &ex_w,
                          std::vector<double> &ex_dx) {
  const int nk = g.nk, np = g.np;
  std::vector<double> lnkArr(nk);
  const double lnkmin = std::log(g.kmin);
  for (int i = 0; i < nk; i++) lnkArr[i] = lnkmin + g.dlnk * i;  // rt:1559-1562
  ex_n0.assign(np, 0);
  ex_w.assign((size_t)4 * np, 0.0);
  ex_dx.assign(np, 0.0);
  for (int ip = 0; ip < np; ip++) {
    const double k = std::exp(g.lnk_pad_min + g.dlnk * ip), lnk = std::log(k);
    const int nguess = (int)((lnk - lnkArr[0]) / g.dlnk);
    int n = (nguess > 2 ? nguess - 2 : 0);
    if (n > nk - 1) n = nk - 1;
    while (n < nk - 1 && lnkArr[n + 1] < lnk) n++;
    int type = 0;
    if (n == 0) type = -1;
    if (n == nk - 2) type = 1;
    if (n >= nk - 1 || lnk > lnkArr[nk - 1]) type = 2;
    double *w = &ex_w[(size_t)4 * ip];
    if (type == 0) {
      const double *p = &lnkArr[n - 1];
      ex_n0[ip] = n - 1;
      w[0] = (lnk - p[1]) * (lnk - p[2]) * (lnk - p[3]) / (p[0] - p[1]) / (p[0] - p[2]) / (p[0] - p[3]);
      w[1] = (lnk - p[0]) * (lnk - p[2]) * (lnk - p[3]) / (p[1] - p[0]) / (p[1] - p[2]) / (p[1] - p[3]);
      w[2] = (lnk - p[0]) * (lnk - p[1]) * (lnk - p[3]) / (p[2] - p[0]) / (p[2] - p[1]) / (p[2] - p[3]);
      w[3] = (lnk - p[0]) * (lnk - p[1]) * (lnk - p[2]) / (p[3] - p[0]) / (p[3] - p[1]) / (p[3] - p[2]);
    } else if (type == 2) {
      ex_n0[ip] = nk - 4;
      w[3] = 1.0;
      ex_dx[ip] = lnk - lnkArr[nk - 1];
    } else {
      const int n0 = std::min(n, nk - 4);
      ex_n0[ip] = n0;
      const double t = (lnk - lnkArr[n]) / (lnkArr[n + 1] - lnkArr[n]);
      w[n - n0] = 1.0 - t;
      w[n + 1 - n0] = t;
    }
  }
}

}  // namespace rtrg
