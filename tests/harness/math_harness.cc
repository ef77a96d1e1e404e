// TEST INFRASTRUCTURE (not part of the product): compiles redtime_b200/csrc/rtrg_math.h -- the
// scalar arithmetic every CUDA kernel is a parallel driver around -- with g++ so that the
// look-up rules, the growth ODE (RK8PD + GSL step control), QAG-61 and the Time-RG right-hand
// side can be checked against the oracle's goldens on a machine without a GPU.  The shipped
// library never runs this: it has no CPU execution path.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../redtime_b200/csrc/gk61_table.h"
#include "../../redtime_b200/csrc/rtrg_math.h"

using namespace rtrg;

namespace {
struct World {
  Cosmo c;
  std::vector<double> a, kb, beta, lnkT, lnT, lna, lnkg, G, dD, Dnorm, brow;
  int n_lna = 100, n_lnk = 50;
  double beta_kmin = 1e-3, beta_kmax = 1.0, a_early = 1e-20;
  BetaTab bt() const {
    BetaTab t;
    t.n_z = c.n_z, t.n_kb = c.n_kb, t.a = a.data(), t.k = kb.data(), t.beta = beta.data(), t.row1 = nullptr;
    t.fn = c.On / c.Om, t.kmin = beta_kmin, t.kmax = beta_kmax;
    return t;
  }
  GrowthTab gt() const {
    GrowthTab g;
    g.n_lna = n_lna, g.n_lnk = n_lnk, g.lna = lna.data(), g.lnk = lnkg.data();
    g.G = G.data(), g.dD = dD.data(), g.Dnorm = Dnorm.data();
    return g;
  }
  LinCtx lin() const {
    LinCtx L;
    L.c = &c, L.bt = bt(), L.gt = gt(), L.lnkT = lnkT.data(), L.lnT = lnT.data(), L.nT = c.nT;
    return L;
  }
};
World W;

double qag(int which) {
  const LinCtx L = W.lin();
  QagState w;
  double fv[122], iv[4];
  auto f = [&](double x) { return which == 0 ? sigma8_integrand(L, x) : sigmav_integrand(L, x); };
  for (int s = 0; s < 61; s++) fv[s] = f(qk61_abscissa(gk61_xgk, -15, 15, s));
  bool done = qag_begin(w, -15, 15, 0, 1e-4, qk61_combine(gk61_xgk, gk61_wgk, gk61_wg, -15, 15, fv));
  while (!done) {
    qag_next(w, &iv[0], &iv[1], &iv[2], &iv[3]);
    for (int s = 0; s < 122; s++) fv[s] = f(qk61_abscissa(gk61_xgk, iv[2 * (s >= 61)], iv[2 * (s >= 61) + 1], s % 61));
    done = qag_update(w, qk61_combine(gk61_xgk, gk61_wgk, gk61_wg, iv[0], iv[1], fv),
                      qk61_combine(gk61_xgk, gk61_wgk, gk61_wg, iv[2], iv[3], fv + 61), 1000);
  }
  return qag_result(w);
}
}  // namespace

extern "C" {

// params[9], z_in, transfer table, interpolation tables (as rtrg_cosmology); builds the beta table,
// the growth tables (hdr:639-731) and the sigma_8 normalisation (hdr:846-879)
int mh_setup(const double *params, double z_in, int nT, const double *kT, const double *TcT, const double *TbT, int n_z,
             const double *z_interp, int n_kb, const double *kb, const double *Tcb, const double *Tnub) {
  Cosmo &c = W.c;
  std::memset(&c, 0, sizeof c);
  c.ns = params[0], c.s8 = params[1], c.h = params[2], c.Om = params[3], c.Ob = params[4], c.On = params[5];
  c.TK = params[6], c.w0 = params[7], c.wa = params[8], c.z_in = z_in;
  cosmo_derive(c);
  c.nT = nT, c.n_z = n_z, c.n_kb = n_kb;
  const double f_b = c.Ob / (c.Om - c.On), f_c = 1.0 - f_b, fn = c.On / c.Om;
  W.lnkT.resize(nT), W.lnT.resize(nT);
  const double T0 = f_b * TbT[0] + f_c * TcT[0];
  for (int i = 0; i < nT; i++) W.lnkT[i] = std::log(kT[i]), W.lnT[i] = std::log((f_b * TbT[i] + f_c * TcT[i]) / T0);
  W.a.resize(n_z), W.kb.assign(kb, kb + n_kb), W.beta.resize((size_t)n_z * n_kb);
  for (int i = 0; i < n_z; i++) W.a[i] = 1.0 / (1.0 + z_interp[i]);
  for (size_t i = 0; i < W.beta.size(); i++) W.beta[i] = fn * Tnub[i] / Tcb[i];
  // growth tables: one RK8PD integration per growth wavenumber through the n_lna+1 legs
  const int nj = W.n_lnk + 1, ni = W.n_lna + 1;
  W.lna.resize(ni), W.lnkg.resize(nj), W.G.assign((size_t)ni * nj, 0), W.dD.assign((size_t)ni * nj, 0), W.Dnorm.resize(nj);
  for (int i = 0; i < ni; i++) W.lna[i] = std::log(GROWTH_A_MIN) + std::log(GROWTH_A_MAX / GROWTH_A_MIN) / W.n_lna * i;
  for (int j = 0; j < nj; j++) W.lnkg[j] = std::log(GROWTH_K_MIN) + std::log(GROWTH_K_MAX / GROWTH_K_MIN) / W.n_lnk * j;
  const PDTableau PD = make_pd_tableau();
  const BetaTab bt = W.bt();
  W.brow.assign((size_t)std::max(n_z, 1) * nj, 0.0);
  for (int j = 0; j < nj; j++) {
    double k = std::exp(W.lnkg[j]);
    k = std::min(std::max(k, W.beta_kmin), W.beta_kmax);
    if (n_z > 0 && fn >= 1e-10) {
      const Stencil st = tab_stencil_y(bt.k, bt.n_kb, k);
      for (int iz = 0; iz < n_z; iz++) W.brow[(size_t)iz * nj + j] = stencil_apply(st, bt.beta + (size_t)iz * n_kb);
    }
    GrowthCtx g;
    g.bg = bg_static(c), g.bt = bt, g.brow = W.brow.data() + j, g.bstride = nj;
    double y[2] = {1.0, 1.0 / W.a_early};
    growth_integrate(PD, g, W.a_early, GROWTH_A_MIN, y);
    W.G[j] = y[0] / GROWTH_A_MIN, W.dD[j] = y[1];
    for (int i = 1; i < ni; i++) {
      const double a1 = std::exp(W.lna[i]);
      growth_integrate(PD, g, std::exp(W.lna[i - 1]), a1, y);
      W.G[(size_t)i * nj + j] = y[0] / a1, W.dD[(size_t)i * nj + j] = y[1];
    }
  }
  const GrowthTab gt = W.gt();
  for (int j = 0; j < nj; j++) W.Dnorm[j] = tab2d(gt.lna, ni, gt.lnk, nj, gt.G, 0.0, gt.lnk[j]);
  c.Norm = c.s8 * c.s8 / qag(0);
  c.sigv2_0 = qag(1) / (6.0 * M_PI * M_PI);
  return 0;
}
double mh_norm() { return W.c.Norm; }
double mh_sigv2_0() { return W.c.sigv2_0; }
void mh_lookups(double z, const double *k, int n, double *D, double *dD, double *beta, double *P, double *Pcb, double *Pnu) {
  const LinCtx L = W.lin();
  for (int i = 0; i < n; i++) {
    growth_D_dD(L.gt, z, k[i], &D[i], &dD[i]);
    beta[i] = beta_P(L.bt, 1.0 / (1.0 + z), k[i]);
    P[i] = plin(L, z, k[i]), Pcb[i] = plin_cb(L, z, k[i]), Pnu[i] = plin_nu(L, z, k[i]);
  }
}
// full Time-RG right-hand side with the oracle's A (64 slots) and R for the state y (rt:1416-1547)
void mh_rhs(double eta, int nk, const double *kgrid, const double *y, const double *A64, const double *R24, int evolve_Q,
            double *dy) {
  static const int JU[14] = {8, 9, 10, 11, 12, 13, 14, 15, 56, 57, 59, 60, 61, 63};
  const Cosmo &c = W.c;
  const BetaTab bt = W.bt();
  const double A = c.a_in * std::exp(eta);
  for (int i = 0; i < nk; i++) {
    double yi[N_U], dyi[N_U], A14[N_UI], R[N_UQ], Om10, Om11;
    for (int j = 0; j < N_U; j++) yi[j] = y[(size_t)j * nk + i];
    for (int j = 0; j < N_UI; j++) A14[j] = A64[(size_t)JU[j] * nk + i];
    for (int j = 0; j < N_UQ; j++) R[j] = R24[(size_t)j * nk + i];
    trg_omega(c, A, beta_P(bt, A, kgrid[i]), &Om10, &Om11);
    trg_rhs_row(eta, kgrid[i], Om10, Om11, 1, evolve_Q, yi, A14, R, dyi);
    for (int j = 0; j < N_U; j++) dy[(size_t)j * nk + i] = dyi[j];
  }
}
// gsl_odeiv std_control_hadjust, control_y_new
int mh_hadjust(double rmax, int ord, double *h) { return gsl_hadjust(rmax, ord, h); }
}
