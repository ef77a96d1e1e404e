// Stage-level oracle driver -- TEST INFRASTRUCTURE ONLY.
//
// Pulls the UNMODIFIED reference translation unit in (its `main` renamed) so that the
// reference's own functions can be called stage by stage from tests:
//   Pab/WP (redTime.cc:117-232), J_MFHB (514-597), PZ_reg (689-727),
//   compute_Aacdbef_Rlabc_PTjm_PMRn_full (740-1282), derivatives (1416-1547),
//   C.D_dD / C.Beta_P / C.Plin_cb / C.Plin_nu / C.sigmaV2 (AU_cosmological_parameters.h).
// No reference source is copied: REF_MAIN_TU is the path of the reference file, given
// by oracle/Makefile.  The reference's global `cosmological_parameters C` is constructed
// when this library is loaded and reads ./params_redTime.dat, so load it with the run
// directory as CWD (one process per cosmology; the reference keeps function-local statics).
#include <cstdio>
#define main redTime_reference_main
#include REF_MAIN_TU
#undef main

extern "C" {

int ref_nk() { return nk; }
int ref_np() { return np; }
int ref_nU() { return nU; }
double ref_dlnk() { return dlnk; }
double ref_z_in() { return C.z_in(); }
int ref_n_out() { return C.n_eta(); }
double ref_z_out(int i) { return C.zsteps(i); }
double ref_eta_out(int i) { return C.etasteps(i); }
int ref_switch(int i) {
  switch (i) {
  case 0: return C.SWITCH_NONLINEAR();
  case 1: return C.SWITCH_1LOOP();
  case 2: return C.PRINTLIN();
  default: return C.PRINTRSD();
  }
}

// redTime.cc:1559-1568: k grid + the serial forced initialisations of main()
void ref_init(double *k_out) {
  for (int i = 0; i < nk; i++) {
    lnkArr[i] = lnkmin + dlnk * i;
    kArr[i] = exp(lnkArr[i]);
    if (k_out) k_out[i] = kArr[i];
  }
  double Ddum[2];
  C.D_dD(C.z_in(), kArr[0], Ddum);
  C.Plin_cb(C.z_in(), kArr[0]);
}

// redTime.cc:1570-1586
void ref_initial_y(double *y) {
  for (int i = 0; i < nk; i++) {
    double D_in[2];
    C.D_dD(C.z_in(), kArr[i], D_in);
    double f_in = C.a_in() * D_in[1] / D_in[0];
    double Pin_i = C.Plin_cb(C.z_in(), kArr[i]);
    y[i] = log(Pin_i);
    y[nk + i] = log(Pin_i * f_in);
    y[2 * nk + i] = log(Pin_i * f_in * f_in);
  }
  for (int i = nUP * nk; i < nU * nk; i++) y[i] = 0.0;
}

// redTime.cc:772-778 (the extrapolated + windowed spectra fed to the integrals)
void ref_extrap_P(const double *y, double *P3np) {
  for (int i = 0; i < np; i++) {
    double k = exp(lnk_pad_min + dlnk * i), Win = WP(lnk_pad_min + dlnk * i);
    P3np[i] = Pab(0, 0, k, y) * Win;
    P3np[i + np] = Pab(0, 1, k, y) * Win;
    P3np[i + 2 * np] = Pab(1, 1, k, y) * Win;
  }
}
double ref_WP(int i) { return WP(lnk_pad_min + dlnk * i); }
double ref_WC(int i) { return WC(i); }

int ref_J_MFHB(int alpha, int beta, int ell, const double *Pa, const double *Pb, double *J) {
  return J_MFHB(alpha, beta, ell, Pa, Pb, J);
}
int ref_PZ_reg(int n, const double *Pq, const double *Pk, double *PZn) {
  return PZ_reg(n, Pq, Pk, PZn);
}
double ref_Zreg_n(int n, double r) { return Zreg_n(n, r); }

int ref_compute_full(double eta, const double *y, double *A64, double *R, double *PTjm,
                     double *PMRn) {
  return compute_Aacdbef_Rlabc_PTjm_PMRn_full(eta, y, A64, R, (double(*)[nk])PTjm,
                                              (double(*)[nk])PMRn);
}
// the wrapper the RHS uses (1-loop cache or full, redTime.cc:1343-1361)
int ref_compute_PTj(double eta, const double *y, double *A64, double *R, double *PT4nk) {
  return compute_Aacdbef_Rlabc_PTj(eta, y, A64, R, PT4nk, PT4nk + nk, PT4nk + 2 * nk,
                                   PT4nk + 3 * nk);
}
int ref_derivatives(double eta, const double *y, double *dy) {
  return derivatives(eta, y, dy, nullptr);
}
double ref_Omega(int i, int j, double A, double k) { return Omega(i, j, A, k); }

void ref_D_dD(double z, const double *k, int n, double *D, double *dDda) {
  for (int i = 0; i < n; i++) {
    double o[2];
    C.D_dD(z, k[i], o);
    D[i] = o[0];
    dDda[i] = o[1];
  }
}
void ref_Beta_P(double a, const double *k, int n, double *beta) {
  for (int i = 0; i < n; i++) beta[i] = C.Beta_P(a, k[i]);
}
void ref_Plin_cb(double z, const double *k, int n, double *P) {
  for (int i = 0; i < n; i++) P[i] = C.Plin_cb(z, k[i]);
}
void ref_Plin_nu(double z, const double *k, int n, double *P) {
  for (int i = 0; i < n; i++) P[i] = C.Plin_nu(z, k[i]);
}
void ref_Plin(double z, const double *k, int n, double *P) {
  for (int i = 0; i < n; i++) P[i] = C.Plin(z, k[i]);
}
double ref_sigmaV2(double z) { return C.sigmaV2(z); }
double ref_H_H0(double a) { return C.H_H0(a); }
double ref_H2_H02(double a) { return C.H2_H02(a); }
double ref_dlnH_dlna(double a) { return C.dlnH_dlna(a); }

// Pbisj (redTime.cc:269-298)
double ref_Pbisj(int i, int j_mu, int m_b, const double *y) { return Pbisj(i, j_mu, m_b, y); }

// run the reference main() itself (prints to stdout)
int ref_main() { return redTime_reference_main(); }

} // extern "C"
