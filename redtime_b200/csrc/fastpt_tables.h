// Cosmology-independent weight tables of the Time-RG mode-coupling integrals.
//
// The reference evaluates J_{alpha,beta,ell}(k) with FAST-PT (FFT-log) transforms
// (src/redTime.cc:411-597) and the P13-type terms with an analytic-mu kernel Z_n
// convolved in log q (src/redTime.cc:599-727).  Both are *linear / bilinear* in the
// tabulated power spectra with weights that depend only on the (nk, kmin, kmax) grid:
//
//   J_n(k_i; A,B) = kfac_n(i) * sum_{j,l} a_j b_l T_n[(i-j) mod np][(i-l) mod np]
//                   a_j = P_A(q_j) q_j^2,  b_l = P_B(q_l) q_l^2
//   PZ_n(k_i;A)   = dlnk/(2 pi^2) k_i^3 P_00(k_i) sum_m P_A(q_m) G_n[i-m]
//
// so the device kernels are dense (q1,q2) quadratures over the padded log-k grid.  This
// file builds T_n and G_n once per grid, in extended precision (long double), so that
// the quadrature reproduces the reference's FFT result to round-off.
#pragma once
#include <vector>

namespace rtrg {

struct GridSpec {
  int nk;       // output wavenumbers (reference: nk=128, redTime.cc:93)
  int np;       // padded grid, 4*nk (redTime.cc:93)
  int nshift;   // (np-nk)/2
  double kmin, kmax, dlnk, lnk_pad_min;
  int jlo;      // first padded index with WP > 0
  int nsup;     // np - jlo : support of the windowed spectra
  int nloMR;    // nshift - nk/2 (redTime.cc:1252)
};

GridSpec make_grid(int nk, double kmin, double kmax);

// window on the padded spectrum (redTime.cc:102-127) and on the Fourier coefficients (:129-138)
double window_P(const GridSpec &g, int ipad);
double window_C(const GridSpec &g, int m);

// Pab on the padded grid (rt:181-232 with the tabulated-function rules itp:68-78) as a 4-point
// stencil per padded sample ip:
//   ln P_ab(k_pad[ip]) = sum_{j<4} w[4 ip + j] lnP_ab[n0[ip] + j] + (n_s - 3) dx[ip]
// Inside the grid it is the 4-point Lagrange interpolant (the first and last interval fall back
// to the linear one), above kmax the power-law tail k^(n_s-3) anchored at the last sample.
void build_extrap_stencil(const GridSpec &g, std::vector<int> &n0, std::vector<double> &w,
                          std::vector<double> &dx);

// The 14 bilinear kernels: 0..6 = J (alpha,-alpha,ell) (redTime.cc:731-732; n=1 regularised),
//                          7..13 = Jn0 (redTime.cc:734-736)
#ifndef RTRG_NKERN_DEFINED
#define RTRG_NKERN_DEFINED
enum { N_JKERN = 14, N_ZKERN = 7 };
#endif
struct KernSpec { int alpha, beta, ell; bool reg; };
KernSpec kern_spec(int n);

// T_n, full circulant np x np, row-major T[u*np+v] (u: alpha-side lag, v: beta-side lag),
// including the reference's constant prefactor; kfac[i], i<np, is the k-dependent prefactor.
void build_T(const GridSpec &g, int n, std::vector<double> &T, std::vector<double> &kfac);

// G_n[d + np - 1], d = i - m in [-(np-1), np-1], for the 7 Z kernels {0,1,-1,3,-3,5,-5}
int zkern_index(int n);  // n -> Z index
double Zreg(int n, double r);
void build_G(const GridSpec &g, int n, std::vector<double> &G);

// ---- linear assembly  J,PZ,Jn0 -> A (14 unique), R (24), PTjm (9), PMRn (8) ----------
// One term:  out[row] += coef * k^kpow * src[index]   (src: 0 = J, 1 = PZ, 2 = Jn0,
// 3 = J[0][00,00] at the low-k row nloMR); index = 9*n + 3*ab + cd as in redTime.cc:785.
// `row` enumerates outputs: 0..13 A (order of JU[], redTime.cc:157), 14..37 R
// ((ell-1)*8+4a+2b+c), 38..46 PTjm, 47..54 PMRn.  Prefactors pre_A = k/(4 pi),
// pre_R = 1/(2 pi k) (redTime.cc:815-816) are folded into coef/kpow.
struct AsmTerm { short row, src, index, kpow; double coef; };
enum { ASM_NROWS = 55, ASM_A0 = 0, ASM_R0 = 14, ASM_PT0 = 38, ASM_PMR0 = 47 };
const std::vector<AsmTerm> &assembly_terms();

}  // namespace rtrg
