// C-ABI access to the cosmology-independent tables (host side; no GPU needed).
#include <cstring>
#include <vector>

#include "../../include/redtime_b200.h"
#include "fastpt_tables.h"

using namespace rtrg;

extern "C" {

int rtrg_grid_info(int nk, double kmin, double kmax, int out[5], double *dlnk,
                   double *lnk_pad_min) {
  if (nk < 16 || (nk % 16) != 0 || !(kmin > 0) || !(kmax > kmin)) return RTRG_EINVAL;
  GridSpec g = make_grid(nk, kmin, kmax);
  if (out) {
    out[0] = g.np;
    out[1] = g.nshift;
    out[2] = g.jlo;
    out[3] = g.nsup;
    out[4] = g.nloMR;
  }
  if (dlnk) *dlnk = g.dlnk;
  if (lnk_pad_min) *lnk_pad_min = g.lnk_pad_min;
  return RTRG_OK;
}

int rtrg_table_T(int nk, double kmin, double kmax, int n, double *T, double *kfac) {
  if (n < 0 || n >= N_JKERN || nk < 16 || (nk % 16) != 0) return RTRG_EINVAL;
  GridSpec g = make_grid(nk, kmin, kmax);
  std::vector<double> Tv, kf;
  build_T(g, n, Tv, kf);
  if (T) std::memcpy(T, Tv.data(), Tv.size() * sizeof(double));
  if (kfac) std::memcpy(kfac, kf.data(), kf.size() * sizeof(double));
  return RTRG_OK;
}

int rtrg_table_G(int nk, double kmin, double kmax, int n, double *G) {
  if (n < 0 || n >= N_ZKERN || nk < 16 || (nk % 16) != 0) return RTRG_EINVAL;
  GridSpec g = make_grid(nk, kmin, kmax);
  std::vector<double> Gv;
  build_G(g, n, Gv);
  if (G) std::memcpy(G, Gv.data(), Gv.size() * sizeof(double));
  return RTRG_OK;
}

int rtrg_table_windows(int nk, double kmin, double kmax, double *WP, double *WC) {
  if (nk < 16 || (nk % 16) != 0) return RTRG_EINVAL;
  GridSpec g = make_grid(nk, kmin, kmax);
  for (int i = 0; i < g.np; i++) {
    if (WP) WP[i] = window_P(g, i);
    if (WC) WC[i] = window_C(g, i);
  }
  return RTRG_OK;
}

int rtrg_table_extrap(int nk, double kmin, double kmax, int *n0, double *w, double *dx) {
  if (nk < 16 || (nk % 16) != 0) return RTRG_EINVAL;
  GridSpec g = make_grid(nk, kmin, kmax);
  std::vector<int> vn0;
  std::vector<double> vw, vdx;
  build_extrap_stencil(g, vn0, vw, vdx);
  if (n0) std::memcpy(n0, vn0.data(), vn0.size() * sizeof(int));
  if (w) std::memcpy(w, vw.data(), vw.size() * sizeof(double));
  if (dx) std::memcpy(dx, vdx.data(), vdx.size() * sizeof(double));
  return RTRG_OK;
}

int rtrg_assembly_terms(int *row, int *src, int *index, int *kpow, double *coef, int cap) {
  const std::vector<AsmTerm> &t = assembly_terms();
  const int n = (int)t.size();
  if (row && src && index && kpow && coef) {
    for (int i = 0; i < n && i < cap; i++) {
      row[i] = t[i].row;
      src[i] = t[i].src;
      index[i] = t[i].index;
      kpow[i] = t[i].kpow;
      coef[i] = t[i].coef;
    }
  }
  return n;
}

}  // extern "C"
