// Device code shared by the stand-alone kernels of one right-hand-side evaluation (k_assemble,
// k_rhs, k_combine, k_final) and by k_stage_post, which runs the same pieces back to back for launches
// too small to fill the GPU (a single cosmology, k-sharded ranks): same functions, same order of
// operations, same bits.
#pragma once
#include "rtrg_device.h"

namespace rtrg {

// ---------------------------------------------------------------------------- assembly
// A_{acd,bef}, R^l_{abc}, P_T,jm, P_MR,n from the 190 raw integrals of a row and the term table of
// assembly_table.cc (rt:820-1278).
__device__ __forceinline__ double kpow_i(double k, double kinv, int p) {
  double r = 1.0;
  if (p > 0)
    for (int i = 0; i < p; i++) r *= k;
  else
    for (int i = 0; i < -p; i++) r *= kinv;
  return r;
}
// ASM_ROWS rows per CTA: 16 for batches (the term table is read once per 16 rows), 4 when the
// launch is small, where more CTAs shorten the critical path.  The arithmetic per row is the same.
enum { ASM_NV = 190, ASM_MAXT = 768 };
template <int ASM_ROWS>
struct AsmShared {
  double vals[ASM_NV][ASM_ROWS + 1];
  // the assembly table (730 terms, 10 KB) once per CTA: the term loops then run out of shared memory
  // instead of chains of dependent global loads
  double coef[ASM_MAXT];
  short src[ASM_MAXT], index[ASM_MAXT], kpow[ASM_MAXT];
  int start[N_SRC + 1];
};
template <int ASM_ROWS>
__device__ __forceinline__ void asm_load_table(const IntegralTabs &tb, AsmShared<ASM_ROWS> &sa) {
  for (int t = threadIdx.x; t < tb.n_terms; t += blockDim.x) {
    sa.coef[t] = tb.t_coef[t];
    sa.src[t] = tb.t_src[t];
    sa.index[t] = tb.t_index[t];
    sa.kpow[t] = tb.t_kpow[t];
  }
  for (int t = threadIdx.x; t <= N_SRC; t += blockDim.x) sa.start[t] = tb.t_start[t];
}
// the 190 raw values of rows r0 .. r0 + rows - 1 of cosmology e (all threads of the CTA)
template <int ASM_ROWS>
__device__ __forceinline__ void asm_gather_vals(const IntegralTabs &tb, int has_jn0, const double *__restrict__ Jpart,
                                                const double *__restrict__ PZb, const double *__restrict__ P3,
                                                const double *__restrict__ Jlo, double *__restrict__ raw, int e, int r0,
                                                int rows, int groups, AsmShared<ASM_ROWS> &sa) {
  // only the values the requested groups consume are read (the others were not computed by this
  // evaluation); with raw != nullptr all of them
  unsigned long long want[3] = {0, 0, 0};
  for (int gi = 0; gi < 4; gi++)
    if (groups & (1 << gi))
      for (int w = 0; w < 3; w++) want[w] |= tb.need_val[gi][w];
  if (raw || (groups & GRP_RAW)) want[0] = want[1] = want[2] = ~0ULL;
  for (int idx = threadIdx.x; idx < ASM_NV * ASM_ROWS; idx += blockDim.x) {
    const int v = idx / ASM_ROWS, rr = idx - v * ASM_ROWS;
    if (rr >= rows) continue;
    if (!((want[v >> 6] >> (v & 63)) & 1ULL)) {
      sa.vals[v][rr] = 0.0;
      continue;
    }
    const int i = r0 + rr, ipad = tb.nshift + i;
    double x = 0.0;
    if (v < 63 || (v >= 126 && v < 189)) {
      const int iJ = (v < 63) ? v : v - 126;
      const int n = iJ / 9 + ((v < 63) ? 0 : 7), pair = iJ % 9;
      if (v < 63 || has_jn0) {
        // the parts k_bilinear wrote for this row block: per split of the beta-side lags, one per
        // CTA along the item axis that holds some of the block's alpha-side lags
        const int nch = tb.nchunk * tb.vsplit, rb = i / BIL_R;
        const int np_rb = (rb * tb.NV + tb.NV - 1) / tb.tpb - (rb * tb.NV) / tb.tpb + 1;
        RT_ASSERT(np_rb >= 1 && np_rb <= tb.nchunk && n < N_JKERN && pair < 9);
        // added in (split, part) order; the loads of four parts are issued together -- one memory
        // round trip per four parts instead of one per part (x + 0.0 is exact for the padding)
        const double *pj = Jpart + ((((long long)e * N_JKERN + n) * nch) * 9 + pair) * tb.nk + i;
        const long long pstride = 9LL * tb.nk;
        const int nq = tb.vsplit * np_rb;
        for (int q0 = 0; q0 < nq; q0 += 4) {
          double t4[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int q = q0 + u, part = (q / np_rb) * tb.nchunk + q % np_rb;
            t4[u] = q < nq ? pj[part * pstride] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) x += t4[u];
        }
        x *= tb.kfac[n * tb.nk + i];
      }
    } else if (v < 126) {
      const int iJ = v - 63, n = iJ / 9, ab = (iJ % 9) / 3, cd = iJ % 3;
      x = PZb[(((long long)e * N_ZKERN + n) * 3 + ab) * tb.nk + i];
      if (cd) {  // rt:797-800
        const double *P = P3 + (long long)e * 3 * tb.np;
        x = x * P[cd * tb.np + ipad] / (P[ipad] + 1e-100);
      }
    } else {
      x = Jlo[e];
    }
    sa.vals[v][rr] = x;
    if (raw) raw[((long long)e * ASM_NV + v) * tb.nk + i] = x;
  }
}
// output groups: A rows 0-13, R 14-37, P_T,jm 38-46, P_MR,n 47-54; rows of groups that were not
// requested keep their old content (their inputs were not computed)
__device__ __forceinline__ bool asm_wanted(int o, int groups) {
  const int grp = o < 14 ? GRP_A : o < 38 ? GRP_R : o < 47 ? GRP_PT : GRP_PMR;
  return (groups & grp) != 0;
}
// source row o at wavenumber i (row rr of the CTA)
template <int ASM_ROWS>
__device__ __forceinline__ double asm_source(const IntegralTabs &tb, const AsmShared<ASM_ROWS> &sa, int o, int rr, int i) {
  const double k = tb.kgrid[i], kinv = 1.0 / k;
  double acc = 0.0;
  const int t1 = sa.start[o + 1];
  for (int t = sa.start[o]; t < t1; t++) {
    const int s = sa.src[t];
    const int v = (s == 3) ? 189 : s * 63 + sa.index[t];
    RT_ASSERT(v >= 0 && v < ASM_NV && t < tb.n_terms);
    acc += sa.coef[t] * kpow_i(k, kinv, sa.kpow[t]) * sa.vals[v][rr];
  }
  return acc;
}

// ---------------------------------------------------------------------------- right-hand side
// linear quantities at grid wavenumber i through the pre-reduced rows (see kernels_linear.cu)
__device__ __forceinline__ double ode_row_beta(const Batch &S, int b, int i, double a) {
  const BetaTab t = beta_tab(S, S.cosmo[b]);
  return beta_row(t, S.bred + (long long)b * S.n_zmax * S.nkk + i, a, S.nkk);
}
__device__ __forceinline__ bool ode_row_D_dD(const Batch &S, int b, int i, double z, double *D,
                                             double *dD) {
  const long long o = (long long)b * (S.n_lna + 1) * S.nk + i;
  return growth_D_dD_row(S.lna, S.n_lna, S.Grow + o, S.dDrow + o, S.nk,
                         S.D0row[(long long)b * S.nk + i], z, D, dD);
}

// ---------------------------------------------------------------------------- k_rhs
// stage < 0: eta = t[b]; otherwise eta = t[b] + c_stage * h_try[b].
// Everything that depends on the time only (background, interpolation weights in a of the beta
// and growth tables, powers of e^eta) is evaluated once per block and shared: the 128 rows of a
// block belong to one cosmology and are evaluated at the same eta.  The rows then read their 41
// state values, 38 sources and a few table entries: HBM traffic 960 B per row.
struct RhsShared {
  double eta, eeta, A, om10_den, Om11, z, a, pre4;
  RowX xb, xg;
  int beta_zero, beta_bad, growth_ok;
};
// time-only part of one right-hand-side evaluation at eta (one thread per block)
__device__ __forceinline__ void rhs_time_setup(const Batch &S, const Cosmo &c, double eta, RhsShared &sh) {
  const double A = c.a_in * exp(eta);  // rt:1430
  sh.eta = eta;
  sh.eeta = exp(eta);
  sh.A = A;
  sh.om10_den = A * A * A * bg_H2(c, A);  // rt:1395-1401
  sh.Om11 = 3.0 + bg_dlnH(c, A);
  // Beta_P(A, k): 0 without massive neutrinos, abort in the reference for A > 1.001 (hdr:523-531)
  sh.beta_zero = (c.n_z == 0 || c.On / c.Om < 1e-10);
  sh.beta_bad = (!sh.beta_zero && A > 1.001);
  if (!sh.beta_zero && !sh.beta_bad) sh.xb = tab_row_x_prepare(S.in + c.offA, c.n_z, A > 1.0 ? 1.0 : A);
  // growth look-up of the 1-loop rescaling (rt:1316-1337)
  sh.z = exp(-eta) * (1.0 + c.z_in) - 1;
  sh.a = 1.0 / (sh.z + 1.0);
  sh.growth_ok = !(sh.a > GROWTH_A_MAX || sh.a < GROWTH_A_MIN);
  if (sh.growth_ok) sh.xg = tab_row_x_prepare(S.lna, S.n_lna + 1, log(sh.a));
  sh.pre4 = exp(-4.0 * eta);
}
// The same, by all threads of a block together (k_rhs): the two table look-ups of rhs_time_setup are
// binary searches -- chains of 4 + 7 dependent global loads, ~7 us of pure latency in front of every
// right-hand side of a single cosmology.  Here every thread compares ONE node with the abscissa and
// __syncthreads_count adds the votes: tab_find(x, n, xq) = #{1 <= i <= n-2 : x[i] < xq} for a sorted
// table, one memory round trip.  Same index, same weights, same bits.
__device__ __forceinline__ void rhs_time_setup_block(const Batch &S, const Cosmo &c, double eta, RhsShared &sh) {
  const double A = c.a_in * exp(eta);
  const bool beta_zero = (c.n_z == 0 || c.On / c.Om < 1e-10), beta_bad = (!beta_zero && A > 1.001);
  const bool need_b = !beta_zero && !beta_bad;
  const double xb = A > 1.0 ? 1.0 : A;
  const double z = exp(-eta) * (1.0 + c.z_in) - 1, a = 1.0 / (z + 1.0);
  const bool growth_ok = !(a > GROWTH_A_MAX || a < GROWTH_A_MIN);
  const double xg = log(a);
  const int nA = c.n_z, nG = S.n_lna + 1, nmax = nA > nG ? nA : nG;
  const double *xa = S.in + c.offA;
  int nb = 0, ng = 0;
  for (int base = 0; base < nmax; base += blockDim.x) {
    const int i = base + threadIdx.x;
    nb += __syncthreads_count(need_b && i >= 1 && i <= nA - 2 && xa[i] < xb);
    ng += __syncthreads_count(growth_ok && i >= 1 && i <= nG - 2 && S.lna[i] < xg);
  }
  if (threadIdx.x == 0) {
    sh.eta = eta;
    sh.eeta = exp(eta);
    sh.A = A;
    sh.om10_den = A * A * A * bg_H2(c, A);  // rt:1395-1401
    sh.Om11 = 3.0 + bg_dlnH(c, A);
    sh.beta_zero = beta_zero;
    sh.beta_bad = beta_bad;
    if (need_b) sh.xb = tab_row_x_prepare_at(xa, nA, xb, nb);
    sh.pre4 = exp(-4.0 * eta);
  }
  if (threadIdx.x == 32 % blockDim.x) {
    sh.z = z;
    sh.a = a;
    sh.growth_ok = growth_ok;
    if (growth_ok) sh.xg = tab_row_x_prepare_at(S.lna, nG, xg, ng);
  }
  __syncthreads();
}
// row-dependent coefficients of that evaluation: Omega_10, and in 1-loop mode the rescaling of the
// z1l cache, sources x (D/D_z1l)^4 e^{-4 eta} f^n (rt:1316-1337): *pre and f = *fz
__device__ __forceinline__ void rhs_row_coeffs(const Batch &S, const Cosmo &c, int b, int i, const RhsShared &sh,
                                               double *Om10, double *pre, double *fz) {
  const int nk = S.nk;
  double beta = 0.0;
  if (sh.beta_bad) beta = NAN;
  else if (!sh.beta_zero) beta = tab_row_x_apply(sh.xb, S.bred + (long long)b * S.n_zmax * S.nkk + i, S.nkk);
  *Om10 = -1.5 * c.Om * (c.fcb + beta) / sh.om10_den;
  *pre = 1.0;
  *fz = 1.0;
  if (c.sw_nl && c.sw_1l) {
    double D = NAN, dD = NAN;
    if (sh.growth_ok) {
      const long long o = (long long)b * (S.n_lna + 1) * nk + i;
      const double D0 = S.D0row[(long long)b * nk + i];
      D = tab_row_x_apply(sh.xg, S.Grow + o, nk) * sh.a / D0;
      dD = tab_row_x_apply(sh.xg, S.dDrow + o, nk) / D0;
    }
    *fz = dD / (D * (1.0 + sh.z));
    const double rD = D / S.D_z1l[(long long)b * nk + i];
    *pre = (rD * rD) * (rD * rD) * sh.pre4;
  }
}

// One row of the right-hand side.  piece < 0: all four pieces (ln P + I, then the three multipoles of
// Q) one after the other; otherwise only that piece.  s1[j sstride] is source j of this row, yb /
// db[j nk] component j of the state / derivative.  At most 17 + 14 + 17 values are live at a time:
// 4x the occupancy of holding all 41 + 38 + 41.
__device__ __forceinline__ void rhs_row(const Batch &S, const Cosmo &c, int b, int i, int piece, const RhsShared &sh,
                                        double k, const double *s1, long long sstride, const double *__restrict__ yb,
                                        double *__restrict__ db) {
  const int nk = S.nk;
  const double eeta = sh.eeta;
  const int one_loop = c.sw_nl && c.sw_1l;
  const int evolve_Q = (S.print_Q || c.sw_pr);
  double Om10, pre, fz;
  rhs_row_coeffs(S, c, b, i, sh, &Om10, &pre, &fz);
  const double Om11 = sh.Om11;
  double fp[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
  if (one_loop) {
#pragma unroll
    for (int p = 1; p < 5; p++) fp[p] = fp[p - 1] * fz;
  }
  if (piece < 0 || piece == 0) {
    double y[N_UP + N_UI], dy[N_UP + N_UI], A14[N_UI];
#pragma unroll
    for (int j = 0; j < N_UP + N_UI; j++) y[j] = yb[(long long)j * nk];
#pragma unroll
    for (int j = 0; j < N_UI; j++)
      A14[j] = !c.sw_nl ? 0.0 : one_loop ? pre * fp[a14_fpow(j)] * s1[(long long)j * sstride] : s1[(long long)j * sstride];
    trg_rhs_PI(eeta, k, Om10, Om11, c.sw_nl, y, A14, dy);
#pragma unroll
    for (int j = 0; j < N_UP + N_UI; j++) db[(long long)j * nk] = dy[j];
  }
#pragma unroll
  for (int l = 0; l < 3; l++) {
    if (piece >= 0 && piece != l + 1) continue;
    double Q[8], R[8], dQ[8];
    const int j0 = N_UP + N_UI + 8 * l;
    if (c.sw_nl && evolve_Q) {
#pragma unroll
      for (int j = 0; j < 8; j++) Q[j] = yb[(long long)(j0 + j) * nk];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const double r = s1[(long long)(N_UI + 8 * l + j) * sstride];
        R[j] = one_loop ? pre * fp[r24_fpow(8 * l + j)] * r : r;
      }
      trg_rhs_Q(eeta, Om10, Om11, Q, R, dQ);
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) dQ[j] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) db[(long long)(j0 + j) * nk] = dQ[j];
  }
}

// ---------------------------------------------------------------------------- Runge-Kutta pieces
// ytmp of stage `stage` for one component: y + h sum_{j < stage} a_{stage,j} k_j (k_j at kj[j ks])
__device__ __forceinline__ double rk_combine(int stage, double y, double h, const double *kj, long long ks) {
  double acc = 0.0;
  for (int j = 0; j < stage; j++) {
    const double a = RKF45::a(stage, j);
    if (a != 0.0) acc += a * kj[j * ks];
  }
  return y + h * acc;
}
// 5th-order solution and error estimate of one component; returns |yerr| / (eps_rel |ynew| + eps_abs)
__device__ __forceinline__ double rk_final(const Batch &S, double y, double h, const double *kj, long long ks, double *yn_out,
                                           double *ye_out) {
  double acc = 0.0, err = 0.0;
#pragma unroll
  for (int j = 0; j < RK_STAGES; j++) {
    const double k = kj[j * ks];
    if (RKF45::b(j) != 0.0) acc += RKF45::b(j) * k;
    if (RKF45::e(j) != 0.0) err += RKF45::e(j) * k;
  }
  const double yn = y + h * acc, ye = h * err;
  *yn_out = yn;
  *ye_out = ye;
  const double D0 = S.eps_rel * fabs(yn) + S.eps_abs;
  double r = fabs(ye) / fabs(D0);
  if (!(r == r)) r = 0.0;  // GSL_MAX_DBL ignores NaN
  return r;
}

}  // namespace rtrg
