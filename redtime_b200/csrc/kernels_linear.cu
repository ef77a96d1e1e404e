// Kernel family (4): linear-theory tables and look-ups, all on the device.
//
//   k_beta_reduce   Beta_P(a,k) table pre-reduced in k   (hdr:513-637, tab:262-328)
//   k_growth_ode    scale-dependent growth tables: one thread integrates one wavenumber
//                   through n_lna+1 legs with RK8PD + GSL step control (hdr:133-190,690-710)
//   k_growth_norm / k_growth_rows   Dnorm and per-grid-k rows (hdr:712-729)
//   k_tgrid         CAMB transfer function at the grid wavenumbers (hdr:790-832)
//   k_qag           sigma_8 normalisation and sigma_v^2 by QAG-61: one warp per cosmology,
//                   lanes evaluate the 61 (122) abscissae, lane 0 runs QUADPACK's bisection
//                   bookkeeping in shared memory (hdr:846-879, 932-963)
//   k_init_state    initial conditions and the z1l linear spectrum (rt:1570-1586, 1299-1306)
//   k_hook_*        stage-level parity hooks (D_dD, Beta_P, Plin*)
#include <utility>

#include "gk61_table.h"
#include "rtrg_device.h"

namespace rtrg {

__constant__ PDTableau c_pd;
__constant__ double c_xgk[31], c_wgk[31], c_wg[15];

int linear_upload_constants() {
  const PDTableau pd = make_pd_tableau();
  cudaError_t e = cudaMemcpyToSymbol(c_pd, &pd, sizeof(pd));
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyToSymbol(c_xgk, gk61_xgk, sizeof(gk61_xgk));
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyToSymbol(c_wgk, gk61_wgk, sizeof(gk61_wgk));
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpyToSymbol(c_wg, gk61_wg, sizeof(gk61_wg));
  return (int)e;
}

// wavenumber of slot kk of the reduced beta table: nk grid values, then the growth-table ones
__device__ __forceinline__ double slot_k(const Batch &S, const double *kgrid, int kk) {
  return kk < S.nk ? kgrid[kk] : exp(S.lnkg[kk - S.nk]);
}

// In-place transform of the raw z=0 transfer columns (once per upload):
//   k_T -> ln k;  Tc_T -> ln(T_cb / T_cb[0]) with T_cb = f_b T_b + f_c T_c  (hdr:804-823)
//   T_nu -> beta = f_nu T_nu / T_c where the tables came in raw                (hdr:556-623)
__global__ void k_prep_T0(Batch S, double *__restrict__ T0) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.B) return;
  const Cosmo &c = S.cosmo[b];
  const double f_b = c.Ob / (c.Om - c.On), f_c = 1.0 - f_b;
  T0[b] = f_b * S.in[c.offTb] + f_c * S.in[c.offLT];
}
__global__ void k_prep_inputs(Batch S, const double *__restrict__ T0) {
  const int b = blockIdx.y;
  const Cosmo &c = S.cosmo[b];
  const double f_b = c.Ob / (c.Om - c.On), f_c = 1.0 - f_b, t0 = T0[b];
  // debug build: every table of this cosmology lies inside the input pool
  RT_ASSERT(c.offT >= 0 && c.offT + c.nT <= S.n_in && c.offLT >= 0 && c.offLT + c.nT <= S.n_in && c.offTb >= 0 &&
            c.offTb + c.nT <= S.n_in);
  RT_ASSERT(c.n_z == 0 || (c.offA >= 0 && c.offA + c.n_z <= S.n_in && c.offKb >= 0 && c.offKb + c.n_kb <= S.n_in));
  RT_ASSERT(c.n_z == 0 || c.offRow1 < 0 || (c.offRow1 + c.n_kb <= S.n_in && c.offBred + (long long)c.n_z * S.nkk <= S.n_in));
  RT_ASSERT(c.n_z == 0 || c.offRow1 >= 0 || c.offB + (long long)c.n_z * c.n_kb <= S.n_in);
  RT_ASSERT(c.offTc < 0 || c.offTc + (long long)c.n_z * c.n_kb <= S.n_in);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.nT; i += gridDim.x * blockDim.x) {
    const double Ti = f_b * S.in[c.offTb + i] + f_c * S.in[c.offLT + i];
    S.in[c.offT + i] = log(S.in[c.offT + i]);
    S.in[c.offLT + i] = log(Ti / t0);
  }
  if (c.offTc >= 0) {  // raw upload from page-locked caller buffers: beta = f_nu T_nu / T_c
    const double fn = c.On / c.Om;
    const long long nB = (long long)c.n_z * c.n_kb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nB; i += (long long)gridDim.x * blockDim.x)
      S.in[c.offB + i] = fn * S.in[c.offB + i] / S.in[c.offTc + i];
  }
}
int launch_prep_inputs(const Batch &S, double *T0, int max_rows, cudaStream_t st, Profiler *prof) {
  RT_TIC(prof, PC_PREP_INPUTS, st);
  k_prep_T0<<<(S.B + 127) / 128, 128, 0, st>>>(S, T0);
  int bx = (max_rows + 255) / 256;
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  k_prep_inputs<<<dim3(bx, S.B), 256, 0, st>>>(S, T0);
  RT_TOC(prof, st);
  return 2;
}

__global__ void k_beta_reduce(Batch S, const double *__restrict__ kgrid) {
  const int b = blockIdx.y, kk = blockIdx.x * blockDim.x + threadIdx.x;
  if (kk >= S.nkk) return;
  const Cosmo &c = S.cosmo[b];
  double *out = S.bred + (long long)b * S.n_zmax * S.nkk + kk;
  if (c.n_z == 0 || c.On / c.Om < 1e-10) {
    for (int j = 0; j < S.n_zmax; j++) out[(long long)j * S.nkk] = 0.0;
    return;
  }
  if (c.offRow1 >= 0) {  // reduced upload: the host already applied the k stencils
    for (int j = 0; j < c.n_z; j++) out[(long long)j * S.nkk] = S.in[c.offBred + (long long)j * S.nkk + kk];
    for (int j = c.n_z; j < S.n_zmax; j++) out[(long long)j * S.nkk] = 0.0;
    return;
  }
  double k = slot_k(S, kgrid, kk);
  if (k < S.beta_kmin) k = S.beta_kmin;  // hdr:538-545
  if (k > S.beta_kmax) k = S.beta_kmax;
  const BetaTab t = beta_tab(S, c);
  const Stencil st = tab_stencil_y(t.k, t.n_kb, k);
  for (int j = 0; j < c.n_z; j++)
    out[(long long)j * S.nkk] = stencil_apply(st, t.beta + (long long)j * t.n_kb);
  for (int j = c.n_z; j < S.n_zmax; j++) out[(long long)j * S.nkk] = 0.0;
}

// Right-hand side of the growth ODE as the device integrates it: the arithmetic of growth_rhs
// (rtrg_math.h) with the a-direction look-up of beta made cheap -- the a nodes and the reciprocal
// Lagrange denominators of every interval sit in shared memory (one block = one cosmology), the
// interval of the previous call is tried first, and E(a) is one exp of p ln a + q (1 - a).
// Differences to growth_rhs are at the 1e-16 level per call (reciprocal products instead of
// chained divisions); the step sequence GSL takes is unchanged.
struct GrowthRhsDev {
  BgStatic bg;
  const double *brow;   // beta at this wavenumber for every a node, stride bstride
  long long bstride;
  const double *s_a;    // [n_z] a nodes (shared)
  const double *s_inv;  // [n_z][4] reciprocal denominators of the cubic on nodes n-1..n+2 (shared)
  int n_z;
  bool has_beta;
  mutable int n_last;
  // the table values around interval n_last, re-read from global memory only when the interval
  // changes (the 13 stage abscissae of an attempt almost always share it)
  mutable int n_cached;
  mutable double bc[4];
  __device__ __forceinline__ double beta(double a) const {
    if (!has_beta) return 0.0;
    if (a > 1.0) a = 1.0;
    int n = n_last;
    const int X = n_z;
    if (!((n == 0 || s_a[n] < a) && (s_a[n + 1] >= a || n == X - 2))) {
      n = tab_find(s_a, X, a);
      n_last = n;
    }
    const bool cubic = n > 0 && n < X - 2;
    if (n != n_cached) {
      n_cached = n;
      if (cubic) {
        bc[0] = brow[(n - 1) * bstride], bc[1] = brow[n * bstride];
        bc[2] = brow[(n + 1) * bstride], bc[3] = brow[(n + 2) * bstride];
      } else {
        bc[1] = brow[n * bstride], bc[2] = brow[(n + 1) * bstride];
      }
    }
    if (cubic) {
      const double d0 = a - s_a[n - 1], d1 = a - s_a[n], d2 = a - s_a[n + 1], d3 = a - s_a[n + 2];
      const double *iv = s_inv + 4 * n;
      return d1 * d2 * d3 * iv[0] * bc[0] + d0 * d2 * d3 * iv[1] * bc[1] + d0 * d1 * d3 * iv[2] * bc[2] +
             d0 * d1 * d2 * iv[3] * bc[3];
    }
    const double f0 = bc[1], f1 = bc[2];
    return f0 + (f1 - f0) * iv_lin(n) * (a - s_a[n]);
  }
  __device__ __forceinline__ double iv_lin(int n) const { return 1.0 / (s_a[n + 1] - s_a[n]); }
  // the part of the right-hand side that depends on a alone (hdr:151-160, 466-500)
  struct Coef {
    double F0, F1;
  };
  static __device__ __forceinline__ Coef coef(const BgStatic &s, double a) {
    const double a2 = a * a, a3 = a2 * a, a4 = a2 * a2, a5 = a4 * a, ainv = 1.0 / a;
    const double E = exp(-3.0 * (1.0 + s.w0 + s.wa) * log(a) - 3.0 * s.wa * (1.0 - a));
    const double dEda = 3.0 * E * (s.wa - (1.0 + s.w0 + s.wa) * ainv);
    const double Y = bgs_Y(s, a), dYda = bgs_dYda(s, a);
    const double H2 = (s.Om - s.On) * (1.0 + Y) / a3 + s.OL * E + s.Og / a4;
    const double dlnH = 0.5 * a / H2 *
                        (s.fc * s.Om * (-3.0 * (1.0 + Y) + a * dYda) / a4 + s.OL * dEda - 4.0 * s.Og / a5);
    Coef c;
    c.F0 = 1.5 * s.Om / (a5 * H2);
    c.F1 = (3.0 + dlnH) * ainv;
    return c;
  }
  __device__ __forceinline__ void apply(double a, const Coef &c, const double y[2], double f[2]) const {
    const double bt = (a < 1e-3) ? bg.fn : beta(a);
    f[0] = y[1];
    f[1] = -c.F1 * y[1] + c.F0 * (bg.fc + bt) * y[0];
  }
  __device__ __forceinline__ void operator()(double a, const double y[2], double f[2]) const {
    apply(a, coef(bg, a), y, f);
  }
};

// Prince-Dormand tableau as compile-time constants: the stage sums of growth_integrate_coop are
// unrolled per stage with the zero entries removed (the generic loop spends two thirds of its
// issue slots on tableau look-ups, tests and branches).  Same terms in the same order.
struct PDc {
  static constexpr double A[13][12] = {RT_PD_ROWS};
  static constexpr double B8[13] = {RT_PD_B8};
  static constexpr double B7[13] = {RT_PD_B7};
};
template <int S, int J>
__device__ __forceinline__ void pd_term(const double (*k)[2], double &a0, double &a1) {
  constexpr double a = PDc::A[S][J];
  if (a != 0.0) {
    a0 += a * k[J][0];
    a1 += a * k[J][1];
  }
}
template <int S, int... J>
__device__ __forceinline__ void pd_acc_impl(const double (*k)[2], double &a0, double &a1,
                                            std::integer_sequence<int, J...>) {
  (pd_term<S, J>(k, a0, a1), ...);
}
__device__ __forceinline__ void pd_acc(int s, const double (*k)[2], double &a0, double &a1) {
  switch (s) {
#define RT_PD_CASE(S) case S: pd_acc_impl<S>(k, a0, a1, std::make_integer_sequence<int, S>()); break;
    RT_PD_CASE(1) RT_PD_CASE(2) RT_PD_CASE(3) RT_PD_CASE(4) RT_PD_CASE(5) RT_PD_CASE(6)
    RT_PD_CASE(7) RT_PD_CASE(8) RT_PD_CASE(9) RT_PD_CASE(10) RT_PD_CASE(11)
    default: pd_acc_impl<12>(k, a0, a1, std::make_integer_sequence<int, 12>()); break;
#undef RT_PD_CASE
  }
}
template <int J>
__device__ __forceinline__ void pd_bterm(const double (*k)[2], double (&s8)[2], double (&s7)[2]) {
  constexpr double b8 = PDc::B8[J], b7 = PDc::B7[J];
  if (b8 != 0.0) {
    s8[0] += b8 * k[J][0];
    s8[1] += b8 * k[J][1];
  }
  if (b7 != 0.0) {
    s7[0] += b7 * k[J][0];
    s7[1] += b7 * k[J][1];
  }
}
template <int... J>
__device__ __forceinline__ void pd_bsum(const double (*k)[2], double (&s8)[2], double (&s7)[2],
                                        std::integer_sequence<int, J...>) {
  (pd_bterm<J>(k, s8, s7), ...);
}

// Warp-cooperative form of growth_integrate (rtrg_math.h; same arithmetic per lane, bit for bit).
// The lanes of a warp integrate the wavenumbers of ONE cosmology, and almost always in lock
// step: every leg restarts from h = 1e-6 a (hdr:170-190), grows by the capped factor 5 and ends
// on the clamped final step, so (t0, h0) of an attempt are the same in all lanes.  Then the 13
// stage abscissae are the same too, and F0(a), F1(a) -- a log, an exp and six divisions, ~85 % of
// the dependent chain of a stage -- are evaluated once per attempt by 13 lanes in parallel and
// passed through shared memory, instead of 13 times in sequence by every lane.  Attempts in
// which the lanes disagree (a rejection at one wavenumber) fall back to per-lane evaluation.
__device__ void growth_integrate_coop(const GrowthRhsDev &g, double a_begin, double a_end, double y[2],
                                      bool valid, double *s_cf) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const double *C = c_pd.C;
  double t = a_begin;
  const double t1 = a_end;
  double h = 1e-6 * t;
  int attempts = 0;
  bool alive = valid, retry = false;
  double t0 = t, dt = 0.0, h0 = h, y0[2] = {y[0], y[1]};
  double k[13][2];
  for (;;) {
    if (alive && !retry && !((t1 - t) * h > 0)) alive = false;
    const unsigned ball = __ballot_sync(FULL, alive);
    if (!ball) break;
    if (alive && !retry) {  // --- gsl_odeiv_evolve_apply starts
      t0 = t, dt = t1 - t0, h0 = h;
      y0[0] = y[0], y0[1] = y[1];
    }
    bool final_step = false;
    if (alive && ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt))) {
      h0 = dt;
      final_step = true;
    }
    const int leader = __ffs(ball) - 1;
    const double lt = __shfl_sync(FULL, t0, leader), lh = __shfl_sync(FULL, h0, leader);
    const bool coop = __all_sync(FULL, !alive || (t0 == lt && h0 == lh));
    if (coop) {
      if (lane < 13) {
        const GrowthRhsDev::Coef c = GrowthRhsDev::coef(g.bg, lane == 0 ? lt : lt + C[lane] * lh);
        s_cf[2 * lane] = c.F0;
        s_cf[2 * lane + 1] = c.F1;
      }
      __syncwarp();
    }
    if (alive) {
      if (!retry) {
        GrowthRhsDev::Coef c;
        if (coop) c.F0 = s_cf[0], c.F1 = s_cf[1];
        else c = GrowthRhsDev::coef(g.bg, t0);
        g.apply(t0, c, y0, k[0]);
      }
      for (int s = 1; s < 13; s++) {
        double acc0 = 0, acc1 = 0;
        pd_acc(s, k, acc0, acc1);
        const double yt[2] = {y0[0] + h0 * acc0, y0[1] + h0 * acc1};
        const double as = t0 + C[s] * h0;
        GrowthRhsDev::Coef c;
        if (coop) c.F0 = s_cf[2 * s], c.F1 = s_cf[2 * s + 1];
        else c = GrowthRhsDev::coef(g.bg, as);
        g.apply(as, c, yt, k[s]);
      }
      double s8[2] = {0, 0}, s7[2] = {0, 0};
      pd_bsum(k, s8, s7, std::make_integer_sequence<int, 13>());
      const double yn[2] = {y0[0] + h0 * s8[0], y0[1] + h0 * s8[1]};
      const double ye[2] = {h0 * (s7[0] - s8[0]), h0 * (s7[1] - s8[1])};
      attempts++;
      const double tn = final_step ? t1 : t0 + h0;
      double rmax = DBL_MIN;  // control_y_new(0, 1e-6), order 8
      for (int i = 0; i < 2; i++) {
        const double D0 = 1e-6 * fabs(yn[i]) + 0.0;
        const double r = fabs(ye[i]) / fabs(D0);
        if (r > rmax) rmax = r;
      }
      const double h_old = h0;
      const int adj = gsl_hadjust(rmax, 8, &h0);
      retry = false;
      if (adj == -1) {
        const double t_next = tn + h0;
        if (fabs(h0) < fabs(h_old) && t_next != tn) retry = true;  // reject, retry smaller
        else h0 = h_old;
      }
      if (!retry) {
        y[0] = yn[0];
        y[1] = yn[1];
        t = tn;
        h = h0;
        if (attempts > 2000000) alive = false;
      }
    }
    __syncwarp();  // s_cf is rewritten by the next attempt
  }
}

__global__ void __launch_bounds__(64, 7) k_growth_ode(Batch S) {  // 7 x 148 >= 1024 cosmologies: one wave
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  const Cosmo &c = S.cosmo[b];
  extern __shared__ double s_g[];  // [n_z] a nodes, [n_z][4] reciprocal denominators
  double *s_a = s_g, *s_inv = s_g + S.n_zmax;
  const bool has_beta = c.n_z > 0 && c.On / c.Om >= 1e-10;
  const double *an = S.in + c.offA;
  for (int n = threadIdx.x; n < c.n_z; n += blockDim.x) {
    s_a[n] = an[n];
    if (n > 0 && n < c.n_z - 2) {
      const double *p = an + n - 1;
      s_inv[4 * n + 0] = 1.0 / ((p[0] - p[1]) * (p[0] - p[2]) * (p[0] - p[3]));
      s_inv[4 * n + 1] = 1.0 / ((p[1] - p[0]) * (p[1] - p[2]) * (p[1] - p[3]));
      s_inv[4 * n + 2] = 1.0 / ((p[2] - p[0]) * (p[2] - p[1]) * (p[2] - p[3]));
      s_inv[4 * n + 3] = 1.0 / ((p[3] - p[0]) * (p[3] - p[1]) * (p[3] - p[2]));
    }
  }
  __syncthreads();
  const bool valid = j <= S.n_lnk;  // lanes beyond the last wavenumber still help with the coefficients
  double *s_cf = s_inv + 4 * S.n_zmax + (threadIdx.x >> 5) * 26;
  GrowthRhsDev g;
  g.bg = bg_static(c);
  g.brow = S.bred + (long long)b * S.n_zmax * S.nkk + (S.nk + (valid ? j : 0));
  g.bstride = S.nkk;
  g.s_a = s_a;
  g.s_inv = s_inv;
  g.n_z = c.n_z;
  g.has_beta = has_beta;
  g.n_last = 0;
  g.n_cached = -1;
  const int nj = S.n_lnk + 1;
  double *G = S.G + (long long)b * (S.n_lna + 1) * nj, *dD = S.dD + (long long)b * (S.n_lna + 1) * nj;
  double y[2] = {1.0, 1.0 / S.a_early};  // hdr:697-698
  growth_integrate_coop(g, S.a_early, GROWTH_A_MIN, y, valid, s_cf);
  if (valid) {
    G[j] = y[0] / GROWTH_A_MIN;
    dD[j] = y[1];
  }
  for (int i = 1; i <= S.n_lna; i++) {
    const double a1 = exp(S.lna[i]);
    growth_integrate_coop(g, exp(S.lna[i - 1]), a1, y, valid, s_cf);
    if (valid) {
      G[(long long)i * nj + j] = y[0] / a1;
      dD[(long long)i * nj + j] = y[1];
    }
  }
}

__global__ void k_growth_norm(Batch S) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > S.n_lnk) return;
  const GrowthTab t = growth_tab(S, b);
  // DnormTab[j] = G_lna_lnk(0, lnkTab[j])  (hdr:716-717)
  S.Dnorm[(long long)b * (S.n_lnk + 1) + j] =
      tab2d(t.lna, t.n_lna + 1, t.lnk, t.n_lnk + 1, t.G, 0.0, t.lnk[j]);
}

__global__ void k_growth_rows(Batch S, const double *__restrict__ kgrid) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nk) return;
  const GrowthTab t = growth_tab(S, b);
  double k = kgrid[i];
  if (k > GROWTH_K_MAX) k = GROWTH_K_MAX;
  if (k < GROWTH_K_MIN) k = GROWTH_K_MIN;
  const double lnk0 = log(k);
  S.D0row[(long long)b * S.nk + i] = tab1d(t.lnk, t.Dnorm, t.n_lnk + 1, lnk0);
  const Stencil st = tab_stencil_y(t.lnk, t.n_lnk + 1, lnk0);
  const int nj = S.n_lnk + 1;
  for (int ia = 0; ia <= S.n_lna; ia++) {
    const long long o = ((long long)b * (S.n_lna + 1) + ia) * S.nk + i;
    S.Grow[o] = stencil_apply(st, t.G + (long long)ia * nj);
    S.dDrow[o] = stencil_apply(st, t.dD + (long long)ia * nj);
  }
}

__global__ void k_tgrid(Batch S, const double *__restrict__ kgrid) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nk) return;
  const LinCtx L = lin_ctx(S, b);
  S.Tgrid[(long long)b * S.nk + i] = transfer_cb(L, kgrid[i]);
}

// ---- QAG-61, one warp per cosmology ---------------------------------------------------
template <int WHICH>
__device__ __forceinline__ double qag_integrand(const LinCtx &L, double x) {
  return WHICH == 0 ? sigma8_integrand(L, x) : sigmav_integrand(L, x);
}
template <int WHICH>
__device__ double qag_warp(const LinCtx &L, QagState &w, double *fv, double *iv, int *status) {
  const int lane = threadIdx.x & 31;
  const double x0 = -15, x1 = 15, epsabs = 0, epsrel = 1e-4;  // hdr:853-854, 942-943
  const int limit = 1000;
  for (int s = lane; s < 61; s += 32) fv[s] = qag_integrand<WHICH>(L, qk61_abscissa(c_xgk, x0, x1, s));
  __syncwarp();
  int done = 0;
  if (lane == 0) {
    const Qk61Out q0 = qk61_combine(c_xgk, c_wgk, c_wg, x0, x1, fv);
    done = qag_begin(w, x0, x1, epsabs, epsrel, q0) ? 1 : 0;
    if (!done) qag_next(w, &iv[0], &iv[1], &iv[2], &iv[3]);
  }
  done = __shfl_sync(0xffffffffu, done, 0);
  while (!done) {
    __syncwarp();
    for (int s = lane; s < 122; s += 32) {
      const int half = s >= 61;
      fv[s] = qag_integrand<WHICH>(L, qk61_abscissa(c_xgk, iv[2 * half], iv[2 * half + 1], s - 61 * half));
    }
    __syncwarp();
    if (lane == 0) {
      const Qk61Out q1 = qk61_combine(c_xgk, c_wgk, c_wg, iv[0], iv[1], fv);
      const Qk61Out q2 = qk61_combine(c_xgk, c_wgk, c_wg, iv[2], iv[3], fv + 61);
      done = qag_update(w, q1, q2, limit) ? 1 : 0;
      if (!done) qag_next(w, &iv[0], &iv[1], &iv[2], &iv[3]);
    }
    done = __shfl_sync(0xffffffffu, done, 0);
  }
  double r = 0;
  if (lane == 0) {
    r = qag_result(w);
    if (w.done == 3 || w.error_type) *status = RTRG_QAG_FAIL;
  }
  return __shfl_sync(0xffffffffu, r, 0);
}

__global__ void __launch_bounds__(32) k_qag(Batch S) {
  const int b = blockIdx.x;
  __shared__ QagState w;
  __shared__ double fv[122];
  __shared__ double iv[4];
  __shared__ int st;
  if (threadIdx.x == 0) st = 0;
  __syncwarp();
  Cosmo &c = S.cosmo[b];
  const LinCtx L = lin_ctx(S, b);
  const double r8 = qag_warp<0>(L, w, fv, iv, &st);
  if (threadIdx.x == 0) c.Norm = c.s8 * c.s8 / r8;  // hdr:874
  __threadfence_block();
  __syncwarp();
  const double rv = qag_warp<1>(L, w, fv, iv, &st);
  if (threadIdx.x == 0) {
    c.sigv2_0 = rv / (6.0 * M_PI * M_PI);  // hdr:961
    if (st) c.status = st;
  }
}

// ---- per grid-wavenumber linear quantities through the pre-reduced rows ----------------
struct RowLin {
  double D, dD, beta;
  bool ok;
};
__device__ __forceinline__ double row_beta(const Batch &S, int b, int i, double a) {
  const Cosmo &c = S.cosmo[b];
  const BetaTab t = beta_tab(S, c);
  return beta_row(t, S.bred + (long long)b * S.n_zmax * S.nkk + i, a, S.nkk);
}
__device__ __forceinline__ bool row_D_dD(const Batch &S, int b, int i, double z, double *D, double *dD) {
  const long long o = (long long)b * (S.n_lna + 1) * S.nk + i;
  return growth_D_dD_row(S.lna, S.n_lna, S.Grow + o, S.dDrow + o, S.nk, S.D0row[(long long)b * S.nk + i], z,
                         D, dD);
}
// Plin / Plin_cb / Plin_nu at a grid wavenumber (hdr:881-930)
__device__ __forceinline__ double row_plin(const Batch &S, int b, int i, double k, double z, int which,
                                           double *D_out, double *dD_out) {
  const Cosmo &c = S.cosmo[b];
  const double a = 1.0 / (1.0 + z), fn = c.On / c.Om, fc = 1.0 - fn;
  const double T = S.Tgrid[(long long)b * S.nk + i];
  const double B = row_beta(S, b, i, a);
  const double F = 1.0 - fn + B;
  double D = NAN, dD = NAN;
  row_D_dD(S, b, i, z, &D, &dD);
  if (D_out) *D_out = D;
  if (dD_out) *dD_out = dD;
  const double P = c.Norm * pow(k, c.ns) * T * T * F * F * D * D;
  if (which == 0) return P;
  if (which == 1) {
    if (fn <= 1e-10) return P;
    const double Rr = 1.0 / (fc + B);
    return P * Rr * Rr;
  }
  if (fn <= 1e-10) return 0.0;
  const double Rr = B / fn / (fc + B);
  return P * Rr * Rr;
}

__global__ void k_init_state(Batch S, const double *__restrict__ kgrid) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S.nk) return;
  const Cosmo &c = S.cosmo[b];
  const int nk = S.nk;
  const double k = kgrid[i];
  double D = 0, dD = 0;
  const double Pin = row_plin(S, b, i, k, c.z_in, 1, &D, &dD);
  const double f_in = c.a_in * dD / D;  // rt:1576
  double *y = S.y + (long long)b * N_U * nk;
  y[i] = log(Pin);
  y[nk + i] = log(Pin * f_in);
  y[2 * nk + i] = log(Pin * f_in * f_in);
  for (int j = N_UP; j < N_U; j++) y[j * nk + i] = 0.0;
  // linear spectrum at z1l for the 1-loop cache (rt:1299-1306)
  double Dz = 0;
  const double Pz = row_plin(S, b, i, k, S.z1l, 1, &Dz, nullptr);
  S.D_z1l[(long long)b * nk + i] = Dz;
  const double l = log(Pz);
  double *yz = S.y_z1l + (long long)b * 3 * nk;
  yz[i] = l;
  yz[nk + i] = l;
  yz[2 * nk + i] = l;
}

// ---- hooks -------------------------------------------------------------------------------
__global__ void k_hook_DdD(Batch S, int b, double z, const double *k, int n, double *D, double *dD,
                           int *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GrowthTab t = growth_tab(S, b);
  if (!growth_D_dD(t, z, k[i], &D[i], &dD[i])) *err = 1;
}
__global__ void k_hook_beta(Batch S, int b, double a, const double *k, int n, double *beta, int *err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const BetaTab t = beta_tab(S, S.cosmo[b]);
  const double v = beta_P(t, a, k[i]);
  beta[i] = v;
  if (v != v) *err = 1;
}
__global__ void k_hook_plin(Batch S, int b, int which, double z, const double *k, int n, double *P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const LinCtx L = lin_ctx(S, b);
  P[i] = which == 0 ? plin(L, z, k[i]) : which == 1 ? plin_cb(L, z, k[i]) : plin_nu(L, z, k[i]);
}

// ---- launch sequence of the device-side initialisation ----------------------------------
int launch_linear_init(const Batch &S, const double *kgrid, cudaStream_t st, Profiler *prof) {
  int n = 0;
  const int B = S.B;
  RT_TIC(prof, PC_BETA_REDUCE, st);
  k_beta_reduce<<<dim3((S.nkk + 127) / 128, B), 128, 0, st>>>(S, kgrid), n++;
  RT_TOC(prof, st);
  RT_TIC(prof, PC_GROWTH_ODE, st);
  k_growth_ode<<<dim3((S.n_lnk + 1 + 63) / 64, B), 64, ((size_t)5 * S.n_zmax + 2 * 26) * sizeof(double), st>>>(S), n++;
  RT_TOC(prof, st);
  RT_TIC(prof, PC_GROWTH_TABS, st);
  k_growth_norm<<<dim3((S.n_lnk + 1 + 63) / 64, B), 64, 0, st>>>(S), n++;
  k_growth_rows<<<dim3((S.nk + 127) / 128, B), 128, 0, st>>>(S, kgrid), n++;
  k_tgrid<<<dim3((S.nk + 127) / 128, B), 128, 0, st>>>(S, kgrid), n++;
  RT_TOC(prof, st);
  RT_TIC(prof, PC_QAG, st);
  k_qag<<<B, 32, 0, st>>>(S), n++;
  RT_TOC(prof, st);
  RT_TIC(prof, PC_INIT_STATE, st);
  k_init_state<<<dim3((S.nk + 127) / 128, B), 128, 0, st>>>(S, kgrid), n++;
  RT_TOC(prof, st);
  return n;
}

void launch_hook_DdD(const Batch &S, int b, double z, const double *k, int n, double *D, double *dD,
                     int *err, cudaStream_t st) {
  k_hook_DdD<<<(n + 127) / 128, 128, 0, st>>>(S, b, z, k, n, D, dD, err);
}
void launch_hook_beta(const Batch &S, int b, double a, const double *k, int n, double *beta, int *err,
                      cudaStream_t st) {
  k_hook_beta<<<(n + 127) / 128, 128, 0, st>>>(S, b, a, k, n, beta, err);
}
void launch_hook_plin(const Batch &S, int b, int which, double z, const double *k, int n, double *P,
                      cudaStream_t st) {
  k_hook_plin<<<(n + 127) / 128, 128, 0, st>>>(S, b, which, z, k, n, P);
}

}  // namespace rtrg
