// Kernel families (2) and (3): the Time-RG right-hand side and the batched, device-resident
// adaptive RKF45 stepper (replaces redTime.cc:1383-1547 and the gsl_odeiv driver loop of
// redTime.cc:1588-1632), plus the output assembly (redTime.cc:1634-1742).
//
// One integrator per cosmology.  All cosmologies advance in lock step over *attempt
// index*: every round each unfinished cosmology either attempts one RKF45 step with its
// own step size, or emits one output table.  The per-cosmology accept / reject / clip
// decisions follow gsl_odeiv_evolve_apply + std_control_hadjust exactly;
// they are taken on the device by k_ctrl_begin / k_ctrl_end, so no host round trip is
// needed inside a step.  State vectors stay in HBM in the reference's component-major
// layout; the error norm is a max over all 41 nk components (order independent).
//
//   k_rhs        dy = f(eta, y)                     HBM bound: 960 nk bytes per evaluation
//   k_combine    ytmp = y + h sum_j a_sj k_j
//   k_final      5th-order solution, error estimate, rmax = max |yerr|/(eps_rel |y| + eps_abs)
//   k_accept     y <- ynew for accepted cosmologies
//   k_output     one nk x ncols table
#include "rtrg_device.h"
#include "stage_device.h"

namespace rtrg {

// SPLIT = false: one thread per row does the four pieces of the row one after the other (batches:
// fewest redundant coefficient evaluations).  SPLIT = true: a thread per (row, piece), 32 rows per
// block -- a quarter of the serial work per thread, for launches too small to fill the GPU (one
// cosmology, k-sharded ranks).  Same arithmetic per component either way (rhs_row, stage_device.h).
template <bool SPLIT>
__global__ void __launch_bounds__(128)
    k_rhs(Batch S, const double *__restrict__ kgrid, const double *__restrict__ yv,
          double *__restrict__ dyv, int stage, const int *__restrict__ mask) {
  const int b = blockIdx.y;
  if (mask && !mask[b]) return;
  const Cosmo &c = S.cosmo[b];
  const int nk = S.nk;
  __shared__ RhsShared sh;
  rhs_time_setup_block(S, c, (stage < 0) ? S.t[b] : S.t[b] + RKF45::c(stage) * S.h_try[b], sh);
  const int piece = SPLIT ? (int)(threadIdx.x >> 5) : -1;
  const int i = S.k_lo + (SPLIT ? blockIdx.x * 32 + (threadIdx.x & 31) : blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= S.k_hi) return;
  const int one_loop = c.sw_nl && c.sw_1l;
  const double *s1 = (one_loop ? S.src_z1l : S.src) + (long long)b * N_SRC * nk + i;
  rhs_row(S, c, b, i, piece, sh, kgrid[i], s1, nk, yv + (long long)b * N_U * nk + i, dyv + (long long)b * N_U * nk + i);
}

// ---------------------------------------------------------------------------- k_attempt_local
// One whole RKF45 attempt of the cosmologies whose right-hand side needs no new integrals (1-loop
// and linear mode): there the 41 equations of a row split into four independent systems -- ln P
// with I (17), and the three multipoles of Q (8 each) -- and no row talks to another one, so
// the six stages, the combinations between them, the 5th-order solution and the error norm are
// done by one thread per (row, system) with the stage derivatives kept in shared memory.  Same
// formulas, in the same order, as k_rhs / k_combine / k_final, which remain the path of the
// full Time-RG cosmologies (their stages are separated by integral evaluations).  HBM traffic:
// y and the sources in, ynew out = 960 B per row instead of ~17.6 KB through the kernel chain.
enum { ATT_ROWS = 32 };
template <int NC, int STAGE>
__device__ __forceinline__ void att_combine_s(const double *s_k, int lane, double h, const double *y, double *yt) {
#pragma unroll
  for (int j = 0; j < NC; j++) {
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < STAGE; m++) {
      const double a = RKF45::a(STAGE, m);
      if (a != 0.0) acc += a * s_k[(m * NC + j) * ATT_ROWS + lane];
    }
    yt[j] = y[j] + h * acc;
  }
}
// ytmp of stage `stage` (the tableau row folds to constants in each case)
template <int NC>
__device__ __forceinline__ void att_combine(const double *s_k, int stage, int lane, double h, const double *y,
                                            double *yt) {
  switch (stage) {
    case 1: att_combine_s<NC, 1>(s_k, lane, h, y, yt); break;
    case 2: att_combine_s<NC, 2>(s_k, lane, h, y, yt); break;
    case 3: att_combine_s<NC, 3>(s_k, lane, h, y, yt); break;
    case 4: att_combine_s<NC, 4>(s_k, lane, h, y, yt); break;
    default: att_combine_s<NC, 5>(s_k, lane, h, y, yt); break;
  }
}
template <int NC>
__device__ __forceinline__ double att_final(const Batch &S, const double *s_k, int lane, double h, const double *y,
                                            double *yn_out, long long stride) {
  double r = 0.0;
#pragma unroll
  for (int j = 0; j < NC; j++) {
    double acc = 0.0, err = 0.0;
#pragma unroll
    for (int m = 0; m < RK_STAGES; m++) {
      const double km = s_k[(m * NC + j) * ATT_ROWS + lane];
      if (RKF45::b(m) != 0.0) acc += RKF45::b(m) * km;
      if (RKF45::e(m) != 0.0) err += RKF45::e(m) * km;
    }
    const double yn = y[j] + h * acc, ye = h * err;
    yn_out[(long long)j * stride] = yn;
    const double D0 = S.eps_rel * fabs(yn) + S.eps_abs;
    double rj = fabs(ye) / fabs(D0);
    if (!(rj == rj)) rj = 0.0;  // GSL_MAX_DBL ignores NaN
    r = fmax(r, rj);
  }
  return r;
}
static_assert(sizeof(RhsShared) <= 48 * sizeof(double), "Batch::att_time slots are 48 doubles");
// time-only quantities of the six stages, once per cosmology instead of once per 32-row block
__global__ void k_attempt_setup(Batch S, const int *__restrict__ mask) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = idx / RK_STAGES, st = idx - b * RK_STAGES;
  if (b >= S.B || !mask[b]) return;
  RhsShared sh;
  rhs_time_setup(S, S.cosmo[b], st == 0 ? S.t[b] : S.t[b] + RKF45::c(st) * S.h_try[b], sh);
  reinterpret_cast<RhsShared *>(reinterpret_cast<double *>(S.att_time) + ((size_t)b * RK_STAGES + st) * 48)[0] = sh;
}
__global__ void __launch_bounds__(128)
    k_attempt_local(Batch S, const double *__restrict__ kgrid, const int *__restrict__ mask) {
  const int b = blockIdx.y;
  if (!mask[b]) return;
  const Cosmo &c = S.cosmo[b];
  const int nk = S.nk, lane = threadIdx.x & 31, piece = threadIdx.x >> 5;
  __shared__ RhsShared sh[RK_STAGES];
  __shared__ double s_om10[RK_STAGES][ATT_ROWS], s_pre[RK_STAGES][ATT_ROWS], s_fz[RK_STAGES][ATT_ROWS];
  __shared__ double s_r[4];
  extern __shared__ double s_kall[];  // PI: [6][17][32]; Q_l: [6][8][32] each
  const double h = S.h_try[b];
  {  // stage quantities prepared by k_attempt_setup
    const double *src = reinterpret_cast<const double *>(S.att_time) + (size_t)b * RK_STAGES * 48;
    constexpr int ND = sizeof(RhsShared) / sizeof(double);
    static_assert(sizeof(RhsShared) % sizeof(double) == 0, "copied as doubles");
    for (int w = threadIdx.x; w < RK_STAGES * ND; w += blockDim.x)
      reinterpret_cast<double *>(&sh[w / ND])[w % ND] = src[(w / ND) * 48 + w % ND];
  }
  __syncthreads();
  const int i0 = S.k_lo + blockIdx.x * ATT_ROWS;
  for (int it = threadIdx.x; it < RK_STAGES * ATT_ROWS; it += blockDim.x) {
    const int st = it / ATT_ROWS, r = it - st * ATT_ROWS;
    if (i0 + r < S.k_hi) rhs_row_coeffs(S, c, b, i0 + r, sh[st], &s_om10[st][r], &s_pre[st][r], &s_fz[st][r]);
  }
  __syncthreads();
  const int i = i0 + lane;
  const int one_loop = c.sw_nl && c.sw_1l;
  const int evolve_Q = (S.print_Q || c.sw_pr);
  double r = 0.0;
  if (i < S.k_hi) {
    const double k = kgrid[i];
    const double *s1 = (one_loop ? S.src_z1l : S.src) + (long long)b * N_SRC * nk + i;
    const double *yb = S.y + (long long)b * N_U * nk + i;
    double *yn = S.ynew + (long long)b * N_U * nk + i;
    if (piece == 0) {
      constexpr int NC = N_UP + N_UI;
      double *s_k = s_kall;
      double y[NC], A0[N_UI];
#pragma unroll
      for (int j = 0; j < NC; j++) y[j] = yb[(long long)j * nk];
#pragma unroll
      for (int j = 0; j < N_UI; j++) A0[j] = s1[(long long)j * nk];
      for (int st = 0; st < RK_STAGES; st++) {
        double yt[NC], dy[NC], A14[N_UI];
        if (st == 0) {
#pragma unroll
          for (int j = 0; j < NC; j++) yt[j] = y[j];
        } else {
          att_combine<NC>(s_k, st, lane, h, y, yt);
        }
        const double pre = s_pre[st][lane], fz = s_fz[st][lane];
        double fp[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
        if (one_loop) {
#pragma unroll
          for (int p = 1; p < 5; p++) fp[p] = fp[p - 1] * fz;
        }
#pragma unroll
        for (int j = 0; j < N_UI; j++)
          A14[j] = !c.sw_nl ? 0.0 : one_loop ? pre * fp[a14_fpow(j)] * A0[j] : A0[j];
        trg_rhs_PI(sh[st].eeta, k, s_om10[st][lane], sh[st].Om11, c.sw_nl, yt, A14, dy);
#pragma unroll
        for (int j = 0; j < NC; j++) s_k[(st * NC + j) * ATT_ROWS + lane] = dy[j];
      }
      r = att_final<NC>(S, s_k, lane, h, y, yn, nk);
    } else {
      constexpr int NC = 8;
      const int l = piece - 1, j0 = N_UP + N_UI + 8 * l;
      double *s_k = s_kall + RK_STAGES * (N_UP + N_UI) * ATT_ROWS + l * RK_STAGES * NC * ATT_ROWS;
      double y[NC], R0[NC];
      const bool on = c.sw_nl && evolve_Q;
#pragma unroll
      for (int j = 0; j < NC; j++) y[j] = yb[(long long)(j0 + j) * nk];
#pragma unroll
      for (int j = 0; j < NC; j++) R0[j] = on ? s1[(long long)(N_UI + 8 * l + j) * nk] : 0.0;
      for (int st = 0; st < RK_STAGES; st++) {
        double yt[NC], dQ[NC], R[NC];
        if (st == 0) {
#pragma unroll
          for (int j = 0; j < NC; j++) yt[j] = y[j];
        } else {
          att_combine<NC>(s_k, st, lane, h, y, yt);
        }
        if (on) {
          const double pre = s_pre[st][lane], fz = s_fz[st][lane];
          double fp[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
          if (one_loop) {
#pragma unroll
            for (int p = 1; p < 5; p++) fp[p] = fp[p - 1] * fz;
          }
#pragma unroll
          for (int j = 0; j < NC; j++) R[j] = one_loop ? pre * fp[r24_fpow(8 * l + j)] * R0[j] : R0[j];
          trg_rhs_Q(sh[st].eeta, s_om10[st][lane], sh[st].Om11, yt, R, dQ);
        } else {
#pragma unroll
          for (int j = 0; j < NC; j++) dQ[j] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < NC; j++) s_k[(st * NC + j) * ATT_ROWS + lane] = dQ[j];
      }
      r = att_final<NC>(S, s_k, lane, h, y, yn + (long long)j0 * nk, nk);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
  if (lane == 0) s_r[piece] = r;
  __syncthreads();
  if (threadIdx.x == 0) {
    r = fmax(fmax(s_r[0], s_r[1]), fmax(s_r[2], s_r[3]));
    atomicMax(&S.rmax_bits[b], (unsigned long long)__double_as_longlong(r));
  }
}

// ---------------------------------------------------------------------------- k_combine
__global__ void k_combine(Batch S, int stage, const int *__restrict__ mask) {
  const int b = blockIdx.y;
  if (mask && !mask[b]) return;
  const long long n = (long long)N_U * S.nk;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const long long o = (long long)b * n + idx, ks = (long long)S.B * n;
  S.ytmp[o] = rk_combine(stage, S.y[o], S.h_try[b], S.kst + o, ks);
}

// ---------------------------------------------------------------------------- k_final
__global__ void __launch_bounds__(256) k_final(Batch S, const int *__restrict__ mask) {
  const int b = blockIdx.y;
  if (mask && !mask[b]) return;
  const long long n = (long long)N_U * S.nk;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double r = 0.0;
  if (idx < n) {
    const long long o = (long long)b * n + idx, ks = (long long)S.B * n;
    double yn, ye;
    const double rj = rk_final(S, S.y[o], S.h_try[b], S.kst + o, ks, &yn, &ye);
    S.ynew[o] = yn;
    S.yerr[o] = ye;
    // only rows owned by this rank contribute (k-sharding); all rows otherwise
    const int i = (int)(idx % S.nk);
    if (i >= S.k_lo && i < S.k_hi) r = rj;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
  __shared__ double s_r[8];
  if ((threadIdx.x & 31) == 0) s_r[threadIdx.x >> 5] = r;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = fmax(r, s_r[w]);
    atomicMax(&S.rmax_bits[b], (unsigned long long)__double_as_longlong(r));
  }
}

// ---------------------------------------------------------------------------- k_stage_post
// Everything between the integrals of one Runge-Kutta stage and the exchange / integrals of the next,
// for launches too small to fill the GPU (one cosmology, k-sharded ranks), where the kernel chain
// k_assemble -> k_rhs -> k_combine (or k_final) is three launches of a few dependent memory round trips
// each: per CTA of STG_ROWS rows of one cosmology
//   1. the sources of the rows from the raw integrals (k_assemble's code),
//   2. the right-hand side of stage `stage` into kst[stage] (k_rhs's code, sources read from shared memory),
//   3. next in 1..5: ytmp of stage `next`;  next = 6: 5th-order solution, error estimate and error norm;
//      next = 0: nothing (the dydt_in evaluation after an accepted step).
// The same device functions in the same order as the separate kernels: same bits.
enum { STG_ROWS = 4, STG_THREADS = 512 };
__global__ void __launch_bounds__(STG_THREADS)
    k_stage_post(IntegralTabs tb, Batch S, const double *__restrict__ kgrid, const double *__restrict__ yv, int stage,
                 int next, const int *__restrict__ mask, int groups) {
  const int b = blockIdx.y;
  if (mask && !mask[b]) return;
  const Cosmo &c = S.cosmo[b];
  const int nk = S.nk, tid = threadIdx.x;
  __shared__ AsmShared<STG_ROWS> sa;
  __shared__ double s_src[N_SRC][STG_ROWS + 1];
  __shared__ RhsShared sh;
  __shared__ double s_r[STG_THREADS / 32];
  const int r0 = S.k_lo + blockIdx.x * STG_ROWS;
  const int rows = min((int)STG_ROWS, S.k_hi - r0);
  asm_load_table(tb, sa);
  asm_gather_vals(tb, c.sw_pr, S.Jpart, S.PZb, S.P3, S.Jlo, nullptr, b, r0, rows, groups, sa);
  // (ends with a barrier: the raw values are in place as well)
  rhs_time_setup_block(S, c, (stage < 0) ? S.t[b] : S.t[b] + RKF45::c(stage) * S.h_try[b], sh);
  for (int idx = tid; idx < N_SRC * STG_ROWS; idx += STG_THREADS) {
    const int o = idx / STG_ROWS, rr = idx - o * STG_ROWS;
    if (rr >= rows || !asm_wanted(o, groups)) continue;  // (the right-hand side reads requested groups only)
    const double v = asm_source(tb, sa, o, rr, r0 + rr);
    s_src[o][rr] = v;
    S.src[((long long)b * N_SRC + o) * nk + r0 + rr] = v;
  }
  __syncthreads();
  const long long n = (long long)N_U * nk, ks = (long long)S.B * n;
  double *dyv = S.kst + (size_t)(stage < 0 ? 0 : stage) * ks;
  if (tid < 4 * STG_ROWS) {
    const int piece = tid / STG_ROWS, rr = tid - piece * STG_ROWS;
    if (rr < rows) {
      const int i = r0 + rr;
      const int one_loop = c.sw_nl && c.sw_1l;
      const double *s1 = one_loop ? S.src_z1l + (long long)b * N_SRC * nk + i : &s_src[0][rr];
      rhs_row(S, c, b, i, piece, sh, kgrid[i], s1, one_loop ? (long long)nk : (long long)(STG_ROWS + 1),
              yv + (long long)b * n + i, dyv + (long long)b * n + i);
    }
  }
  if (next == 0) return;
  __syncthreads();  // (the derivatives just written are read back by other threads of the CTA)
  double r = 0.0;
  for (int idx = tid; idx < N_U * STG_ROWS; idx += STG_THREADS) {
    const int j = idx / STG_ROWS, rr = idx - j * STG_ROWS;
    if (rr >= rows) continue;
    const long long o = (long long)b * n + (long long)j * nk + r0 + rr;
    if (next < RK_STAGES) {
      S.ytmp[o] = rk_combine(next, S.y[o], S.h_try[b], S.kst + o, ks);
    } else {
      double yn, ye;
      r = fmax(r, rk_final(S, S.y[o], S.h_try[b], S.kst + o, ks, &yn, &ye));
      S.ynew[o] = yn;
      S.yerr[o] = ye;
    }
  }
  if (next >= RK_STAGES) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_xor_sync(0xffffffffu, r, o));
    if ((tid & 31) == 0) s_r[tid >> 5] = r;
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < STG_THREADS / 32; w++) r = fmax(r, s_r[w]);
      atomicMax(&S.rmax_bits[b], (unsigned long long)__double_as_longlong(r));
    }
  }
}

// ---------------------------------------------------------------------------- control
// has_full: some cosmology runs the full Time-RG (integrals inside the RHS).
__global__ void k_ctrl_begin(Batch S) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.B) return;
  S.flag_out[b] = S.flag_step[b] = S.flag_acc[b] = 0;
  S.m_full_step[b] = S.m_full_acc[b] = S.m_out_int[b] = S.m_loc_step[b] = 0;
  if (S.done[b]) return;
  const Cosmo &c = S.cosmo[b];
  RT_ASSERT(S.i_out[b] >= 0 && S.i_out[b] < MAX_OUT && S.i_out[b] < c.n_out);
  const double target = S.etaout[(long long)b * MAX_OUT + S.i_out[b]];
  const double t = S.t[b], h = S.h[b];
  if ((target - t) * h > 0) {  // rt:1614
    double h0 = h;
    const double dt = target - t;
    int fin = 0;
    if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) {
      h0 = dt;
      fin = 1;
    }
    S.h_try[b] = h0;
    S.final_step[b] = fin;
    S.flag_step[b] = 1;
    S.m_full_step[b] = (c.sw_nl && !c.sw_1l);
    S.m_loc_step[b] = !S.m_full_step[b];
    if (S.m_full_step[b]) S.counters[4LL * b + 3] += 5;  // integral evaluations of stages 2..6
    S.rmax_bits[b] = (unsigned long long)__double_as_longlong(DBL_MIN);
  } else {
    S.flag_out[b] = 1;
    S.m_out_int[b] = (c.sw_nl && c.sw_1l);  // rt:1646
    if (S.m_out_int[b]) S.counters[4LL * b + 3] += 1;
  }
}

__global__ void k_ctrl_end(Batch S, int max_attempts) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.B) return;
  if (S.done[b]) return;
  Cosmo &c = S.cosmo[b];
  long long *cnt = S.counters + 4LL * b;
  if (S.flag_out[b]) {
    S.i_out[b]++;
    if (S.i_out[b] >= c.n_out) {
      S.done[b] = 1;
      atomicSub(S.n_active, 1);
    }
    return;
  }
  if (!S.flag_step[b]) return;
  const double rmax = __longlong_as_double((long long)S.rmax_bits[b]);
  const double target = S.etaout[(long long)b * MAX_OUT + S.i_out[b]];
  const double h_old = S.h_try[b];
  const double t_new = S.final_step[b] ? target : S.t[b] + h_old;
  if (cnt[0] < RMAX_HIST) S.rmax_hist[(long long)b * RMAX_HIST + cnt[0]] = rmax;
  cnt[0]++;
  cnt[2] += 5;
  double h0 = h_old;
  const int adj = gsl_hadjust(rmax, 5, &h0);
  if (adj == -1) {
    const double t_next = t_new + h0;
    if (fabs(h0) < fabs(h_old) && t_next != t_new) {
      S.h[b] = h0;  // rejected: retry from the same (t, y) with the smaller step
      cnt[1]++;
      if (cnt[0] >= max_attempts) {
        c.status = RTRG_ODE_FAIL;
        S.done[b] = 1;
        atomicSub(S.n_active, 1);
      }
      return;
    }
    h0 = h_old;
  }
  S.t[b] = t_new;
  S.h[b] = h0;
  S.flag_acc[b] = 1;
  S.m_full_acc[b] = (c.sw_nl && !c.sw_1l);
  if (S.m_full_acc[b]) cnt[3] += 1;
  cnt[2] += 1;
  if (cnt[0] >= max_attempts) {
    c.status = RTRG_ODE_FAIL;
    S.done[b] = 1;
    atomicSub(S.n_active, 1);
  }
}

__global__ void k_accept(Batch S) {
  const int b = blockIdx.y;
  if (!S.flag_acc[b]) return;
  const long long n = (long long)N_U * S.nk;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  S.y[(long long)b * n + idx] = S.ynew[(long long)b * n + idx];
}

// ---------------------------------------------------------------------------- k_stash
// Round loop, cosmologies that reached an output redshift: keep the state for the deferred
// output stage (the evolution itself never consumes the output integrals, rt:1646-1653).
__global__ void k_stash(Batch S) {
  const int b = blockIdx.y;
  if (!S.flag_out[b]) return;
  const long long n = (long long)N_U * S.nk;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = S.vbase[b] + S.i_out[b];
  RT_ASSERT(v >= 0 && v < S.NO && S.i_out[b] < S.cosmo[b].n_out);
  if (idx == 0) {
    S.vc_have[v] = 1;
    S.t_stash[v] = S.t[b];
  }
  if (idx < n) S.ystash[(long long)v * n + idx] = S.y[(long long)b * n + idx];
}
// after the loop: which virtual cosmologies need the 1-loop output integrals (rt:1646), and the
// per-virtual copy of the cosmology record the integral kernels index
__global__ void k_vprep(Batch S) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= S.NO) return;
  const Cosmo c = S.cosmo[S.vc_b[v]];
  S.cosmo_v[v] = c;
  S.vc_mask[v] = S.vc_have[v] && c.sw_nl && c.sw_1l;
}

// ---------------------------------------------------------------------------- k_output
// columns of one row as in rt:1670-1737, for the virtual cosmologies [v0, v0 + gridDim.y);
// src_v holds the output integrals of that chunk
__global__ void __launch_bounds__(128) k_output(Batch S, const double *__restrict__ kgrid, int v0) {
  const int v = v0 + blockIdx.y;
  if (!S.vc_have[v]) return;
  const int b = S.vc_b[v], io = S.vc_io[v];
  const Cosmo &c = S.cosmo[b];
  const int nk = S.nk;
  const double z = S.zout[(long long)b * MAX_OUT + io], a = S.aout[(long long)b * MAX_OUT + io];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    double *h = S.hdr + ((long long)b * MAX_OUT + io) * 5;
    const GrowthTab gt = growth_tab(S, b);
    double D = NAN, dD = NAN;
    growth_D_dD(gt, z, 1e-3, &D, &dD);  // hdr:964-965
    h[0] = S.t_stash[v];
    h[1] = a;
    h[2] = z;
    h[3] = sqrt(bg_H2(c, a)) * H0H;
    h[4] = D * D * c.sigv2_0;
    if (io == 0) {
      double D0 = NAN;
      growth_D_dD(gt, 0.0, 1e-3, &D0, &dD);
      S.hdr0[2 * b] = log(1.0 / c.a_in);  // rt:1598
      S.hdr0[2 * b + 1] = D0 * D0 * c.sigv2_0;
    }
  }
  if (i >= nk || i < S.k_lo || i >= S.k_hi) return;
  const double k = kgrid[i];
  const int ncols = S.ncols[b];
  double *row = S.out + S.out_off[b] + ((long long)io * nk + i) * ncols;
  const double *y = S.ystash + (long long)v * N_U * nk + i;
  const double a_ain = a / c.a_in, a2 = a_ain * a_ain, a3 = a2 * a_ain, a4 = a2 * a2;
  int col = 0;
  row[col++] = k;
  if (c.sw_pl) {
    double D = NAN, dD = NAN;
    const double Pcb = [&] {
      // Plin_cb and D,dD share the look-ups
      const double fn = c.On / c.Om, fc = 1.0 - fn;
      const double T = S.Tgrid[(long long)b * nk + i];
      const double B = ode_row_beta(S, b, i, a);
      const double F = 1.0 - fn + B;
      ode_row_D_dD(S, b, i, z, &D, &dD);
      const double P = c.Norm * pow(k, c.ns) * T * T * F * F * D * D;
      if (fn <= 1e-10) return P;
      const double Rr = 1.0 / (fc + B);
      return P * Rr * Rr;
    }();
    const double aL = a * 0.999, aR = fmin(1.0, a * 1.001);  // rt:1660-1661
    const double B_eta = ode_row_beta(S, b, i, a), B1 = ode_row_beta(S, b, i, 1.0);
    const double B_left = ode_row_beta(S, b, i, aL), B_right = ode_row_beta(S, b, i, aR);
    const double dlnB = (c.fnu < 1e-10) ? 0.0 : (a / B_eta) * (B_right - B_left) / (aR - aL);
    double Pnu = 0.0;
    {
      const double fn = c.On / c.Om, fc = 1.0 - fn;
      if (fn > 1e-10) {
        const double T = S.Tgrid[(long long)b * nk + i];
        const double F = 1.0 - fn + B_eta;
        const double P = c.Norm * pow(k, c.ns) * T * T * F * F * D * D;
        const double Rr = B_eta / fn / (fc + B_eta);
        Pnu = P * Rr * Rr;
      }
    }
    row[col++] = D;
    row[col++] = a * dD / D;
    row[col++] = Pcb;
    row[col++] = B_eta / (B1 + 1e-100);
    row[col++] = dlnB;
    row[col++] = Pnu;
  }
  row[col++] = exp(y[0]) * a2;
  row[col++] = exp(y[(long long)nk]) * a2;
  row[col++] = exp(y[2LL * nk]) * a2;
  const bool have_int = (c.sw_nl && c.sw_1l);  // rt:1646; otherwise the reference prints
                                               // uninitialised memory (zeros in its goldens)
  const double *s1 = S.src_v + (long long)(v - v0) * N_SRC * nk + i;
  if (S.print_A)
    for (int j = 0; j < N_UI; j++) row[col++] = have_int ? s1[(long long)j * nk] : 0.0;
  if (S.print_I)
    for (int j = 0; j < N_UI; j++) row[col++] = y[(long long)(N_UP + j) * nk];
  if (c.sw_pr) {
    double Q[N_UQ];
    for (int j = 0; j < N_UQ; j++) Q[j] = y[(long long)(N_UP + N_UI + j) * nk];
    const double pk = M_PI * k;
    if (S.print_bias) {
      row[col++] = pk * pbis_comb(Q, 2, 2) * a3;
      row[col++] = pk * pbis_comb(Q, 2, 1) * a3;
      row[col++] = pk * pbis_comb(Q, 4, 1) * a3;
      row[col++] = pk * pbis_comb(Q, 4, 0) * a3;
      row[col++] = pk * pbis_comb(Q, 6, 0) * a3;
      for (int j = 0; j < 9; j++) row[col++] = have_int ? s1[(long long)(38 + j) * nk] * a4 : 0.0;
      for (int j = 0; j < 8; j++) row[col++] = have_int ? s1[(long long)(47 + j) * nk] * a4 : 0.0;
    } else {
      row[col++] = (pk * pbis_comb(Q, 2, 2) + pk * pbis_comb(Q, 2, 1)) * a3;
      row[col++] = (pk * pbis_comb(Q, 4, 1) + pk * pbis_comb(Q, 4, 0)) * a3;
      row[col++] = pk * pbis_comb(Q, 6, 0) * a3;
      double PT[4] = {0, 0, 0, 0};
      if (have_int) {  // rt:1353-1358
        PT[0] = s1[38LL * nk] + s1[39LL * nk] + s1[40LL * nk];
        PT[1] = s1[41LL * nk] + s1[42LL * nk] + s1[43LL * nk];
        PT[2] = s1[44LL * nk] + s1[45LL * nk];
        PT[3] = s1[46LL * nk];
      }
      for (int j = 0; j < 4; j++) row[col++] = PT[j] * a4;
    }
  }
  if (S.print_Q)
    for (int j = 0; j < N_UQ; j++) row[col++] = y[(long long)(N_UP + N_UI + j) * nk] * a3;
  RT_ASSERT(col == ncols && (row - S.out) + ncols <= S.n_out_total && v < S.NO && io < c.n_out);
}

// ---------------------------------------------------------------------------- launchers
void launch_rhs(const Batch &S, const double *kgrid, const double *yv, double *dyv, int stage,
                const int *mask, cudaStream_t st) {
  const int nrows = S.k_hi - S.k_lo;
  if ((long long)S.B * nrows <= 2048)
    k_rhs<true><<<dim3((nrows + 31) / 32, S.B), 128, 0, st>>>(S, kgrid, yv, dyv, stage, mask);
  else
    k_rhs<false><<<dim3((nrows + 127) / 128, S.B), 128, 0, st>>>(S, kgrid, yv, dyv, stage, mask);
}
static size_t attempt_smem_bytes() { return (size_t)RK_STAGES * (N_UP + N_UI + 24) * ATT_ROWS * sizeof(double); }
int ode_configure() {
  return (int)cudaFuncSetAttribute(k_attempt_local, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)attempt_smem_bytes());
}
void launch_attempt_local(const Batch &S, const double *kgrid, const int *mask, cudaStream_t st) {
  const int nrows = S.k_hi - S.k_lo;
  k_attempt_setup<<<(S.B * RK_STAGES + 127) / 128, 128, 0, st>>>(S, mask);
  k_attempt_local<<<dim3((nrows + ATT_ROWS - 1) / ATT_ROWS, S.B), 128, attempt_smem_bytes(), st>>>(S, kgrid, mask);
}
// small launches: the fused path (k_stage_post) replaces k_assemble + k_rhs + k_combine / k_final
bool stage_post_applies(const Batch &S) { return (long long)S.B * (S.k_hi - S.k_lo) <= 2048; }
void launch_stage_post(const IntegralTabs &tb, const Batch &S, const double *kgrid, const double *yv, int stage, int next,
                       const int *mask, int groups, cudaStream_t st) {
  const int nrows = S.k_hi - S.k_lo;
  k_stage_post<<<dim3((nrows + STG_ROWS - 1) / STG_ROWS, S.B), STG_THREADS, 0, st>>>(tb, S, kgrid, yv, stage, next, mask, groups);
}
void launch_combine(const Batch &S, int stage, const int *mask, cudaStream_t st) {
  const long long n = (long long)N_U * S.nk;
  k_combine<<<dim3((unsigned)((n + 255) / 256), S.B), 256, 0, st>>>(S, stage, mask);
}
void launch_final(const Batch &S, const int *mask, cudaStream_t st) {
  const long long n = (long long)N_U * S.nk;
  k_final<<<dim3((unsigned)((n + 255) / 256), S.B), 256, 0, st>>>(S, mask);
}
void launch_ctrl_begin(const Batch &S, cudaStream_t st) { k_ctrl_begin<<<(S.B + 127) / 128, 128, 0, st>>>(S); }
void launch_ctrl_end(const Batch &S, int max_attempts, cudaStream_t st) {
  k_ctrl_end<<<(S.B + 127) / 128, 128, 0, st>>>(S, max_attempts);
}
void launch_accept(const Batch &S, cudaStream_t st) {
  const long long n = (long long)N_U * S.nk;
  k_accept<<<dim3((unsigned)((n + 255) / 256), S.B), 256, 0, st>>>(S);
}
void launch_stash(const Batch &S, cudaStream_t st) {
  const long long n = (long long)N_U * S.nk;
  k_stash<<<dim3((unsigned)((n + 255) / 256), S.B), 256, 0, st>>>(S);
}
void launch_vprep(const Batch &S, cudaStream_t st) { k_vprep<<<(S.NO + 127) / 128, 128, 0, st>>>(S); }
void launch_output(const Batch &S, const double *kgrid, int v0, int nv, cudaStream_t st) {
  k_output<<<dim3((S.nk + 127) / 128, nv), 128, 0, st>>>(S, kgrid, v0);
}
// device-side loop condition of rtrg_run's conditional (while) graph node
__global__ void k_loop_cond(cudaGraphConditionalHandle handle, const int *n_active, long long *rounds,
                            long long max_rounds) {
  const long long r = ++(*rounds);
  cudaGraphSetConditional(handle, (*n_active > 0 && r < max_rounds) ? 1u : 0u);
}
void launch_loop_cond(unsigned long long handle, const Batch &S, long long max_rounds, cudaStream_t st) {
  k_loop_cond<<<1, 1, 0, st>>>((cudaGraphConditionalHandle)handle, S.n_active, S.rounds, max_rounds);
}

}  // namespace rtrg
