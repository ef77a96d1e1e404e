// Device-side data layout of one batch of cosmologies (all pointers are device pointers).
//
// HBM layout (doubles unless noted; B = cosmologies in the batch, nk output wavenumbers,
// np = 4 nk padded grid, N_EQ = 41 nk):
//   state vectors      y, ytmp, ynew, yerr, k[6]     [B][41][nk]   component-major like the
//                                                    reference's y[c*nk+i] (rt:1418-1423)
//   beta rows          bred   [B][n_zmax][nkk]       Beta_P pre-reduced in k at the nk grid
//                                                    wavenumbers and the n_lnk+1 growth ones
//   growth tables      G, dD  [B][n_lna+1][n_lnk+1]; Dnorm [B][n_lnk+1]
//   growth rows        Grow, dDrow [B][n_lna+1][nk]; D0row [B][nk]
//   integrals          Prev   [B][3][LP]             reversed, zero-padded P q^2 (TMA source)
//                      P3     [B][3][np]             extrapolated, windowed spectra
//                      Jpart  [B][14][nchunk*vsplit][9][nk] partial bilinear sums
//                      PZb    [B][7][3][nk]
//                      src    [B][55][nk]            A14, R24, PTjm9, PMRn8
//   inputs             in     [sum over cosmologies of 3 nT + n_z + n_kb + 2 n_z n_kb]
//                                                    raw CAMB columns, transformed in place
//   weight tables      Tc     [14+3][NUp/8][ldT/8][32][2]  compact circulant kernels (+ 3 transposes) as 8 x 8 tiles in
//                                                    DMMA operand order (L2 resident)
#pragma once
#include "rtrg_math.h"
#ifdef __CUDACC__
#include <cuda_runtime.h>
#include <vector>
#endif

// -DRTRG_BOUNDS: the debug build (libredtime_b200_bounds.so) checks every computed index against
// the extent of the array it addresses with device-side asserts -- the stand-in for
// compute-sanitizer's memcheck, which is closed on the GPU pool.  tests/test_gpu_bounds.py runs the
// end-to-end paths on that build; a violated assert surfaces as cudaErrorAssert.
#ifdef RTRG_BOUNDS
#include <cassert>
#define RT_ASSERT(cond) assert(cond)
#else
#define RT_ASSERT(cond) ((void)0)
#endif

namespace rtrg {

enum { RK_STAGES = 6, MAX_OUT = 64, N_SRC = 55, RMAX_HIST = 128 };
#ifndef RTRG_NKERN_DEFINED
#define RTRG_NKERN_DEFINED
enum { N_JKERN = 14, N_ZKERN = 7 };  // bilinear kernels (J + Jn0) and Z kernels
#endif
enum { BIL_R = 8 };      // output rows per row block of the bilinear kernel
// Transposed copies of the asymmetric Jn0 kernels (alpha, beta, l) = (0, 2, l), n = 7, 8, 9: table
// N_JKERN + n - 7 holds T_n^T, with which J_n(ab, cd) = a_ab^T T_n b_cd is evaluated as b_cd^T T_n^T a_ab,
// i.e. with the matrix-vector product on the ab side.  The default output columns need the pairs
// (ab, cd) = (2, 1), (2, 2) of these kernels: one product T_n^T a_2 instead of T_n b_1 and T_n b_2.
enum { N_TKERN = 3, TKERN_FIRST = 7 };
// output groups of one evaluation of the mode-coupling integrals
enum { GRP_A = 1, GRP_R = 2, GRP_PT = 4, GRP_PMR = 8, GRP_ALL = 15, GRP_RAW = 16 };
enum { RTRG_QAG_FAIL = 101, RTRG_ODE_FAIL = 102, RTRG_RANGE_FAIL = 103 };  // Cosmo::status

struct IntegralTabs {
  int nk, np, nshift, jlo, nsup, nloMR;
  int NV;      // nsup + BIL_R - 1 rounded up to whole 8-lag tiles: lag range a row block touches
  int NVp;     // NV rounded up to a multiple of BIL_R
  int LP;      // padded length of one reversed spectrum (even)
  int NUp;     // rows of the compact kernel table (>= nk + NVp)
  int ldT;     // leading dimension of the compact kernel table
  int nchunk;  // partial sums per row block: the most CTAs along the item axis one row block spans
  int tpb;     // threads per CTA of the bilinear kernel
  long long n_Tc;  // doubles in Tc (bounds checks of the debug build)
  int vsplit;  // CTAs along the beta-side lag dimension (rtrg_config.v_split)
  double dlnk;     // grid spacing in ln k
  double kfac_lo;  // k-dependent prefactor of kernel 0 at the padded row nloMR
  const double *Tc;    // [14 + N_TKERN][NUp/8][ldT/8][32][2]  tile (v''/8, u''/8), lane 4 (u''&7) + (v''&7)/2, half v''&1 = T_n[u][v]
  const double *Tlo;   // [nsup][nsup]    kernel 0 at the low-k row nloMR (reversed indices)
  const double *kfac;  // [14][nk]
  const double *G;     // [7][2np-1]
  const double *WP;    // [np]
  const double *kpad;  // [np]
  const double *kgrid; // [nk]
  // extrapolation stencil of Pab (rt:181-232) for every padded sample
  const int *ex_n0;       // [np] first node of the 4-point stencil
  const double *ex_w;     // [np][4]
  const double *ex_dx;    // [np] lnk - lnk[nk-1] for the power-law extrapolation, else 0
  // beta-side spectra (bit c = P_{cd=c}) of kernel n consumed by output group g = A, R, PT, PMR
  unsigned char need_cd[4][N_JKERN];
  unsigned char need_ab[4][N_JKERN];  // the same for the alpha side (bit a = P_{ab=a})
  unsigned int need_pz[4];  // bit 3 n + ab: the PZ_n(P_ab) log-convolutions an output group consumes
  unsigned long long need_val[4][3];  // bit v: raw value v (of 190, stage_device.h) is consumed by the group
  // assembly table (sorted by output row)
  int n_terms;
  const int *t_start;     // [56]
  const short *t_src, *t_index, *t_kpow;
  const double *t_coef;
};

struct Batch {
  int B, nk, np, n_lna, n_lnk, nkk, n_zmax;
  double eps_abs, eps_rel, z1l, beta_kmin, beta_kmax, a_early;
  int print_A, print_I, print_Q, print_bias;
  int k_lo, k_hi;  // k-rows owned by this rank (k-sharding); [0,nk) otherwise
  long long n_in, n_out_total, n_slots;  // extents of in, out and of the integral work space (in cosmologies)
  Cosmo *cosmo;            // [B]
  double *zout, *aout, *etaout;  // [B][MAX_OUT] output redshifts, 1/(1+z), ln(a/a_in)
  // pooled input tables (one buffer; per-cosmology offsets in Cosmo)
  double *in;
  // linear-theory work
  double *bred;            // [B][n_zmax][nkk]
  const double *lna, *lnkg;// [n_lna+1], [n_lnk+1]
  double *G, *dD, *Dnorm;
  double *Grow, *dDrow, *D0row;
  double *Tgrid;           // [B][nk] transfer function at the grid wavenumbers
  // 1-loop cache at z1l (rt:1291-1313)
  double *src_z1l;         // [B][55][nk]
  double *D_z1l;           // [B][nk]
  double *y_z1l;           // [B][3][nk] ln P_lin,cb(z1l) three times (rt:1299-1306)
  // ODE state
  double *y, *ytmp, *ynew, *yerr, *kst;   // kst: [6][B][41][nk]
  double *src;             // [B][55][nk]  sources of the current RHS evaluation
  // integral work space
  double *Prev, *P3, *Jpart, *PZb, *Jlo;
  // stepper control (per cosmology)
  double *t, *h, *h_try;   // current eta, suggested step, step of this attempt
  unsigned long long *rmax_bits;
  int *i_out, *flag_out, *flag_step, *flag_acc, *final_step, *done;
  int *m_full_step, *m_full_acc, *m_out_int;  // masks of the integral launches
  int *m_loc_step;  // stepping cosmologies whose RHS needs no new integrals (k_attempt_local)
  void *att_time;   // [B][RK_STAGES] time-only stage quantities of the current attempt (k_attempt_setup)
  long long *counters;     // [B][4]
  double *rmax_hist;       // [B][RMAX_HIST] error norm of every attempt (diagnostics: decision margins)
  long long *matvecs;      // [B] (kernel, spectrum) matrix-vector sets executed since device_init
  int *act, *nact;         // [B], [1] compacted list of the cosmologies of the current launch
  int *n_active;           // [1]
  long long *rounds;       // [1] rounds executed by the device-side loop of rtrg_run
  // deferred outputs: the round loop only stashes the state at every output redshift; the
  // 1-loop output integrals (rt:1646-1653) and the tables are then produced for all NO =
  // sum n_out (cosmology, output) pairs ("virtual cosmologies") in a few large launches
  int NO;
  double *ystash;          // [NO][41][nk]
  double *t_stash;         // [NO] eta reached when the output was taken
  int *vbase;              // [B] first virtual index of cosmology b
  int *vc_b, *vc_io;       // [NO] cosmology and output index of virtual cosmology v
  int *vc_have;            // [NO] stashed in this run
  int *vc_mask;            // [NO] needs the output integrals
  Cosmo *cosmo_v;          // [NO] copy of cosmo[vc_b[v]] (the integral kernels index it by v)
  long long *matvecs_v;    // [NO]
  double *src_v;           // [VCH][55][nk] sources of one chunk of virtual cosmologies
  // outputs
  double *out;             // concatenated tables
  long long *out_off;      // [B] offset of cosmology b in out
  int *ncols;              // [B]
  double *hdr;             // [B][MAX_OUT][5]
  double *hdr0;            // [B][2]
};

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline
// numbers).  Off by default: a null Profiler* costs nothing.
enum ProfCat {
  PC_EXTRAP, PC_BILINEAR, PC_JLO, PC_PZ, PC_ASSEMBLE, PC_RHS, PC_COMBINE, PC_FINAL, PC_CTRL,
  PC_ACCEPT, PC_OUTPUT, PC_ATTEMPT, PC_STAGE, PC_XCH, PC_PREP_INPUTS, PC_BETA_REDUCE, PC_GROWTH_ODE, PC_GROWTH_TABS, PC_QAG, PC_INIT_STATE,
  PC_NCAT
};
#ifdef __CUDACC__
struct Profiler {
  struct Rec { int cat; cudaEvent_t e0, e1; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
  void tic(int cat, cudaStream_t st) {
    Rec r = {cat, get(), get()};
    cudaEventRecord(r.e0, st);
    recs.push_back(r);
  }
  void toc(cudaStream_t st) { cudaEventRecord(recs.back().e1, st); }
  void reset() { recs.clear(); used = 0; }
  ~Profiler() { for (cudaEvent_t e : pool) cudaEventDestroy(e); }
};
// second stream of a handle for launch branches that may run beside the main one (launch_integrals)
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
#define RT_TIC(prof, cat, st) do { if (prof) (prof)->tic(cat, st); } while (0)
#define RT_TOC(prof, st) do { if (prof) (prof)->toc(st); } while (0)
#endif

RT_HD BetaTab beta_tab(const Batch &S, const Cosmo &c) {
  BetaTab t;
  t.n_z = c.n_z;
  t.n_kb = c.n_kb;
  t.a = S.in + c.offA;
  t.k = S.in + c.offKb;
  t.beta = c.offRow1 >= 0 ? nullptr : S.in + c.offB;
  t.row1 = c.offRow1 >= 0 ? S.in + c.offRow1 : nullptr;
  t.fn = c.On / c.Om;
  t.kmin = S.beta_kmin;
  t.kmax = S.beta_kmax;
  return t;
}
RT_HD GrowthTab growth_tab(const Batch &S, int b) {
  GrowthTab g;
  g.n_lna = S.n_lna;
  g.n_lnk = S.n_lnk;
  g.lna = S.lna;
  g.lnk = S.lnkg;
  const long long n = (long long)(S.n_lna + 1) * (S.n_lnk + 1);
  g.G = S.G + b * n;
  g.dD = S.dD + b * n;
  g.Dnorm = S.Dnorm + (long long)b * (S.n_lnk + 1);
  return g;
}
RT_HD LinCtx lin_ctx(const Batch &S, int b) {
  LinCtx L;
  L.c = &S.cosmo[b];
  L.bt = beta_tab(S, S.cosmo[b]);
  L.gt = growth_tab(S, b);
  L.lnkT = S.in + S.cosmo[b].offT;
  L.lnT = S.in + S.cosmo[b].offLT;
  L.nT = S.cosmo[b].nT;
  return L;
}

}  // namespace rtrg
