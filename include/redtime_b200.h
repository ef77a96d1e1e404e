/* redtime_b200 -- C-ABI of the B200-native Time-RG hot path.
 *
 * Drop-in boundary for michaelbuehlmann/redTime (reference paths below are relative to
 * the reference repository).  The reference has no plugin API; its seams are
 *   (i)  the executable's file/stdout contract            src/redTime.cc:1551-1745
 *   (ii) the GSL callback `derivatives(eta,y,dy,params)`  src/redTime.cc:1416,1596
 * and one level below: compute_Aacdbef_Rlabc_PTjm_PMRn_full (src/redTime.cc:740-744),
 * C.D_dD (src/AU_cosmological_parameters.h:733), C.Beta_P (:632).
 * This header exposes those seams, batched over cosmologies, with plain pointers and
 * sizes only.  All functions return 0 on success, a negative RTRG_E* code otherwise;
 * nothing aborts or throws across the boundary.  A handle is thread-confined.  The
 * caller owns every host buffer; the library copies.  There is NO CPU fallback: every
 * compute entry point fails with RTRG_ENOGPU when no CUDA device is usable.
 */
#ifndef REDTIME_B200_H
#define REDTIME_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTRG_OK 0
#define RTRG_EINVAL (-1)   /* bad argument / malformed input                       */
#define RTRG_ENOGPU (-2)   /* no usable CUDA device (there is no CPU path)           */
#define RTRG_ECUDA (-3)    /* CUDA runtime error, see rtrg_last_error()             */
#define RTRG_ERANGE (-4)   /* look-up outside the reference's abort() bounds         */
#define RTRG_EODE (-5)     /* integrator failed for at least one cosmology           */
#define RTRG_ENOMEM (-6)

#define RTRG_MAX_OUT 64    /* output redshifts per cosmology                         */
#define RTRG_MAX_Z 200     /* interpolation redshifts (MAXTRANSFER, hdr:56)          */

typedef struct rtrg_handle rtrg_handle;

/* Compile-time constants of the reference turned into run-time fields, same defaults.
 * Reference: nk/np src/redTime.cc:90-97; tolerances :141-145; z1l :1285; PRINT* :65;
 * beta clamp src/AU_cosmological_parameters.h:537; n_lnk,a_early :664,697. */
typedef struct rtrg_config {
  int nk;              /* 128  (np = 4*nk)                                          */
  double kmin, kmax;   /* 1e-3, 1  [h/Mpc]                                          */
  double z1l;          /* 10   redshift of the cached 1-loop evaluation             */
  double eps_abs;      /* 1e-7 RKF45 control_y_new                                  */
  double eps_rel;      /* 1e-2                                                      */
  double beta_kmin;    /* 1e-3 clamp of Beta_P(a,k)                                 */
  double beta_kmax;    /* 1                                                         */
  int n_lnk;           /* 50   growth table: n_lnk+1 wavenumbers                    */
  int n_lna;           /* 100  growth table: n_lna+1 scale factors                  */
  double a_early;      /* 1e-20 start of the growth integration                     */
  int print_A, print_I, print_Q, print_bias; /* 0: optional output column groups    */
  int device;          /* CUDA device ordinal                                       */
  int max_attempts;    /* 100000: cap on RKF45 step attempts per cosmology          */
  int k_shards;        /* 1; >1: this handle owns k-rows [k_rank*nk/k_shards, ...)   */
  int k_rank;          /* 0                                                         */
  int v_split;         /* 1: CTAs per row block along the beta-side lags of the bilinear
                          quadrature.  > 1 shortens the serial work of a CTA when the grid is too
                          small to fill the GPU (one cosmology, k-sharded ranks); the partial sums
                          are added in a different order, so results move by round-off       */
  int reduce_beta;     /* 0: the full Beta_P(a,k) tables go to the device (every hook works);
                          1: the host sends only what rtrg_run consumes -- beta at a = 1 and the
                             table pre-reduced in k at the grid / growth wavenumbers (5x fewer
                             PCIe bytes); rtrg_Beta_P / rtrg_Plin then answer for a = 1 only   */
} rtrg_config;

/* One cosmology = the content of params_redTime.dat plus the CAMB tables it names
 * (src/AU_cosmological_parameters.h:231-353, :547-627, :790-832). */
typedef struct rtrg_cosmology {
  double params[9];      /* n_s sigma_8 h Omega_m Omega_b Omega_nu T_cmb w0 wa      */
  int switches[4];       /* nonlinear, 1loop, print_linear, print_rsd               */
  double z_in;
  int n_out;
  const double *z_out;   /* [n_out], greatest to least                              */
  int n_T;               /* rows of the z=0 transfer file                           */
  const double *k_T;     /* [n_T] column 0                                          */
  const double *Tc_T;    /* [n_T] column 1 (delta_cdm/k^2)                          */
  const double *Tb_T;    /* [n_T] column 2 (delta_b/k^2)                            */
  int n_z;               /* interpolation redshifts (0 => Beta_P == 0)              */
  const double *z_interp;/* [n_z] greatest to least (atof of the strings)           */
  int n_kb;              /* rows of each interpolation file                         */
  const double *k_b;     /* [n_kb] column 0 of the first file                       */
  const double *Tc_b;    /* [n_z][n_kb] column 1                                    */
  const double *Tnu_b;   /* [n_z][n_kb] column 5 (delta_massive_nu/k^2)             */
} rtrg_cosmology;

void rtrg_default_config(rtrg_config *cfg);
const char *rtrg_last_error(void);
const char *rtrg_version(void);

/* Builds the cosmology-independent quadrature weights (FAST-PT-equivalent kernels T_n,
 * Z kernels G_n, windows) and uploads them.  Replaces the per-call recomputation in
 * src/redTime.cc:306-355,411-597,689-727. */
int rtrg_create(const rtrg_config *cfg, rtrg_handle **out);
int rtrg_destroy(rtrg_handle *h);
/* Launch on a caller-provided CUDA stream (cudaStream_t passed as void*); NULL = own. */
int rtrg_set_stream(rtrg_handle *h, void *cuda_stream);

/* rtrg_add_cosmology copies the tables, so the caller's buffers may be freed right after the
 * call -- EXCEPT tables that live in page-locked memory (cudaHostAlloc / cudaHostRegister / torch
 * pin_memory): the four small ones k_T, Tc_T, Tb_T, k_b when all of them are page-locked, and the
 * two interpolation tables Tc_b, Tnu_b when both are (and reduce_beta = 0).  Those are sent to the
 * device directly by the copy engine, without a host-side copy, and must stay valid and unchanged
 * until rtrg_prepare() has returned (and for every later rtrg_prepare() on the same batch).    */
int rtrg_clear_cosmologies(rtrg_handle *h);
int rtrg_add_cosmology(rtrg_handle *h, const rtrg_cosmology *c);
/* the same for n cosmologies at once; the table copies run on several host threads */
int rtrg_add_cosmologies(rtrg_handle *h, int n, const rtrg_cosmology *const *list);
int rtrg_num_cosmologies(const rtrg_handle *h);
/* columns of the table of cosmology i (src/redTime.cc:1670-1737): 17 by default */
int rtrg_num_columns(const rtrg_handle *h, int icosmo);

/* Host->device upload of all added cosmologies, then on the device: Beta_P tables,
 * growth tables (RK8PD), sigma_8 normalisation and sigma_v^2 (QAG-61), initial
 * conditions, 1-loop cache.  Replaces src/redTime.cc:1559-1586 +
 * src/AU_cosmological_parameters.h:513-971 first-call initialisations. */
int rtrg_prepare(rtrg_handle *h);
/* Device-side part of rtrg_prepare() alone, on the inputs already resident in HBM (no
 * host->device traffic): used to time the hot path with resident inputs. */
int rtrg_device_init(rtrg_handle *h);

/* The whole job (src/redTime.cc:1588-1742): evolve every cosmology with the
 * device-resident RKF45 stepper and write the output tables.
 *   out   : concatenated per cosmology, [n_out][nk][ncols] each (row-major)
 *   hdr   : [n_cosmo][RTRG_MAX_OUT][5] = eta, a, z, H, sigma_v^2 per output
 *   hdr0  : [n_cosmo][2] = eta_fin, sigmaV2(z=0)
 *   status: [n_cosmo] 0 = ok (may be NULL)
 * out / hdr / hdr0 may be NULL: the tables then stay on the device (no D2H copy).     */
int rtrg_run(rtrg_handle *h, double *out, size_t out_len, double *hdr, double *hdr0,
             int *status);
/* Zero-copy read-back after rtrg_run(h, NULL, 0, NULL, NULL, status): one device->host copy
 * into page-locked buffers owned by the handle.  The returned pointers (layouts as in rtrg_run)
 * stay valid until the next rtrg_prepare / rtrg_fetch_outputs / rtrg_destroy on this handle. */
int rtrg_fetch_outputs(rtrg_handle *h, const double **out, size_t *out_len, const double **hdr,
                       const double **hdr0);

/* ---- double-buffered batch pipeline (what replaces the sequential model loop of
 * scripts/runRedTimeBatch:91-99) -----------------------------------------------------------
 * `depth` handles on cfg->device (2 = double buffering) behind two worker threads: while batch i
 * evolves (rtrg_run + rtrg_fetch_outputs on the run thread), batch i+1 is staged, uploaded and
 * initialised (rtrg_add_cosmologies + rtrg_prepare on the stage thread), so host staging and the
 * PCIe transfers hide behind the GPU time of the previous batch.  Results are bit-identical to the
 * serial add / prepare / run sequence.
 *   submit : returns at once with a ticket; the caller's tables (and the list array's entries)
 *            must stay valid and unchanged until rtrg_pipeline_wait(ticket) has returned
 *   wait   : blocks until the batch is done; out / hdr / hdr0 (layouts as in rtrg_run) point into
 *            page-locked memory of the pipeline and status[n] holds the per-cosmology status;
 *            all stay valid until rtrg_pipeline_release(ticket).  Returns the batch's error code
 *   release: frees the batch's slot for a later submission (at most `depth` unreleased batches
 *            are in flight; further submissions queue)                                      */
typedef struct rtrg_pipeline rtrg_pipeline;
int rtrg_pipeline_create(const rtrg_config *cfg, int depth, rtrg_pipeline **out);
int rtrg_pipeline_submit(rtrg_pipeline *p, int n, const rtrg_cosmology *const *list, long long *ticket);
int rtrg_pipeline_wait(rtrg_pipeline *p, long long ticket, const double **out, size_t *out_len,
                       const double **hdr, const double **hdr0, const int **status);
/* columns of cosmology i of a finished batch (rtrg_num_columns) */
int rtrg_pipeline_columns(rtrg_pipeline *p, long long ticket, int icosmo);
int rtrg_pipeline_release(rtrg_pipeline *p, long long ticket);
/* Diagnostics: host-clock seconds since rtrg_pipeline_create at which the job's stage (staging,
 * H2D, initialisation), run (evolution) and fetch (D2H) phases began and ended:
 * t = {stage0, stage1, run0, run1, fetch0, fetch1}.  Valid once rtrg_pipeline_wait() has returned. */
int rtrg_pipeline_times(rtrg_pipeline *p, long long ticket, double t[6]);
int rtrg_pipeline_destroy(rtrg_pipeline *p);

/* ---- one high-resolution cosmology sharded over its k-rows ---------------
 * Create one handle per rank with cfg.k_shards = number of ranks and cfg.k_rank = this rank
 * (nk/8 must be divisible by k_shards), add the SAME cosmologies to each, attach a transport,
 * then call rtrg_prepare / rtrg_run on every rank.  Before each integral evaluation the ranks
 * all-gather their block of the three ln P_ab components (3*nk/k_shards doubles per rank),
 * and after each RKF45 attempt they max-reduce the error norm, so all ranks replay the same
 * accept/reject sequence; at the end every rank holds the complete output tables.
 *   NCCL     : one process per GPU.  Rank 0 calls rtrg_kshard_nccl_id and ships the 128 bytes
 *              to the others (MPI, torch.distributed, a file ...).
 *   loopback : all ranks live in one process, each handle driven by its own host thread;
 *              blocks move with peer copies (single-GPU testing, or NVLink P2P without NCCL). */
int rtrg_kshard_nccl_id(char id[128]);
int rtrg_kshard_init_nccl(rtrg_handle *h, const char id[128]);
/* description of the transport attached to the handle ("none" without one) */
const char *rtrg_kshard_transport(const rtrg_handle *h);
typedef struct rtrg_loopback rtrg_loopback;
int rtrg_kshard_loopback_create(int nranks, rtrg_loopback **out);
int rtrg_kshard_init_loopback(rtrg_handle *h, rtrg_loopback *g);
void rtrg_kshard_loopback_free(rtrg_loopback *g);

/* Work counters of the last rtrg_run for cosmology i:
 * counters[0]=RKF45 attempts, [1]=rejected, [2]=RHS evaluations, [3]=integral evaluations */
int rtrg_counters(const rtrg_handle *h, int icosmo, long long counters[4]);
/* Error norms rmax = max_i |yerr_i| / (eps_rel |y_i| + eps_abs) of the first (up to 128) RKF45
 * attempts of cosmology i in the last rtrg_run, in order; returns how many were written (<= cap).
 * GSL's controller rejects an attempt when rmax > 1.1 (gsl odeiv control.c: std_control_hadjust,
 * driven from src/redTime.cc:1616): an rmax within round-off of 1.1 marks a cosmology whose step
 * sequence -- and with it P(k) at the 1e-4 level -- flips under 1-ulp changes of the inputs.   */
int rtrg_rmax_history(rtrg_handle *h, int icosmo, double *out, int cap);
/* (kernel, beta-side spectrum) matrix-vector sets k_bilinear executed for cosmology i since the
 * last rtrg_device_init, up to the end of the last rtrg_run; one set = nk rows x (2 nsup^2 +
 * 6 nsup) FLOP.  Only the products the requested outputs consume are computed.            */
long long rtrg_matvec_sets(const rtrg_handle *h, int icosmo);
/* number of kernel launches issued by this handle so far */
long long rtrg_launch_count(const rtrg_handle *h);
/* Per-kernel device timing (CUDA events on the launching stream), off by default.
 * Switching it on or off zeroes the totals.  Categories 0..rtrg_profile_categories()-1. */
int rtrg_set_profiling(rtrg_handle *h, int on);
int rtrg_profile_categories(void);
const char *rtrg_profile_name(int cat);
int rtrg_profile_query(const rtrg_handle *h, int cat, long long *n_launches, double *total_ms);

/* Tuning aid: evaluate the mode-coupling integrals of the linear spectra at z1l for every
 * cosmology of the batch `reps` times; read the timings with rtrg_profile_query() and the work
 * with rtrg_matvec_sets().  groups: bit 0 A, 1 R, 2 P_T,jm, 3 P_MR,n, 4 every raw product;
 * identical != 0 uses the identical-spectra shortcut of the 1-loop cache.  Results go to the
 * scratch source buffer only.                                                            */
int rtrg_bench_integrals(rtrg_handle *h, int reps, int groups, int identical);
/* FP64 pipe peak of `device` in TFLOP/s, best launch over about `seconds` of device time.
 * rtrg_bench_dmma: register-resident DMMA.8x8x4 loop (the instruction the bilinear kernel runs on;
 * reaches the nominal 64 FMA/clk/SM) -- the roofline denominator of the integral kernels.
 * rtrg_bench_dfma: register-resident scalar DFMA loop (about 92 % of the nominal rate).   */
int rtrg_bench_dmma(int device, double seconds, double *tflops);
int rtrg_bench_dfma(int device, double seconds, double *tflops);

/* ---- stage-level hooks with the reference's array layouts (parity tests) ---------- */
/* src/redTime.cc:772-778: y[0..3nk) -> P[3][np] extrapolated and windowed            */
int rtrg_extrap_P(rtrg_handle *h, int icosmo, const double *lnP3nk, double *P3np);
/* src/redTime.cc:740-1282: A[64*nk], R[24*nk], PTjm[9*nk], PMRn[8*nk]                */
int rtrg_integrals_full(rtrg_handle *h, int icosmo, const double *lnP3nk, double *A64,
                        double *R24, double *PTjm9, double *PMRn8);
/* raw bilinear quadratures J[63][nk], PZ[63][nk], Jn0[63][nk] at the nk output rows
 * (src/redTime.cc:781-811; index 9n+3ab+cd) and J[0][nloMR]                          */
int rtrg_integrals_raw(rtrg_handle *h, int icosmo, const double *lnP3nk, double *J63,
                       double *PZ63, double *Jn063, double *Jlo);
/* src/redTime.cc:1416-1547: the Time-RG right-hand side, y,dy of length 41*nk        */
int rtrg_derivatives(rtrg_handle *h, int icosmo, double eta, const double *y, double *dy);
/* src/AU_cosmological_parameters.h:639-738 and :513-637                              */
int rtrg_D_dD(rtrg_handle *h, int icosmo, double z, const double *k, int n, double *D,
              double *dDda);
int rtrg_Beta_P(rtrg_handle *h, int icosmo, double a, const double *k, int n, double *beta);
/* src/AU_cosmological_parameters.h:834-930: which = 0 Plin, 1 Plin_cb, 2 Plin_nu     */
int rtrg_Plin(rtrg_handle *h, int icosmo, int which, double z, const double *k, int n,
              double *P);
/* initial state y[41*nk] (src/redTime.cc:1570-1586) and scalars
 * scal[0]=sigma8 Norm, [1]=sigmaV2(z=0) (src/AU_cosmological_parameters.h:874,961)   */
int rtrg_initial_state(rtrg_handle *h, int icosmo, double *y, double scal[2]);

/* ---- cosmology-independent tables (host side, usable without a GPU) ---------------- */
/* grid: out[0]=np, [1]=nshift, [2]=jlo, [3]=nsup, [4]=nloMR */
int rtrg_grid_info(int nk, double kmin, double kmax, int out[5], double *dlnk,
                   double *lnk_pad_min);
/* T_n (n<14): np*np doubles, row-major [u][v]; kfac: np doubles */
int rtrg_table_T(int nk, double kmin, double kmax, int n, double *T, double *kfac);
/* G_n (n<7): 2*np-1 doubles, G[d+np-1] */
int rtrg_table_G(int nk, double kmin, double kmax, int n, double *G);
/* windows WP[np], WC[np] */
int rtrg_table_windows(int nk, double kmin, double kmax, double *WP, double *WC);
/* Pab stencil on the padded grid (replaces Pab(), src/redTime.cc:181-232):
   ln P(k_pad[ip]) = sum_j w[4 ip + j] lnP[n0[ip] + j] + (n_s - 3) dx[ip];  n0[np], w[4 np], dx[np] */
int rtrg_table_extrap(int nk, double kmin, double kmax, int *n0, double *w, double *dx);
/* assembly terms: returns count; arrays may be NULL to query the count */
int rtrg_assembly_terms(int *row, int *src, int *index, int *kpow, double *coef, int cap);

/* ---- host-side I/O mirroring the reference executable ------------------------------ */
typedef struct rtrg_run_inputs rtrg_run_inputs;
/* parse <dir>/params_redTime.dat and the CAMB files it names (hdr:231-353,547-627,790-832);
 * camb_modern != 0 selects the 13-column layout (hdr:76-80) */
int rtrg_read_run_dir(const char *dir, int camb_modern, rtrg_run_inputs **out);
/* n run directories at once, spread over the host threads; out[n] receives the handles */
int rtrg_read_run_dirs(int n, const char *const *dirs, int camb_modern, rtrg_run_inputs **out);
const rtrg_cosmology *rtrg_inputs_cosmology(const rtrg_run_inputs *in);
void rtrg_free_run_inputs(rtrg_run_inputs *in);
/* print one cosmology's result exactly as the reference's main() does (stdout format of
 * src/redTime.cc:1602-1603,1639-1641,1670-1741; banner line of hdr:236) */
int rtrg_print_result(void *cfile, const char *paramfile_name, int nk, int ncols, int n_out,
                      const double *out, const double *hdr, const double *hdr0);

/* the printer's number format, "%20.12g", for n values: dst receives 20 n characters + NUL */
int rtrg_format_g12(const double *v, int n, char *dst);

#ifdef __cplusplus
}
#endif
#endif /* REDTIME_B200_H */
