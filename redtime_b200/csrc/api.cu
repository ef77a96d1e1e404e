// C-ABI of the B200-native Time-RG hot path (see include/redtime_b200.h).
// Host orchestration only: table upload, batch staging, kernel launch sequences.  Every
// numerical result is produced by the CUDA kernels; there is no CPU execution path.
#include <cuda_runtime.h>

#include <sched.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/redtime_b200.h"
#include "fastpt_tables.h"
#include "kshard.h"
#include "rtrg_device.h"

namespace rtrg {
// kernels_integrals.cu
int launch_integrals(const IntegralTabs &tb, const Batch &S, const double *y, long long ystride,
                     double *src, double *raw, const int *mask, int groups, int identical,
                     cudaStream_t st, Profiler *prof, const SideStream *side = nullptr, bool assemble = true);
void launch_extrap_only(const IntegralTabs &tb, const Batch &S, const double *y, long long ystride,
                        const int *mask, cudaStream_t st);
int integrals_configure(const IntegralTabs &tb);
size_t bilinear_smem_bytes(const IntegralTabs &tb);
int bilinear_tpb();
// host_stage.cc
void beta_row_cubic(const double *tn, const double *tc, size_t n, double fn, const double *x, double xq, double *row1);
void beta_row_linear(const double *tn, const double *tc, size_t n, double fn, double x0, double x1, double xq, double *row1);
// kernels_linear.cu
int linear_upload_constants();
int launch_linear_init(const Batch &S, const double *kgrid, cudaStream_t st, Profiler *prof);
int launch_prep_inputs(const Batch &S, double *T0, int max_rows, cudaStream_t st, Profiler *prof);
void launch_hook_DdD(const Batch &S, int b, double z, const double *k, int n, double *D, double *dD,
                     int *err, cudaStream_t st);
void launch_hook_beta(const Batch &S, int b, double a, const double *k, int n, double *beta, int *err,
                      cudaStream_t st);
void launch_hook_plin(const Batch &S, int b, int which, double z, const double *k, int n, double *P,
                      cudaStream_t st);
// kernels_ode.cu
int ode_configure();
void launch_rhs(const Batch &S, const double *kgrid, const double *yv, double *dyv, int stage,
                const int *mask, cudaStream_t st);
void launch_attempt_local(const Batch &S, const double *kgrid, const int *mask, cudaStream_t st);
void launch_combine(const Batch &S, int stage, const int *mask, cudaStream_t st);
bool stage_post_applies(const Batch &S);
void launch_stage_post(const IntegralTabs &tb, const Batch &S, const double *kgrid, const double *yv, int stage, int next,
                       const int *mask, int groups, cudaStream_t st);
void launch_final(const Batch &S, const int *mask, cudaStream_t st);
void launch_ctrl_begin(const Batch &S, cudaStream_t st);
void launch_ctrl_end(const Batch &S, int max_attempts, cudaStream_t st);
void launch_accept(const Batch &S, cudaStream_t st);
void launch_stash(const Batch &S, cudaStream_t st);
void launch_vprep(const Batch &S, cudaStream_t st);
void launch_output(const Batch &S, const double *kgrid, int v0, int nv, cudaStream_t st);
void launch_loop_cond(unsigned long long handle, const Batch &S, long long max_rounds, cudaStream_t st);
}  // namespace rtrg

using namespace rtrg;

static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(RTRG_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                  \
  } while (0)

struct HostCosmo {
  Cosmo c;  // offsets point into the device input pool
  std::vector<double> z_out;
  // what the host copies into the staging arena: slice [stage_off, stage_off + stage_len) of the
  // arena = slice of the pool at pool_off
  size_t stage_off = 0, stage_len = 0, pool_off = 0;
  // page-locked caller tables are sent as they are (no host copy): up to 6 direct segments
  // k_T, Tc_T, Tb_T, k_b and the raw interpolation tables T_nu, T_c
  int n_direct = 0;
  size_t direct_len = 0;
  const double *dsrc[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t dlen[6] = {0, 0, 0, 0, 0, 0};
  long long doff[6] = {0, 0, 0, 0, 0, 0};
  bool small_direct = false;  // k_T, Tc_T, Tb_T, k_b go out directly
  bool direct = false;        // T_nu, T_c go out directly (beta is formed on the device)
};

// Pinned host arena mirroring the device input pool: rtrg_add_cosmology copies the caller's
// buffers here (the caller owns its buffers and may free them right away); rtrg_prepare sends
// the not-yet-uploaded tail to the device with one asynchronous copy.
struct StagingArena {
  double *base = nullptr;
  size_t cap = 0, used = 0;
  int reserve(size_t n_total) {
    if (n_total <= cap) return RTRG_OK;
    size_t ncap = std::max(n_total, cap + cap / 2);
    double *nb = nullptr;
    if (cudaHostAlloc((void **)&nb, ncap * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return RTRG_ENOMEM;
    }
    if (used) std::memcpy(nb, base, used * sizeof(double));
    if (base) cudaFreeHost(base);
    base = nb;
    cap = ncap;
    return RTRG_OK;
  }
  void release() {
    if (base) cudaFreeHost(base);
    base = nullptr;
    cap = used = 0;
  }
};

// One device allocation carved into the batch work buffers; kept across rtrg_prepare calls
// while it is large enough (no cudaMalloc/cudaFree churn between batches).
struct DeviceArena {
  char *base = nullptr;
  size_t cap = 0, used = 0;
  template <class T>
  T *take(size_t n) {
    used = (used + 255) & ~(size_t)255;
    T *p = (T *)(base ? base + used : nullptr);
    used += (n ? n : 1) * sizeof(T);
    return p;
  }
};

struct rtrg_handle {
  rtrg_config cfg;
  GridSpec grid;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  // device-side initialisation (rtrg_device_init) runs on a stream of the highest priority: its
  // latency-bound kernels (growth ODE, QAG) get the SM slots another handle's k_bilinear launch frees,
  // i.e. they run beside it instead of after it (pipelines of several handles on one GPU)
  cudaStream_t init_stream = nullptr;
  IntegralTabs tb;
  std::vector<void *> table_allocs, batch_allocs;
  StagingArena stage;
  DeviceArena work;
  double *d_in = nullptr;     // device input pool (mirror of stage)
  size_t d_in_cap = 0, d_in_used = 0;
  bool d_in_transformed = false;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_done = nullptr;
  SideStream side;  // parallel launch branch of the integral evaluations
  // pinned host mirror of the outputs (rtrg_fetch_outputs)
  double *h_out = nullptr;
  size_t h_out_cap = 0;
  double *d_T0 = nullptr;
  double *d_kgrid = nullptr;
  std::vector<double> kgrid;
  std::vector<double> slot_k;  // reduce_beta: clamped wavenumbers of the nkk pre-reduced beta columns
  std::vector<HostCosmo> cos;
  Batch S;
  bool prepared = false, uploaded = false;
  long long launches = 0;
  std::vector<long long> out_off, counters, matvecs;
  std::vector<int> ncols;
  size_t out_total = 0;
  int vch = 0;  // virtual cosmologies (cosmology, output) per launch of the deferred output stage
  bool any_full = false, any_1loop = false, any_pr = false, any_local = false;
  double *d_yinit = nullptr, *d_raw = nullptr, *d_scratch = nullptr;
  int *d_hookmask = nullptr, *d_err = nullptr, *d_minit = nullptr;
  size_t scratch_len = 0;
  std::unique_ptr<Exchange> xch;  // k-shard transport (k_shards > 1)
  GatherPlan plan_lnP, plan_out;  // what the ranks exchange: ln P rows of a state vector, output rows
  Profiler profiler;
  Profiler *prof = nullptr;  // &profiler when profiling is switched on
  double prof_ms[PC_NCAT] = {0};
  long long prof_n[PC_NCAT] = {0};
};

// timed launch of one of the ODE-side kernels
#define ODE_LAUNCH(cat, call)        \
  do {                               \
    RT_TIC(h->prof, cat, st);        \
    call;                            \
    RT_TOC(h->prof, st);             \
    h->launches++;                   \
  } while (0)

// Output groups an evaluation has to deliver (see assembly need table):
//   inside the RHS / for the z1l cache: A, and R when some cosmology evolves Q (rt:1516)
//   at an output redshift in 1-loop mode (rt:1646): what the printed columns consume
static int grp_rhs(const rtrg_handle *h) { return GRP_A | ((h->cfg.print_Q || h->any_pr) ? GRP_R : 0); }
static int grp_out(const rtrg_handle *h) {
  return (h->any_pr ? GRP_PT : 0) | (h->cfg.print_A ? GRP_A : 0) | ((h->cfg.print_bias && h->any_pr) ? GRP_PMR : 0);
}

// fold the recorded events into the per-category totals (stream must be idle)
static void prof_collect(rtrg_handle *h) {
  if (!h->prof) return;
  for (const Profiler::Rec &r : h->profiler.recs) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      h->prof_ms[r.cat] += ms;
      h->prof_n[r.cat]++;
    }
  }
  cudaGetLastError();
  h->profiler.reset();
}

// ---- on-disk cache of the cosmology-independent kernels T_n (24 MB at nk=128, 95 MB at
// nk=256).  Building them takes ~1.3 s (5 s) of long-double FFTs on the host; the drop-in
// executable is started once per model, so it reloads them instead.  RTRG_CACHE_DIR selects the
// directory (default $XDG_CACHE_HOME/redtime_b200 or ~/.cache/redtime_b200; "off" disables).
// Bump TABLE_CACHE_VERSION whenever build_T (fastpt_tables.cc), the packing below (Tc layout,
// BIL_R, the lag window) or this header changes: a file written by an older build is then rebuilt
// instead of trusted.
enum { TABLE_CACHE_VERSION = 5 };
struct TableCacheHeader {
  char magic[8];
  int version, nk, np, nsup, n_tc, n_tlo, n_kfac, bil_r;
  double kmin, kmax, kfac_lo;
  unsigned long long checksum;  // FNV-1a over every byte of the three arrays
};
static std::string table_cache_path(const GridSpec &g) {
  const char *dir = std::getenv("RTRG_CACHE_DIR");
  std::string d;
  if (dir && *dir) {
    if (std::string(dir) == "off") return "";
    d = dir;
  } else if (const char *x = std::getenv("XDG_CACHE_HOME")) {
    d = std::string(x) + "/redtime_b200";
  } else if (const char *hme = std::getenv("HOME")) {
    d = std::string(hme) + "/.cache/redtime_b200";
  } else {
    return "";
  }
  char name[160];
  std::snprintf(name, sizeof name, "/T_v%d_nk%d_%.17g_%.17g.bin", (int)TABLE_CACHE_VERSION, g.nk, g.kmin, g.kmax);
  return d + name;
}
static unsigned long long table_checksum(const std::vector<double> &a, const std::vector<double> &b,
                                         const std::vector<double> &c) {
  // FNV-1a, 64 bit, over all bytes, eight at a time (the arrays are doubles)
  unsigned long long hsh = 1469598103934665603ULL;
  for (const auto *v : {&a, &b, &c}) {
    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(v->data());
    for (size_t i = 0; i < v->size(); i++) {
      hsh ^= w[i];
      hsh *= 1099511628211ULL;
    }
  }
  return hsh;
}
static bool load_table_cache(const std::string &path, const GridSpec &g, std::vector<double> &Tc,
                             std::vector<double> &Tlo, std::vector<double> &kfac, double *kfac_lo) {
  if (path.empty()) return false;
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  TableCacheHeader h;
  bool ok = std::fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, "RTRGTAB", 8) == 0 && h.version == TABLE_CACHE_VERSION &&
            h.bil_r == BIL_R && h.nk == g.nk && h.np == g.np && h.nsup == g.nsup && h.kmin == g.kmin && h.kmax == g.kmax &&
            (size_t)h.n_tc == Tc.size() && (size_t)h.n_tlo == Tlo.size() && (size_t)h.n_kfac == kfac.size();
  ok = ok && std::fread(Tc.data(), sizeof(double), Tc.size(), f) == Tc.size();
  ok = ok && std::fread(Tlo.data(), sizeof(double), Tlo.size(), f) == Tlo.size();
  ok = ok && std::fread(kfac.data(), sizeof(double), kfac.size(), f) == kfac.size();
  std::fclose(f);
  ok = ok && table_checksum(Tc, Tlo, kfac) == h.checksum;
  if (ok) *kfac_lo = h.kfac_lo;
  else std::fill(Tc.begin(), Tc.end(), 0.0);
  return ok;
}
static void save_table_cache(const std::string &path, const GridSpec &g, const std::vector<double> &Tc,
                             const std::vector<double> &Tlo, const std::vector<double> &kfac, double kfac_lo) {
  if (path.empty()) return;
  const std::string dir = path.substr(0, path.rfind('/'));
  for (size_t i = 1; i <= dir.size(); i++)  // mkdir -p
    if (i == dir.size() || dir[i] == '/') ::mkdir(dir.substr(0, i).c_str(), 0755);
  const std::string tmp = path + ".tmp" + std::to_string((long long)::getpid());
  FILE *f = std::fopen(tmp.c_str(), "wb");
  if (!f) return;
  TableCacheHeader h;
  std::memset(&h, 0, sizeof h);
  std::memcpy(h.magic, "RTRGTAB", 8);
  h.version = TABLE_CACHE_VERSION, h.bil_r = BIL_R, h.nk = g.nk, h.np = g.np, h.nsup = g.nsup;
  h.n_tc = (int)Tc.size(), h.n_tlo = (int)Tlo.size(), h.n_kfac = (int)kfac.size();
  h.kmin = g.kmin, h.kmax = g.kmax, h.kfac_lo = kfac_lo, h.checksum = table_checksum(Tc, Tlo, kfac);
  bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(Tc.data(), sizeof(double), Tc.size(), f) == Tc.size() &&
            std::fwrite(Tlo.data(), sizeof(double), Tlo.size(), f) == Tlo.size() &&
            std::fwrite(kfac.data(), sizeof(double), kfac.size(), f) == kfac.size();
  ok = (std::fclose(f) == 0) && ok;
  if (ok) std::rename(tmp.c_str(), path.c_str());  // atomic: concurrent writers cannot tear the file
  else std::remove(tmp.c_str());
}

template <class T>
static int dev_alloc(std::vector<void *> &pool, T **p, size_t n, bool zero = true) {
  void *q = nullptr;
  if (n == 0) n = 1;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T));
  if (e != cudaSuccess) return fail(RTRG_ENOMEM, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
  if (zero) cudaMemset(q, 0, n * sizeof(T));
  pool.push_back(q);
  *p = (T *)q;
  return RTRG_OK;
}
template <class T>
static int dev_upload(std::vector<void *> &pool, const T **p, const std::vector<T> &v) {
  T *q = nullptr;
  int rc = dev_alloc(pool, &q, v.size(), false);
  if (rc) return rc;
  if (!v.empty()) {
    cudaError_t e = cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(RTRG_ECUDA, "cudaMemcpy H2D: %s", cudaGetErrorString(e));
  }
  *p = q;
  return RTRG_OK;
}
static void free_pool(std::vector<void *> &pool) {
  for (void *p : pool) cudaFree(p);
  pool.clear();
}

static int num_columns(const rtrg_config &cfg, const Cosmo &c) {
  // rt:1670-1737
  int n = 1 + 3;
  if (c.sw_pl) n += 6;
  if (cfg.print_A) n += 14;
  if (cfg.print_I) n += 14;
  if (c.sw_pr) n += cfg.print_bias ? (5 + 9 + 8) : 7;
  if (cfg.print_Q) n += 24;
  return n;
}

extern "C" {

void rtrg_default_config(rtrg_config *cfg) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof *cfg);
  cfg->nk = 128;
  cfg->kmin = 1e-3;
  cfg->kmax = 1.0;
  cfg->z1l = 10.0;
  cfg->eps_abs = 1e-7;
  cfg->eps_rel = 1e-2;
  cfg->beta_kmin = 1e-3;
  cfg->beta_kmax = 1.0;
  cfg->n_lnk = 50;
  cfg->n_lna = 100;
  cfg->a_early = 1e-20;
  cfg->device = 0;
  cfg->max_attempts = 100000;
  cfg->k_shards = 1;
  cfg->k_rank = 0;
  cfg->v_split = 1;
}

const char *rtrg_last_error(void) { return g_err.c_str(); }
const char *rtrg_version(void) { return "redtime_b200 0.1 (sm_100a)"; }

int rtrg_create(const rtrg_config *cfg, rtrg_handle **out) {
  if (!cfg || !out) return fail(RTRG_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->nk < 16 || cfg->nk % 16 || !(cfg->kmin > 0) || !(cfg->kmax > cfg->kmin) || cfg->n_lnk < 3 ||
      cfg->n_lna < 3 || cfg->k_shards < 1 || cfg->k_rank < 0 || cfg->k_rank >= cfg->k_shards ||
      (cfg->nk / BIL_R) % cfg->k_shards || cfg->v_split < 1 || cfg->v_split > 16)
    return fail(RTRG_EINVAL, "invalid configuration");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(RTRG_ENOGPU, "no CUDA device available; redtime_b200 has no CPU path");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(RTRG_EINVAL, "device %d out of range", cfg->device);
  CU(cudaSetDevice(cfg->device));
  rtrg_handle *h = new rtrg_handle();
  h->cfg = *cfg;
  h->grid = make_grid(cfg->nk, cfg->kmin, cfg->kmax);
  std::memset(&h->S, 0, sizeof h->S);
  std::memset(&h->tb, 0, sizeof h->tb);
  {
    cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
      int least = 0, greatest = 0;
      cudaDeviceGetStreamPriorityRange(&least, &greatest);
      e = cudaStreamCreateWithPriority(&h->init_stream, cudaStreamNonBlocking, greatest);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->copy_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->side.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side.fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side.join, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      rtrg_destroy(h);
      return fail(RTRG_ECUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
    }
  }
  h->stream = h->own_stream;
  if (linear_upload_constants() != 0) {
    rtrg_destroy(h);
    return fail(RTRG_ECUDA, "constant upload failed");
  }

  const GridSpec &g = h->grid;
  const int nk = g.nk, np = g.np;
  IntegralTabs &tb = h->tb;
  tb.nk = nk;
  tb.np = np;
  tb.nshift = g.nshift;
  tb.jlo = g.jlo;
  tb.nsup = g.nsup;
  tb.nloMR = g.nloMR;
  // lag range a row block touches, in whole 8-lag DMMA tiles (nsup + BIL_R - 1 rounded up: the one
  // or two extra lags multiply zero padding of the spectra)
  tb.NV = (g.nsup + BIL_R - 1 + 7) / 8 * 8;
  tb.NVp = (tb.NV + BIL_R - 1) / BIL_R * BIL_R;
  tb.LP = tb.NVp + BIL_R;
  tb.NUp = nk - BIL_R + tb.NVp;
  tb.ldT = (nk + g.nsup - 1 + 7) / 8 * 8;
  tb.tpb = bilinear_tpb();
  tb.nchunk = (tb.NV + tb.tpb - 2) / tb.tpb + 1;
  tb.vsplit = cfg->v_split;
  tb.dlnk = g.dlnk;
  {
    // the bilinear kernel stages 9 spectra of LP samples in shared memory, two CTAs per SM
    int smem_optin = 0;
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
    const size_t need = bilinear_smem_bytes(tb) + 1024;
    if (need > (size_t)smem_optin) {
      rtrg_destroy(h);
      return fail(RTRG_EINVAL, "nk = %d needs %zu bytes of shared memory per CTA, the device offers %d (nk <= 512 fits)",
                  cfg->nk, need, smem_optin);
    }
    if (integrals_configure(tb) != 0 || ode_configure() != 0) {
      cudaGetLastError();
      rtrg_destroy(h);
      return fail(RTRG_ECUDA, "cudaFuncSetAttribute(shared memory opt-in) failed");
    }
  }

  // --- circulant kernels, built in parallel on the host, packed into the compact layout
  const int UMIN = g.nshift - (np - 1), NU = nk + g.nsup - 1;
  std::vector<double> Tc((size_t)(N_JKERN + N_TKERN) * tb.NUp * tb.ldT, 0.0), kfac((size_t)N_JKERN * nk);
  std::vector<double> Tlo((size_t)g.nsup * g.nsup);
  double kfac_lo = 0;
  const std::string cache = table_cache_path(g);
  if (!load_table_cache(cache, g, Tc, Tlo, kfac, &kfac_lo)) {
    std::vector<std::thread> th;
    for (int n = 0; n < N_JKERN; n++)
      th.emplace_back([&, n]() {
        std::vector<double> T, kf;
        build_T(g, n, T, kf);
        double *dst = &Tc[(size_t)n * tb.NUp * tb.ldT];
        double *dstT = (n >= TKERN_FIRST && n < TKERN_FIRST + N_TKERN) ? &Tc[(size_t)(N_JKERN + n - TKERN_FIRST) * tb.NUp * tb.ldT] : nullptr;
        for (int vv = 0; vv < NU; vv++) {
          const int v = ((vv + UMIN) % np + np) % np;
          for (int uu = 0; uu < NU; uu++) {
            const int u = ((uu + UMIN) % np + np) % np;
            // 8 x 8 tiles (alpha-side lags x beta-side lags) stored in the operand order of
            // DMMA.8x8x4: lane 4 g + t of a warp holds, as one 16-byte element, the beta-side lags
            // 2 t and 2 t + 1 of alpha-side lag g -- the A fragments of two consecutive MMAs
            const size_t tile = (size_t)(vv >> 3) * (tb.ldT / 8) + (uu >> 3);
            const size_t at = (tile * 32 + 4 * (uu & 7) + ((vv & 7) >> 1)) * 2 + (vv & 1);
            dst[at] = T[(size_t)u * np + v];
            if (dstT) dstT[at] = T[(size_t)v * np + u];  // the transposed copy (N_TKERN)
          }
        }
        for (int i = 0; i < nk; i++) kfac[(size_t)n * nk + i] = kf[g.nshift + i];
        if (n == 0) {
          kfac_lo = kf[g.nloMR];
          for (int ll = 0; ll < g.nsup; ll++) {
            const int v = ((g.nloMR - (np - 1 - ll)) % np + np) % np;
            for (int jj = 0; jj < g.nsup; jj++) {
              const int u = ((g.nloMR - (np - 1 - jj)) % np + np) % np;
              Tlo[(size_t)ll * g.nsup + jj] = T[(size_t)u * np + v];
            }
          }
        }
      });
    for (auto &t : th) t.join();
    save_table_cache(cache, g, Tc, Tlo, kfac, kfac_lo);
  }
  tb.kfac_lo = kfac_lo;
  std::vector<double> G((size_t)N_ZKERN * (2 * np - 1));
  for (int n = 0; n < N_ZKERN; n++) {
    std::vector<double> Gn;
    build_G(g, n, Gn);
    std::copy(Gn.begin(), Gn.end(), G.begin() + (size_t)n * (2 * np - 1));
  }
  std::vector<double> WP(np), kpad(np);
  h->kgrid.resize(nk);
  std::vector<double> lnkArr(nk);
  const double lnkmin = std::log(cfg->kmin);
  for (int i = 0; i < nk; i++) {  // rt:1559-1562
    lnkArr[i] = lnkmin + g.dlnk * i;
    h->kgrid[i] = std::exp(lnkArr[i]);
  }
  // wavenumbers of the pre-reduced beta columns: the nk grid values, then the growth-table ones,
  // clamped as Beta_P does (hdr:538-545)
  {
    const double lnk_min = std::log(GROWTH_K_MIN), dl = std::log(GROWTH_K_MAX / GROWTH_K_MIN) / cfg->n_lnk;
    for (int i = 0; i < nk; i++) h->slot_k.push_back(h->kgrid[i]);
    for (int j = 0; j <= cfg->n_lnk; j++) h->slot_k.push_back(std::exp(lnk_min + dl * j));
    for (double &k : h->slot_k) k = std::min(std::max(k, cfg->beta_kmin), cfg->beta_kmax);
  }
  // --- Pab stencil for every padded sample (rt:181-232, itp:68-78)
  std::vector<int> ex_n0;
  std::vector<double> ex_w, ex_dx;
  build_extrap_stencil(g, ex_n0, ex_w, ex_dx);
  for (int ip = 0; ip < np; ip++) {
    WP[ip] = window_P(g, ip);
    kpad[ip] = std::exp(g.lnk_pad_min + g.dlnk * ip);
  }
  // --- assembly table sorted by output row.  Kernels with alpha = beta are symmetric,
  // T_n[u][v] = T_n[v][u], so J_n(ab,cd) = J_n(cd,ab): terms are redirected to the pair with
  // ab <= cd, which lets k_bilinear skip beta-side spectra that only the mirrored pairs used
  // (e.g. the four (2,2,l) Jn0 kernels of P_T,jm need P_11 alone instead of P_01 and P_11).
  std::vector<AsmTerm> terms = assembly_terms();
  for (AsmTerm &t : terms) {
    if (t.src != 0 && t.src != 2) continue;
    const KernSpec ks = kern_spec(t.index / 9 + (t.src == 2 ? 7 : 0));
    if (ks.alpha != ks.beta) continue;
    const int ab = (t.index % 9) / 3, cd = t.index % 3;
    if (ab > cd) t.index = (short)(9 * (t.index / 9) + 3 * cd + ab);
  }
  std::stable_sort(terms.begin(), terms.end(), [](const AsmTerm &a, const AsmTerm &b) { return a.row < b.row; });
  std::vector<int> t_start(N_SRC + 1, 0);
  std::vector<short> t_src, t_index, t_kpow;
  std::vector<double> t_coef;
  for (const AsmTerm &t : terms) {
    t_start[t.row + 1]++;
    t_src.push_back(t.src);
    t_index.push_back(t.index);
    t_kpow.push_back(t.kpow);
    t_coef.push_back(t.coef);
  }
  for (int r = 0; r < N_SRC; r++) t_start[r + 1] += t_start[r];
  tb.n_terms = (int)terms.size();
  if (tb.n_terms > 768) {  // k_assemble keeps the table in shared memory (ASM_MAXT)
    rtrg_destroy(h);
    return fail(RTRG_EINVAL, "assembly table has %d terms, the kernel holds 768", tb.n_terms);
  }
  // which (kernel, beta-side spectrum) products each output group consumes
  std::memset(tb.need_cd, 0, sizeof tb.need_cd);
  std::memset(tb.need_ab, 0, sizeof tb.need_ab);
  std::memset(tb.need_pz, 0, sizeof tb.need_pz);
  std::memset(tb.need_val, 0, sizeof tb.need_val);
  for (const AsmTerm &t : terms) {
    {  // the raw value the term reads: J 0-62, PZ 63-125, Jn0 126-188, J_lo 189
      const int gi = t.row < 14 ? 0 : t.row < 38 ? 1 : t.row < 47 ? 2 : 3;
      const int v = t.src == 3 ? 189 : t.src * 63 + t.index;
      tb.need_val[gi][v >> 6] |= 1ULL << (v & 63);
    }
    if (t.src == 1) {  // PZ: index = 9 n + 3 ab + cd, the convolution itself is PZ_n(P_ab)
      const int gi = t.row < 14 ? 0 : t.row < 38 ? 1 : t.row < 47 ? 2 : 3;
      tb.need_pz[gi] |= 1u << (3 * (t.index / 9) + (t.index % 9) / 3);
    }
    if (t.src != 0 && t.src != 2) continue;  // J (kernels 0-6) and Jn0 (7-13)
    const int gi = t.row < 14 ? 0 : t.row < 38 ? 1 : t.row < 47 ? 2 : 3;
    const int n = t.index / 9 + (t.src == 2 ? 7 : 0), cd = t.index % 3;
    tb.need_cd[gi][n] |= (unsigned char)(1 << cd);
    tb.need_ab[gi][n] |= (unsigned char)(1 << ((t.index % 9) / 3));
  }

  int rc = 0;
  auto &P = h->table_allocs;
  tb.n_Tc = (long long)Tc.size();
  rc = rc ? rc : dev_upload(P, &tb.Tc, Tc);
  rc = rc ? rc : dev_upload(P, &tb.Tlo, Tlo);
  rc = rc ? rc : dev_upload(P, &tb.kfac, kfac);
  rc = rc ? rc : dev_upload(P, &tb.G, G);
  rc = rc ? rc : dev_upload(P, &tb.WP, WP);
  rc = rc ? rc : dev_upload(P, &tb.kpad, kpad);
  rc = rc ? rc : dev_upload(P, &tb.kgrid, h->kgrid);
  rc = rc ? rc : dev_upload(P, &tb.ex_n0, ex_n0);
  rc = rc ? rc : dev_upload(P, &tb.ex_w, ex_w);
  rc = rc ? rc : dev_upload(P, &tb.ex_dx, ex_dx);
  rc = rc ? rc : dev_upload(P, &tb.t_start, t_start);
  rc = rc ? rc : dev_upload(P, &tb.t_src, t_src);
  rc = rc ? rc : dev_upload(P, &tb.t_index, t_index);
  rc = rc ? rc : dev_upload(P, &tb.t_kpow, t_kpow);
  rc = rc ? rc : dev_upload(P, &tb.t_coef, t_coef);
  if (rc) {
    rtrg_destroy(h);
    return rc;
  }
  h->d_kgrid = const_cast<double *>(tb.kgrid);
  *out = h;
  return RTRG_OK;
}

int rtrg_destroy(rtrg_handle *h) {
  if (!h) return RTRG_OK;
  cudaSetDevice(h->cfg.device);
  free_pool(h->batch_allocs);
  free_pool(h->table_allocs);
  if (h->work.base) cudaFree(h->work.base);
  if (h->d_in) cudaFree(h->d_in);
  h->stage.release();
  if (h->h_out) cudaFreeHost(h->h_out);
  if (h->copy_done) cudaEventDestroy(h->copy_done);
  if (h->side.fork) cudaEventDestroy(h->side.fork);
  if (h->side.join) cudaEventDestroy(h->side.join);
  if (h->side.stream) cudaStreamDestroy(h->side.stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->init_stream) cudaStreamDestroy(h->init_stream);
  delete h;
  return RTRG_OK;
}

int rtrg_set_stream(rtrg_handle *h, void *cuda_stream) {
  if (!h) return fail(RTRG_EINVAL, "null handle");
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return RTRG_OK;
}

int rtrg_clear_cosmologies(rtrg_handle *h) {
  if (!h) return fail(RTRG_EINVAL, "null handle");
  h->cos.clear();
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  h->stage.used = 0;
  h->d_in_used = 0;
  h->d_in_transformed = false;
  h->prepared = h->uploaded = false;
  return RTRG_OK;
}

int rtrg_num_cosmologies(const rtrg_handle *h) { return h ? (int)h->cos.size() : 0; }

int rtrg_num_columns(const rtrg_handle *h, int i) {
  if (!h || i < 0 || i >= (int)h->cos.size()) return RTRG_EINVAL;
  return num_columns(h->cfg, h->cos[i].c);
}

// host threads one handle may use for staging: RTRG_HOST_THREADS, else the cores this process
// may run on divided by the ranks of the node (LOCAL_WORLD_SIZE, set by torchrun), capped
static int host_threads(int cap) {
  if (const char *v = std::getenv("RTRG_HOST_THREADS")) {
    const int n = std::atoi(v);
    if (n > 0) return std::min(n, 64);
  }
  int n = (int)std::thread::hardware_concurrency();
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) n = std::min(n, (int)CPU_COUNT(&set));
  if (const char *w = std::getenv("LOCAL_WORLD_SIZE")) {
    const int ws = std::atoi(w);
    if (ws > 1) n = n / ws;
  }
  return std::max(1, std::min(n, cap));
}

// Stream of the device work of rtrg_prepare / rtrg_device_init: the handle's high-priority stream
// (see rtrg_handle::init_stream) unless the caller supplied a stream.  Both functions end with a
// synchronisation, so what follows on h->stream is ordered by the host.
static cudaStream_t prep_stream(const rtrg_handle *h) {
  static const bool init_prio = !std::getenv("RTRG_NO_INIT_PRIORITY");
  return (init_prio && h->stream == h->own_stream && h->init_stream) ? h->init_stream : h->stream;
}

static int check_cosmology(const rtrg_cosmology *in) {
  if (!in) return fail(RTRG_EINVAL, "null cosmology");
  if (in->n_out < 1 || in->n_out > RTRG_MAX_OUT || !in->z_out) return fail(RTRG_EINVAL, "bad n_out");
  if (in->n_T < 4 || !in->k_T || !in->Tc_T || !in->Tb_T) return fail(RTRG_EINVAL, "bad transfer table");
  if (in->n_z < 0 || in->n_z > RTRG_MAX_Z || in->n_z == 1) return fail(RTRG_EINVAL, "bad n_z");
  if (in->n_z > 0 && (in->n_kb < 4 || !in->z_interp || !in->k_b || !in->Tc_b || !in->Tnu_b))
    return fail(RTRG_EINVAL, "bad interpolation tables");
  return RTRG_OK;
}
static bool is_page_locked(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}
// Scalars + layout of one cosmology.  The part the host copies (a nodes; the small tables k_T,
// Tc_T, Tb_T, k_b unless they are page-locked; beta or its pre-reduction) starts at arena offset
// `soff` = pool offset `off`; page-locked caller tables are sent directly and are placed by the
// caller afterwards (place_direct).
static HostCosmo describe_cosmology(const rtrg_handle *h, const rtrg_cosmology *in, size_t off, size_t soff) {
  HostCosmo hc;
  Cosmo &c = hc.c;
  std::memset(&c, 0, sizeof c);
  const double *p = in->params;
  c.ns = p[0], c.s8 = p[1], c.h = p[2], c.Om = p[3], c.Ob = p[4], c.On = p[5], c.TK = p[6], c.w0 = p[7], c.wa = p[8];
  c.z_in = in->z_in;
  cosmo_derive(c);
  c.sw_nl = in->switches[0], c.sw_1l = in->switches[1], c.sw_pl = in->switches[2], c.sw_pr = in->switches[3];
  c.n_out = in->n_out;
  hc.z_out.assign(in->z_out, in->z_out + in->n_out);
  c.nT = in->n_T;
  c.n_z = in->n_z > 0 ? in->n_z : 0;
  c.n_kb = c.n_z > 0 ? in->n_kb : 0;
  const size_t nT = c.nT, nz = c.n_z, nkb = c.n_kb;
  hc.stage_off = soff;
  hc.pool_off = off;
  hc.small_direct = is_page_locked(in->k_T) && is_page_locked(in->Tc_T) && is_page_locked(in->Tb_T) &&
                    (nz == 0 || is_page_locked(in->k_b));
  size_t o = off;
  if (hc.small_direct) {
    hc.dsrc[0] = in->k_T, hc.dlen[0] = nT;
    hc.dsrc[1] = in->Tc_T, hc.dlen[1] = nT;
    hc.dsrc[2] = in->Tb_T, hc.dlen[2] = nT;
    hc.dsrc[3] = in->k_b, hc.dlen[3] = nkb;
    hc.n_direct = nz ? 4 : 3;
  } else {
    c.offT = (long long)o, c.offLT = (long long)(o + nT), c.offTb = (long long)(o + 2 * nT);
    o += 3 * nT;
  }
  c.offA = (long long)o;
  o += nz;
  if (!hc.small_direct) {
    c.offKb = (long long)o;
    o += nkb;
  }
  c.offB = (long long)o;
  c.offTc = -1;
  c.offRow1 = c.offBred = -1;
  if (h->cfg.reduce_beta && nz > 0) {
    // only what the run consumes: beta(a=1, k_b) and beta at the nkk slot wavenumbers
    const size_t nkk = h->slot_k.size();
    c.offRow1 = c.offB;
    c.offBred = c.offB + (long long)nkb;
    o += nkb + nz * nkk;
  } else {
    hc.direct = nz > 0 && is_page_locked(in->Tc_b) && is_page_locked(in->Tnu_b);
    if (hc.direct) {
      hc.dsrc[hc.n_direct] = in->Tnu_b, hc.dlen[hc.n_direct++] = nz * nkb;
      hc.dsrc[hc.n_direct] = in->Tc_b, hc.dlen[hc.n_direct++] = nz * nkb;
    } else {
      o += nz * nkb;
    }
  }
  hc.stage_len = o - off;
  for (int i = 0; i < hc.n_direct; i++) hc.direct_len += hc.dlen[i];
  return hc;
}
// pool offsets of the direct segments, starting at off_direct
static void place_direct(HostCosmo &hc, size_t off_direct) {
  Cosmo &c = hc.c;
  size_t o = off_direct;
  int i = 0;
  if (hc.small_direct) {
    hc.doff[0] = c.offT = (long long)o, o += hc.dlen[0];
    hc.doff[1] = c.offLT = (long long)o, o += hc.dlen[1];
    hc.doff[2] = c.offTb = (long long)o, o += hc.dlen[2];
    i = 3;
    if (c.n_z > 0) hc.doff[3] = c.offKb = (long long)o, o += hc.dlen[3], i = 4;
  }
  if (hc.direct) {
    hc.doff[i] = c.offB = (long long)o, o += hc.dlen[i], i++;
    hc.doff[i] = c.offTc = (long long)o, o += hc.dlen[i], i++;
  }
}
// copy what the host has to touch into the staging arena; pageable interpolation tables are
// reduced to beta = f_nu T_nu / T_c on the way (hdr:556-623), which halves their PCIe bytes
static void stage_cosmology(const rtrg_handle *h, const rtrg_cosmology *in, const HostCosmo &hc, double *base) {
  const Cosmo &c = hc.c;
  const size_t nT = c.nT, nz = c.n_z, nkb = c.n_kb;
  double *s = base + hc.stage_off;
  if (!hc.small_direct) {
    std::memcpy(s, in->k_T, nT * sizeof(double));
    std::memcpy(s + nT, in->Tc_T, nT * sizeof(double));
    std::memcpy(s + 2 * nT, in->Tb_T, nT * sizeof(double));
    s += 3 * nT;
  }
  double *a_nodes = s;
  for (size_t i = 0; i < nz; i++) s[i] = 1.0 / (1.0 + in->z_interp[i]);
  s += nz;
  if (nz && !hc.small_direct) {
    std::memcpy(s, in->k_b, nkb * sizeof(double));
    s += nkb;
  }
  if (nz && c.offRow1 >= 0) {
    // pre-reduction of the beta table: the 2-D rule interpolates in a first, column
    // by column, and then in k (tab:262-328), so (i) the row at a = 1 and (ii) the k-stencil
    // applied to every a row are all the run needs.  Same arithmetic as beta_P / k_beta_reduce.
    const double fn = c.On / c.Om;
    const double *tn = in->Tnu_b, *tc = in->Tc_b, *a = a_nodes, *kb = in->k_b;
    auto beta = [&](size_t j, size_t i) { return fn * tn[j * nkb + i] / tc[j * nkb + i]; };
    double *row1 = s, *bred = row1 + nkb;
    const int X = (int)nz, nx = tab_find(a, X, 1.0);
    // (vectorised, same bits as cub4 / lin2 per column: host_stage.cc)
    if (nx > 0 && nx < X - 2)
      beta_row_cubic(tn + (size_t)(nx - 1) * nkb, tc + (size_t)(nx - 1) * nkb, nkb, fn, a + nx - 1, 1.0, row1);
    else
      beta_row_linear(tn + (size_t)nx * nkb, tc + (size_t)nx * nkb, nkb, fn, a[nx], a[nx + 1], 1.0, row1);
    const size_t nkk = h->slot_k.size();
    for (size_t kk = 0; kk < nkk; kk++) {
      const Stencil st = tab_stencil_y(kb, (int)nkb, h->slot_k[kk]);
      for (size_t j = 0; j < nz; j++) {
        double r = 0;
        for (int m = 0; m < 4; m++)
          if (st.w[m] != 0.0) r += st.w[m] * beta(j, (size_t)(st.n0 + m));
        bred[j * nkk + kk] = r;
      }
    }
    return;
  }
  if (nz && !hc.direct) {
    const double fn = c.On / c.Om;
    const double *tn = in->Tnu_b, *tc = in->Tc_b;
    const size_t n = nz * nkb;
    for (size_t i = 0; i < n; i++) s[i] = fn * tn[i] / tc[i];
  }
}
// host -> device copies of the direct segments of one cosmology, straight from the caller's
// page-locked buffers
static int upload_direct(rtrg_handle *h, const HostCosmo &hc) {
  for (int i = 0; i < hc.n_direct; i++)
    if (hc.dlen[i])
      CU(cudaMemcpyAsync(h->d_in + hc.doff[i], hc.dsrc[i], hc.dlen[i] * sizeof(double), cudaMemcpyHostToDevice,
                         h->copy_stream));
  return RTRG_OK;
}
// (re-)send everything of the cosmologies [b0, b1)
static int upload_range(rtrg_handle *h, size_t b0, size_t b1) {
  if (b0 >= b1) return RTRG_OK;
  // staged parts: each add call left one contiguous slice in the arena and in the pool
  size_t i = b0;
  while (i < b1) {
    size_t j = i, len = 0;
    while (j < b1 && h->cos[j].pool_off == h->cos[i].pool_off + len && h->cos[j].stage_off == h->cos[i].stage_off + len)
      len += h->cos[j++].stage_len;
    if (len)
      CU(cudaMemcpyAsync(h->d_in + h->cos[i].pool_off, h->stage.base + h->cos[i].stage_off, len * sizeof(double),
                         cudaMemcpyHostToDevice, h->copy_stream));
    i = j;
  }
  for (size_t b = b0; b < b1; b++) {
    int rc = upload_direct(h, h->cos[b]);
    if (rc) return rc;
  }
  return RTRG_OK;
}

int rtrg_add_cosmologies(rtrg_handle *h, int n, const rtrg_cosmology *const *list) {
  if (!h || n < 0 || (n > 0 && !list)) return fail(RTRG_EINVAL, "null argument");
  for (int i = 0; i < n; i++) {
    int rc = check_cosmology(list[i]);
    if (rc) return rc;
  }
  CU(cudaSetDevice(h->cfg.device));
  const size_t first = h->cos.size();
  const size_t d_in_used0 = h->d_in_used, stage_used0 = h->stage.used;
  auto roll_back = [&]() {  // leave the handle as it was before this call
    h->cos.resize(first);
    h->d_in_used = d_in_used0;
    h->stage.used = stage_used0;
  };
  // layout: the small tables of this call form one slice (mirrored in the arena), the big raw
  // tables of the direct cosmologies follow it
  size_t off = h->d_in_used, soff = h->stage.used;
  for (int i = 0; i < n; i++) {
    h->cos.push_back(describe_cosmology(h, list[i], off, soff));
    off += h->cos.back().stage_len;
    soff += h->cos.back().stage_len;
  }
  for (int i = 0; i < n; i++) {
    place_direct(h->cos[first + i], off);
    off += h->cos[first + i].direct_len;
  }
  h->prepared = h->uploaded = false;
  if (soff > h->stage.cap && cudaStreamSynchronize(h->copy_stream) != cudaSuccess) {  // the arena is about to move
    roll_back();
    return fail(RTRG_ECUDA, "cudaStreamSynchronize(copy stream): %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (h->stage.reserve(soff) != RTRG_OK) {
    roll_back();
    return fail(RTRG_ENOMEM, "pinned staging arena of %zu bytes", soff * sizeof(double));
  }
  h->stage.used = soff;
  h->d_in_used = off;
  if (h->d_in_used > h->d_in_cap) {
    if (cudaStreamSynchronize(h->copy_stream) != cudaSuccess) {
      roll_back();
      return fail(RTRG_ECUDA, "cudaStreamSynchronize(copy stream): %s", cudaGetErrorString(cudaGetLastError()));
    }
    double *q = nullptr;
    const size_t ncap = h->d_in_used + h->d_in_used / 8;
    cudaError_t e = cudaMalloc((void **)&q, ncap * sizeof(double));
    if (e != cudaSuccess) {
      roll_back();
      return fail(RTRG_ENOMEM, "cudaMalloc(%zu bytes): %s", ncap * sizeof(double), cudaGetErrorString(e));
    }
    if (h->d_in) cudaFree(h->d_in);
    h->d_in = q;
    h->d_in_cap = ncap;
    int rc = upload_range(h, 0, first);  // the pool moved: what was sent before goes out again
    if (rc) {
      roll_back();
      return rc;
    }
  }
  // Staging runs on several host threads (the copies are memory-bound, one thread moves only
  // ~10 GB/s); every chunk is sent as soon as it is staged, so the PCIe transfer overlaps the
  // staging of the next chunk.  The big tables of page-locked cosmologies are not touched by
  // the host at all: the copy engine reads them from the caller's buffers.
  int nth = host_threads(16);
  nth = std::max(1, std::min(nth, n));
  double *base = h->stage.base;
  const int chunk = std::max(nth, 64);
  for (int c0 = 0; c0 < n; c0 += chunk) {
    const int c1 = std::min(n, c0 + chunk);
    if (nth == 1) {
      for (int i = c0; i < c1; i++) stage_cosmology(h, list[i], h->cos[first + i], base);
    } else {
      std::vector<std::thread> th;
      for (int t = 0; t < nth; t++)
        th.emplace_back([=]() {
          for (int i = c0 + t; i < c1; i += nth) stage_cosmology(h, list[i], h->cos[first + i], base);
        });
      for (auto &t : th) t.join();
    }
    int rc = upload_range(h, first + c0, first + c1);
    if (rc) {
      cudaStreamSynchronize(h->copy_stream);
      roll_back();
      return rc;
    }
  }
  h->prepared = h->uploaded = false;
  return RTRG_OK;
}

int rtrg_add_cosmology(rtrg_handle *h, const rtrg_cosmology *in) { return rtrg_add_cosmologies(h, 1, &in); }

int rtrg_prepare(rtrg_handle *h) {
  if (!h) return fail(RTRG_EINVAL, "null handle");
  if (h->cos.empty()) return fail(RTRG_EINVAL, "no cosmologies");
  if (h->cos.size() > 65535)  // the batch index is a grid y/z dimension of the kernels
    return fail(RTRG_EINVAL, "%zu cosmologies in one batch; the limit is 65535 (split the batch)", h->cos.size());
  CU(cudaSetDevice(h->cfg.device));
  free_pool(h->batch_allocs);  // hook scratch buffers of the previous batch
  h->prepared = h->uploaded = false;
  const rtrg_config &cfg = h->cfg;
  const int B = (int)h->cos.size(), nk = cfg.nk, np = 4 * nk;
  Batch &S = h->S;
  std::memset(&S, 0, sizeof S);
  S.B = B, S.nk = nk, S.np = np, S.n_lna = cfg.n_lna, S.n_lnk = cfg.n_lnk;
  S.nkk = nk + cfg.n_lnk + 1;
  S.eps_abs = cfg.eps_abs, S.eps_rel = cfg.eps_rel, S.z1l = cfg.z1l;
  S.beta_kmin = cfg.beta_kmin, S.beta_kmax = cfg.beta_kmax, S.a_early = cfg.a_early;
  S.print_A = cfg.print_A, S.print_I = cfg.print_I, S.print_Q = cfg.print_Q, S.print_bias = cfg.print_bias;
  const int rows_per = nk / cfg.k_shards;
  S.k_lo = cfg.k_rank * rows_per;
  S.k_hi = S.k_lo + rows_per;

  // --- per-cosmology scalars
  std::vector<Cosmo> cs(B);
  std::vector<double> zout((size_t)B * MAX_OUT, 0.0), aout((size_t)B * MAX_OUT, 1.0), etaout((size_t)B * MAX_OUT, 0.0);
  h->out_off.assign(B, 0);
  h->ncols.assign(B, 0);
  size_t off = 0;
  int n_zmax = 2, max_rows = 1;
  h->any_full = h->any_1loop = h->any_pr = h->any_local = false;
  for (int b = 0; b < B; b++) {
    const HostCosmo &hc = h->cos[b];
    const Cosmo &c = hc.c;
    n_zmax = std::max(n_zmax, c.n_z);
    max_rows = std::max(max_rows, c.nT + c.n_z * c.n_kb);
    for (int i = 0; i < c.n_out; i++) {  // hdr:274-277
      const double z = hc.z_out[i], a = 1.0 / (1.0 + z);
      zout[(size_t)b * MAX_OUT + i] = z;
      aout[(size_t)b * MAX_OUT + i] = a;
      etaout[(size_t)b * MAX_OUT + i] = std::log(a / c.a_in);
    }
    // The reference abort()s on look-ups outside its tables: D_dD needs 1e-3 <= a <= 1.1
    // (hdr:644-649) at z_in, z1l and every output, Beta_P needs a <= 1.001 (hdr:528-531).  Such a
    // cosmology is flagged and left out of the run; the rest of the batch is unaffected.
    {
      bool ok = c.a_in >= GROWTH_A_MIN && c.a_in <= GROWTH_A_MAX;
      if (c.sw_nl && c.sw_1l) ok = ok && (1.0 / (1.0 + cfg.z1l) >= GROWTH_A_MIN);
      for (int i = 0; i < c.n_out; i++) {
        const double a = 1.0 / (1.0 + hc.z_out[i]);
        ok = ok && a >= c.a_in && a <= (c.n_z > 0 && c.fnu >= 1e-10 ? 1.001 : GROWTH_A_MAX);
        if (i > 0) ok = ok && hc.z_out[i] <= hc.z_out[i - 1];  // greatest to least (hdr:274)
      }
      if (!ok) cs[b].status = RTRG_RANGE_FAIL;
    }
    h->ncols[b] = num_columns(cfg, c);
    h->out_off[b] = (long long)off;
    off += (size_t)c.n_out * nk * h->ncols[b];
    if (c.sw_nl && !c.sw_1l) h->any_full = true;
    else h->any_local = true;
    if (c.sw_nl && c.sw_1l) h->any_1loop = true;
    if (c.sw_pr) h->any_pr = true;
    const int st_keep = cs[b].status;
    cs[b] = c;
    cs[b].status = st_keep;
    h->cos[b].c.status = st_keep;
  }
  h->out_total = off;
  S.n_zmax = n_zmax;
  // virtual cosmologies of the deferred output stage: one per (cosmology, output redshift)
  std::vector<int> vbase(B), vc_b, vc_io;
  for (int b = 0; b < B; b++) {
    vbase[b] = (int)vc_b.size();
    for (int io = 0; io < h->cos[b].c.n_out; io++) vc_b.push_back(b), vc_io.push_back(io);
  }
  S.NO = (int)vc_b.size();
  h->vch = std::min(S.NO, 8192);
  const size_t BV = (size_t)std::max(B, h->vch);  // integral work space: per cosmology or per virtual one
  // growth-table axes (hdr:677-687)
  std::vector<double> lna(cfg.n_lna + 1), lnkg(cfg.n_lnk + 1);
  {
    const double lna_min = std::log(GROWTH_A_MIN), dlna = std::log(GROWTH_A_MAX / GROWTH_A_MIN) / cfg.n_lna;
    const double lnk_min = std::log(GROWTH_K_MIN), dlnk = std::log(GROWTH_K_MAX / GROWTH_K_MIN) / cfg.n_lnk;
    for (int i = 0; i <= cfg.n_lna; i++) lna[i] = lna_min + dlna * i;
    for (int j = 0; j <= cfg.n_lnk; j++) lnkg[j] = lnk_min + dlnk * j;
  }

  // --- carve the work arena (dry run for the size first)
  const size_t NE = (size_t)B * N_U * nk, ng = (size_t)(cfg.n_lna + 1) * (cfg.n_lnk + 1);
  const IntegralTabs &tb = h->tb;
  double *d_lna = nullptr, *d_lnkg = nullptr;
  // k-sharding: segment tables of the two gathers.  [0] the three ln P components of a state vector
  // (segment = one component of one cosmology, nk doubles, rank r owns rows [r nk/G, (r+1) nk/G));
  // [1] the output tables (segment = the nk x ncols block of one output, rank r owns its rows)
  const bool sharded_cfg = cfg.k_shards > 1;
  std::vector<long long> plan_off[2], plan_pre[2];
  std::vector<int> plan_len[2];
  long long plan_total[2] = {0, 0};
  if (sharded_cfg) {
    for (int b = 0; b < B; b++)
      for (int c = 0; c < N_UP; c++) {
        plan_off[0].push_back(((long long)b * N_U + c) * nk);
        plan_len[0].push_back(rows_per);
      }
    for (int b = 0; b < B; b++)
      for (int io = 0; io < h->cos[b].c.n_out; io++) {
        plan_off[1].push_back(h->out_off[b] + (long long)io * nk * h->ncols[b]);
        plan_len[1].push_back(rows_per * h->ncols[b]);
      }
    for (int p = 0; p < 2; p++)
      for (int l : plan_len[p]) {
        plan_pre[p].push_back(plan_total[p]);
        plan_total[p] += l;
      }
  }
  const size_t n_seg[2] = {plan_off[0].size(), plan_off[1].size()};
  long long *d_plan_off[2] = {nullptr, nullptr}, *d_plan_pre[2] = {nullptr, nullptr};
  int *d_plan_len[2] = {nullptr, nullptr};
  auto carve = [&](DeviceArena &A) {
    A.used = 0;
    S.cosmo = A.take<Cosmo>(B);
    S.zout = A.take<double>((size_t)B * MAX_OUT);
    S.aout = A.take<double>((size_t)B * MAX_OUT);
    S.etaout = A.take<double>((size_t)B * MAX_OUT);
    d_lna = A.take<double>(lna.size());
    d_lnkg = A.take<double>(lnkg.size());
    S.out_off = A.take<long long>(B);
    S.ncols = A.take<int>(B);
    h->d_T0 = A.take<double>(B);
    S.bred = A.take<double>((size_t)B * n_zmax * S.nkk);
    S.G = A.take<double>(B * ng);
    S.dD = A.take<double>(B * ng);
    S.Dnorm = A.take<double>((size_t)B * (cfg.n_lnk + 1));
    S.Grow = A.take<double>((size_t)B * (cfg.n_lna + 1) * nk);
    S.dDrow = A.take<double>((size_t)B * (cfg.n_lna + 1) * nk);
    S.D0row = A.take<double>((size_t)B * nk);
    S.Tgrid = A.take<double>((size_t)B * nk);
    S.src_z1l = A.take<double>((size_t)B * N_SRC * nk);
    S.D_z1l = A.take<double>((size_t)B * nk);
    S.y_z1l = A.take<double>((size_t)B * 3 * nk);
    S.y = A.take<double>(NE);
    S.ytmp = A.take<double>(NE);
    S.ynew = A.take<double>(NE);
    S.yerr = A.take<double>(NE);
    S.kst = A.take<double>(RK_STAGES * NE);
    h->d_yinit = A.take<double>(NE);
    S.src = A.take<double>((size_t)B * N_SRC * nk);
    S.Prev = A.take<double>(BV * 3 * tb.LP);
    S.P3 = A.take<double>(BV * 3 * np);
    S.Jpart = A.take<double>(BV * N_JKERN * tb.nchunk * tb.vsplit * 9 * nk);
    S.PZb = A.take<double>(BV * N_ZKERN * 3 * nk);
    S.Jlo = A.take<double>(BV);
    S.t = A.take<double>(B);
    S.h = A.take<double>(B);
    S.h_try = A.take<double>(B);
    S.rmax_bits = A.take<unsigned long long>(B);
    S.i_out = A.take<int>(B);
    S.flag_out = A.take<int>(B);
    S.flag_step = A.take<int>(B);
    S.flag_acc = A.take<int>(B);
    S.final_step = A.take<int>(B);
    S.done = A.take<int>(B);
    S.m_full_step = A.take<int>(B);
    S.m_full_acc = A.take<int>(B);
    S.m_out_int = A.take<int>(B);
    S.m_loc_step = A.take<int>(B);
    S.att_time = A.take<double>((size_t)B * RK_STAGES * 48);  // RhsShared is < 48 doubles
    S.counters = A.take<long long>((size_t)4 * B);
    S.rmax_hist = A.take<double>((size_t)B * RMAX_HIST);
    S.matvecs = A.take<long long>(B);
    S.act = A.take<int>(BV);
    S.nact = A.take<int>(1);
    S.n_active = A.take<int>(1);
    S.rounds = A.take<long long>(1);
    S.ystash = A.take<double>((size_t)S.NO * N_U * nk);
    S.t_stash = A.take<double>(S.NO);
    S.vbase = A.take<int>(B);
    S.vc_b = A.take<int>(S.NO);
    S.vc_io = A.take<int>(S.NO);
    S.vc_have = A.take<int>(S.NO);
    S.vc_mask = A.take<int>(S.NO);
    S.cosmo_v = A.take<Cosmo>(S.NO);
    S.matvecs_v = A.take<long long>(S.NO);
    S.src_v = A.take<double>((size_t)h->vch * N_SRC * nk);
    // k-shard gather plans (segment tables)
    d_plan_off[0] = A.take<long long>(n_seg[0]), d_plan_len[0] = A.take<int>(n_seg[0]), d_plan_pre[0] = A.take<long long>(n_seg[0]);
    d_plan_off[1] = A.take<long long>(n_seg[1]), d_plan_len[1] = A.take<int>(n_seg[1]), d_plan_pre[1] = A.take<long long>(n_seg[1]);
    S.out = A.take<double>(h->out_total);
    S.hdr = A.take<double>((size_t)B * MAX_OUT * 5);
    S.hdr0 = A.take<double>((size_t)B * 2);
    h->d_hookmask = A.take<int>(B);
    h->d_minit = A.take<int>(B);
    h->d_err = A.take<int>(1);
  };
  {
    DeviceArena dry;
    carve(dry);
    const size_t need = dry.used + 256;
    if (need > h->work.cap) {
      if (h->work.base) cudaFree(h->work.base);
      h->work.base = nullptr;
      h->work.cap = 0;
      void *q = nullptr;
      const size_t ncap = need + need / 8;
      cudaError_t e = cudaMalloc(&q, ncap);
      if (e != cudaSuccess) return fail(RTRG_ENOMEM, "cudaMalloc(%zu bytes): %s", ncap, cudaGetErrorString(e));
      h->work.base = (char *)q;
      h->work.cap = ncap;
    }
    carve(h->work);
  }
  h->d_raw = nullptr;
  h->d_scratch = nullptr;
  h->scratch_len = 0;
  S.lna = d_lna;
  S.lnkg = d_lnkg;

  cudaStream_t st = prep_stream(h);
  // zero everything once (the Prev padding, the Q/I state rows ... rely on it)
  CU(cudaMemsetAsync(h->work.base, 0, h->work.used, st));
  // --- input pool: rtrg_add_cosmologies already sent the staged tables on the copy stream.  A
  // repeated rtrg_prepare (the log transform below consumed the raw columns) sends them again.
  if (h->d_in_transformed) {
    int rc = upload_range(h, 0, h->cos.size());
    if (rc) return rc;
  }
  CU(cudaEventRecord(h->copy_done, h->copy_stream));
  CU(cudaStreamWaitEvent(st, h->copy_done, 0));
  h->d_in_transformed = true;
  S.in = h->d_in;
  S.n_in = (long long)h->d_in_used;
  S.n_out_total = (long long)h->out_total;
  S.n_slots = (long long)BV;
  CU(cudaMemcpyAsync(S.cosmo, cs.data(), B * sizeof(Cosmo), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(S.zout, zout.data(), zout.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(S.aout, aout.data(), aout.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(S.etaout, etaout.data(), etaout.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_lna, lna.data(), lna.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(d_lnkg, lnkg.data(), lnkg.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(S.out_off, h->out_off.data(), B * sizeof(long long), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(S.ncols, h->ncols.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(S.vbase, vbase.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(S.vc_b, vc_b.data(), vc_b.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(S.vc_io, vc_io.data(), vc_io.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  if (sharded_cfg) {
    for (int p = 0; p < 2; p++) {
      CU(cudaMemcpyAsync(d_plan_off[p], plan_off[p].data(), n_seg[p] * sizeof(long long), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(d_plan_len[p], plan_len[p].data(), n_seg[p] * sizeof(int), cudaMemcpyHostToDevice, st));
      CU(cudaMemcpyAsync(d_plan_pre[p], plan_pre[p].data(), n_seg[p] * sizeof(long long), cudaMemcpyHostToDevice, st));
      GatherPlan &gp = p ? h->plan_out : h->plan_lnP;
      gp.nseg = (int)n_seg[p], gp.total = plan_total[p];
      gp.off = d_plan_off[p], gp.len = d_plan_len[p], gp.prefix = d_plan_pre[p];
    }
    if (h->xch) {
      std::string xerr;
      if (h->xch->reserve(std::max(plan_total[0], plan_total[1]), (size_t)B, &xerr) != 0)
        return fail(RTRG_ENOMEM, "%s", xerr.c_str());
    }
  }
  h->launches += launch_prep_inputs(S, h->d_T0, max_rows, st, h->prof);
  CU(cudaStreamSynchronize(st));  // the host vectors above go out of scope
  h->uploaded = true;
  return rtrg_device_init(h);
}

int rtrg_device_init(rtrg_handle *h) {
  if (!h) return fail(RTRG_EINVAL, "null handle");
  if (!h->uploaded) return fail(RTRG_EINVAL, "rtrg_prepare() has not uploaded any cosmology");
  CU(cudaSetDevice(h->cfg.device));
  h->prepared = false;
  Batch &S = h->S;
  const IntegralTabs &tb = h->tb;
  const int B = S.B, nk = S.nk;
  const size_t NE = (size_t)B * N_U * nk;
  // (everything issued before is complete: rtrg_prepare synchronised; this function synchronises at its end)
  cudaStream_t st = prep_stream(h);
  CU(cudaMemsetAsync(S.matvecs, 0, B * sizeof(long long), st));
  CU(cudaMemsetAsync(S.matvecs_v, 0, S.NO * sizeof(long long), st));
  h->launches += launch_linear_init(S, h->d_kgrid, st, h->prof);
  CU(cudaMemcpyAsync(h->d_yinit, S.y, NE * sizeof(double), cudaMemcpyDeviceToDevice, st));
  if (h->any_1loop) {
    // 1-loop cache: the integrals of the linear spectrum at z1l (rt:1295-1313)
    std::vector<int> m(B);
    for (int b = 0; b < B; b++) m[b] = h->cos[b].c.sw_nl && h->cos[b].c.sw_1l;
    CU(cudaMemcpyAsync(h->d_minit, m.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
    h->launches += launch_integrals(tb, S, S.y_z1l, 3LL * nk, S.src_z1l, nullptr, h->d_minit, grp_rhs(h),
                                    /*identical=*/1, st, h->prof);
  }
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  prof_collect(h);
  // status of the device-side initialisation
  std::vector<Cosmo> cs(B);
  CU(cudaMemcpy(cs.data(), S.cosmo, B * sizeof(Cosmo), cudaMemcpyDeviceToHost));
  for (int b = 0; b < B; b++) h->cos[b].c.Norm = cs[b].Norm, h->cos[b].c.sigv2_0 = cs[b].sigv2_0, h->cos[b].c.status = cs[b].status;
  h->prepared = true;
  return RTRG_OK;
}

int rtrg_run(rtrg_handle *h, double *out, size_t out_len, double *hdr, double *hdr0, int *status) {
  if (!h) return fail(RTRG_EINVAL, "null handle");
  if (!h->prepared) return fail(RTRG_EINVAL, "rtrg_prepare() has not been called");
  if (out && out_len < h->out_total) return fail(RTRG_EINVAL, "output buffer too small (%zu < %zu)", out_len, h->out_total);
  CU(cudaSetDevice(h->cfg.device));
  Batch &S = h->S;
  const IntegralTabs &tb = h->tb;
  const int B = S.B, nk = S.nk;
  const size_t NE = (size_t)B * N_U * nk;
  cudaStream_t st = h->stream;

  // --- reset the integrator state (rt:1598-1599): eta = 0, deta = 1e-2 (eta_fin - eta)
  std::vector<double> h0(B);
  for (int b = 0; b < B; b++) h0[b] = 1e-2 * (std::log(1.0 / h->cos[b].c.a_in) - 0.0);
  CU(cudaMemcpyAsync(S.y, h->d_yinit, NE * sizeof(double), cudaMemcpyDeviceToDevice, st));
  CU(cudaMemsetAsync(S.t, 0, B * sizeof(double), st));
  CU(cudaMemcpyAsync(S.h, h0.data(), B * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(S.i_out, 0, B * sizeof(int), st));
  std::vector<int> done0(B);
  int n_active = 0;
  for (int b = 0; b < B; b++) {
    done0[b] = h->cos[b].c.status != 0;  // flagged by rtrg_prepare / the device-side initialisation
    n_active += !done0[b];
  }
  CU(cudaMemcpyAsync(S.done, done0.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(S.counters, 0, 4 * B * sizeof(long long), st));
  CU(cudaMemcpyAsync(S.n_active, &n_active, sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(S.rounds, 0, sizeof(long long), st));
  CU(cudaMemsetAsync(S.vc_have, 0, S.NO * sizeof(int), st));
  CU(cudaMemsetAsync(S.hdr, 0, (size_t)B * MAX_OUT * 5 * sizeof(double), st));  // a = 0 marks "never produced"

  // --- k-sharded mode: all-gather of the ranks' ln P rows, max-reduction of the error norm
  const bool sharded = h->cfg.k_shards > 1;
  if (sharded && (!h->xch || h->xch->nranks() != h->cfg.k_shards || h->xch->rank() != h->cfg.k_rank))
    return fail(RTRG_EINVAL, "k_shards > 1 needs rtrg_kshard_init_nccl() or rtrg_kshard_init_loopback()");
  std::string xerr;
  if (sharded && h->xch->reserve(std::max(h->plan_lnP.total, h->plan_out.total), (size_t)B, &xerr) != 0)
    return fail(RTRG_ENOMEM, "%s", xerr.c_str());
  auto gather_lnP = [&](double *yv) -> int { return sharded ? h->xch->gather(yv, h->plan_lnP, true, st, &xerr) : 0; };
  // a rank that leaves with an error wakes the ranks that would wait for it (loopback transport)
#define XCH(call)                                                             \
  do {                                                                        \
    RT_TIC(sharded ? h->prof : nullptr, PC_XCH, st);                          \
    const int xrc_ = (call);                                                  \
    RT_TOC(sharded ? h->prof : nullptr, st);                                  \
    if (xrc_ != 0) {                                                          \
      h->xch->abort();                                                        \
      return fail(RTRG_ECUDA, "k-shard exchange: %s", xerr.c_str());          \
    }                                                                         \
  } while (0)

  // small launches (one cosmology, k-sharded ranks): k_stage_post runs the assembly, the right-hand side
  // and the combination of the next stage (or the error estimate) as one kernel
  const bool fused = stage_post_applies(S) && !std::getenv("RTRG_NO_STAGE_FUSION");
  // --- dydt_in of the first step
  if (h->any_full) {
    std::vector<int> m(B);
    for (int b = 0; b < B; b++) m[b] = h->cos[b].c.sw_nl && !h->cos[b].c.sw_1l;
    CU(cudaMemcpyAsync(h->d_minit, m.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
    if (fused) {
      h->launches += launch_integrals(tb, S, S.y, (long long)N_U * nk, S.src, nullptr, h->d_minit, grp_rhs(h), 0, st, h->prof, &h->side, false);
      ODE_LAUNCH(PC_STAGE, launch_stage_post(tb, S, h->d_kgrid, S.y, -1, 0, h->d_minit, grp_rhs(h), st));
    } else {
      h->launches += launch_integrals(tb, S, S.y, (long long)N_U * nk, S.src, nullptr, h->d_minit, grp_rhs(h), 0, st, h->prof, &h->side);
      ODE_LAUNCH(PC_RHS, launch_rhs(S, h->d_kgrid, S.y, S.kst, -1, h->d_minit, st));
    }
  }

  long long rounds = 0;
  const long long max_rounds = (long long)h->cfg.max_attempts + RTRG_MAX_OUT + 8;
  // One round = every unfinished cosmology attempts one RKF45 step or reaches one output
  // redshift (its state is stashed; the tables are produced after the loop).  The launch
  // sequence of a round is the same every time -- what differs lives in device-side masks -- so
  // it is captured once.  Unsharded runs put it into the body of a conditional WHILE graph node
  // whose condition (cosmologies still active) is set by the last kernel of the body: the whole
  // evolution is ONE graph launch and one host synchronisation.  k-sharded runs over NCCL replay a
  // per-round graph (NCCL's capture may add host nodes, which a conditional body cannot hold) and
  // read one int back per round.
  auto round_body = [&]() -> int {
    ODE_LAUNCH(PC_CTRL, launch_ctrl_begin(S, st));
    ODE_LAUNCH(PC_OUTPUT, launch_stash(S, st));
    // cosmologies without integrals inside the RHS (1-loop, linear): the whole attempt in one kernel
    if (h->any_local) {
      ODE_LAUNCH(PC_ATTEMPT, launch_attempt_local(S, h->d_kgrid, S.m_loc_step, st));
      h->launches++;  // k_attempt_setup + k_attempt_local
    }
    // full Time-RG: the stages are separated by integral evaluations (and the k-shard exchange)
    if (h->any_full && fused) {
      ODE_LAUNCH(PC_COMBINE, launch_combine(S, 1, S.m_full_step, st));
      for (int s = 1; s < RK_STAGES; s++) {
        XCH(gather_lnP(S.ytmp));
        h->launches += launch_integrals(tb, S, S.ytmp, (long long)N_U * nk, S.src, nullptr, S.m_full_step, grp_rhs(h), 0, st, h->prof, &h->side, false);
        // sources, right-hand side of stage s, then ytmp of stage s + 1 (after the last stage: the
        // 5th-order solution and the error norm)
        ODE_LAUNCH(PC_STAGE, launch_stage_post(tb, S, h->d_kgrid, S.ytmp, s, s + 1, S.m_full_step, grp_rhs(h), st));
      }
    } else if (h->any_full) {
      for (int s = 1; s < RK_STAGES; s++) {
        ODE_LAUNCH(PC_COMBINE, launch_combine(S, s, S.m_full_step, st));
        XCH(gather_lnP(S.ytmp));
        h->launches += launch_integrals(tb, S, S.ytmp, (long long)N_U * nk, S.src, nullptr, S.m_full_step, grp_rhs(h), 0, st, h->prof, &h->side);
        ODE_LAUNCH(PC_RHS, launch_rhs(S, h->d_kgrid, S.ytmp, S.kst + (size_t)s * NE, s, S.m_full_step, st));
      }
      ODE_LAUNCH(PC_FINAL, launch_final(S, S.m_full_step, st));
    }
    if (sharded) XCH(h->xch->allreduce_max_u64(S.rmax_bits, B, st, &xerr));
    ODE_LAUNCH(PC_CTRL, launch_ctrl_end(S, h->cfg.max_attempts, st));
    ODE_LAUNCH(PC_ACCEPT, launch_accept(S, st));
    XCH(gather_lnP(S.y));  // keep ln P of the accepted state complete on every rank
    if (h->any_full && fused) {  // dydt_in of the next attempt
      h->launches += launch_integrals(tb, S, S.y, (long long)N_U * nk, S.src, nullptr, S.m_full_acc, grp_rhs(h), 0, st, h->prof, &h->side, false);
      ODE_LAUNCH(PC_STAGE, launch_stage_post(tb, S, h->d_kgrid, S.y, -1, 0, S.m_full_acc, grp_rhs(h), st));
    } else if (h->any_full) {
      h->launches += launch_integrals(tb, S, S.y, (long long)N_U * nk, S.src, nullptr, S.m_full_acc, grp_rhs(h), 0, st, h->prof, &h->side);
      ODE_LAUNCH(PC_RHS, launch_rhs(S, h->d_kgrid, S.y, S.kst, -1, S.m_full_acc, st));
    }
    return RTRG_OK;
  };
  // no graph while per-kernel timing is on (events between the nodes), for the host-driven
  // loopback exchange, or when RTRG_NO_GRAPH is set
  const bool use_graph = !h->prof && (!sharded || h->xch->capturable()) && !std::getenv("RTRG_NO_GRAPH");
  const bool use_while = use_graph && (!sharded || h->xch->device_side()) && !std::getenv("RTRG_NO_WHILE");
  cudaGraphExec_t gexec = nullptr;
  long long launches_per_round = 0;
  bool whole_loop_graph = false;
  if (use_while && n_active > 0) {
    // graph = { WHILE(cond) { round ; cond = n_active > 0 && rounds < max } }
    cudaGraph_t graph = nullptr, body = nullptr;
    cudaGraphConditionalHandle cond = 0;
    cudaGraphNode_t node = nullptr;
    cudaGraphNodeParams np_ = {};
    cudaError_t ce = cudaGraphCreate(&graph, 0);
    if (ce == cudaSuccess) ce = cudaGraphConditionalHandleCreate(&cond, graph, 1, cudaGraphCondAssignDefault);
    if (ce == cudaSuccess) {
      np_.type = cudaGraphNodeTypeConditional;
      np_.conditional.handle = cond;
      np_.conditional.type = cudaGraphCondTypeWhile;
      np_.conditional.size = 1;
      ce = cudaGraphAddNode(&node, graph, nullptr, 0, &np_);
    }
    if (ce == cudaSuccess) {
      body = np_.conditional.phGraph_out[0];
      ce = cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    }
    if (ce == cudaSuccess) {
      const long long l0 = h->launches;
      const int rc_body = round_body();
      launch_loop_cond((unsigned long long)cond, S, max_rounds, st);
      cudaGraph_t captured = nullptr;
      ce = cudaStreamEndCapture(st, &captured);
      launches_per_round = h->launches - l0 + 1;
      h->launches = l0;
      if (rc_body != RTRG_OK) {
        cudaGraphDestroy(graph);
        return rc_body;
      }
      if (ce == cudaSuccess) ce = cudaGraphInstantiate(&gexec, graph, 0);
    }
    if (graph) cudaGraphDestroy(graph);
    if (ce == cudaSuccess) {
      whole_loop_graph = true;
    } else {
      cudaGetLastError();  // conditional nodes unavailable (old driver): per-round graph below
      gexec = nullptr;
    }
  }
  if (whole_loop_graph) {
    cudaError_t e1 = cudaGraphLaunch(gexec, st);
    if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(&n_active, S.n_active, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(&rounds, S.rounds, sizeof(long long), cudaMemcpyDeviceToHost, st);
    if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(st);
    cudaGraphExecDestroy(gexec);
    gexec = nullptr;
    if (e1 != cudaSuccess) return fail(RTRG_ECUDA, "evolution graph: %s", cudaGetErrorString(e1));
    h->launches += launches_per_round * rounds;
  } else {
    if (use_graph && n_active > 0 &&
        cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();  // e.g. the legacy default stream cannot be captured: plain launches then
    } else if (use_graph && n_active > 0) {
      cudaGraph_t graph = nullptr;
      const long long l0 = h->launches;
      const int rc_body = round_body();
      const cudaError_t ce = cudaStreamEndCapture(st, &graph);
      launches_per_round = h->launches - l0;
      h->launches = l0;
      if (rc_body != RTRG_OK) {
        if (graph) cudaGraphDestroy(graph);
        return rc_body;
      }
      if (ce != cudaSuccess) return fail(RTRG_ECUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
      const cudaError_t ie = cudaGraphInstantiate(&gexec, graph, 0);
      cudaGraphDestroy(graph);
      if (ie != cudaSuccess) return fail(RTRG_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie));
    }
    while (n_active > 0 && rounds < max_rounds) {
      if (gexec) {
        const cudaError_t ge = cudaGraphLaunch(gexec, st);
        if (ge != cudaSuccess) {
          cudaGraphExecDestroy(gexec);
          return fail(RTRG_ECUDA, "cudaGraphLaunch: %s", cudaGetErrorString(ge));
        }
        h->launches += launches_per_round;
      } else {
        const int rc_body = round_body();
        if (rc_body != RTRG_OK) return rc_body;
      }
      cudaError_t e1 = cudaMemcpyAsync(&n_active, S.n_active, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(st);
      if (e1 != cudaSuccess) {
        if (gexec) cudaGraphExecDestroy(gexec);
        return fail(RTRG_ECUDA, "round %lld: %s", rounds, cudaGetErrorString(e1));
      }
      rounds++;
    }
    if (gexec) cudaGraphExecDestroy(gexec);
  }
  CU(cudaGetLastError());

  // --- deferred output stage: the 1-loop output integrals of every stashed (cosmology, output)
  // state (rt:1646-1653) in launches of up to vch virtual cosmologies, then the tables
  ODE_LAUNCH(PC_CTRL, launch_vprep(S, st));
  const bool out_int = h->any_1loop && grp_out(h);
  for (int v0 = 0; v0 < S.NO; v0 += h->vch) {
    const int nv = std::min(h->vch, S.NO - v0);
    if (out_int) {
      Batch SV = S;
      SV.B = nv;
      SV.cosmo = S.cosmo_v + v0;
      SV.matvecs = S.matvecs_v + v0;
      h->launches += launch_integrals(tb, SV, S.ystash + (size_t)v0 * N_U * nk, (long long)N_U * nk, S.src_v, nullptr,
                                      S.vc_mask + v0, grp_out(h), 0, st, h->prof, &h->side);
    }
    ODE_LAUNCH(PC_OUTPUT, launch_output(S, h->d_kgrid, v0, nv, st));
  }
  if (sharded) XCH(h->xch->gather(S.out, h->plan_out, false, st, &xerr));  // every rank ends up with the complete tables
#undef XCH

  if (out) CU(cudaMemcpyAsync(out, S.out, h->out_total * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (hdr) CU(cudaMemcpyAsync(hdr, S.hdr, (size_t)B * MAX_OUT * 5 * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (hdr0) CU(cudaMemcpyAsync(hdr0, S.hdr0, (size_t)B * 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  h->counters.assign((size_t)4 * B, 0);
  h->matvecs.assign(B, 0);
  CU(cudaMemcpyAsync(h->counters.data(), S.counters, 4 * B * sizeof(long long), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->matvecs.data(), S.matvecs, B * sizeof(long long), cudaMemcpyDeviceToHost, st));
  std::vector<long long> mv_v(S.NO);
  CU(cudaMemcpyAsync(mv_v.data(), S.matvecs_v, S.NO * sizeof(long long), cudaMemcpyDeviceToHost, st));
  std::vector<Cosmo> cs(B);
  CU(cudaMemcpyAsync(cs.data(), S.cosmo, B * sizeof(Cosmo), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (sharded && h->xch->check(&xerr) != 0) return fail(RTRG_ECUDA, "k-shard exchange: %s", xerr.c_str());
  prof_collect(h);
  int worst = RTRG_OK;
  for (int b = 0; b < B; b++) {
    // + the evaluation at z1l (1-loop cache, rtrg_device_init) or of the first dydt_in (full)
    if (h->cos[b].c.sw_nl) h->counters[4 * b + 3] += 1;
    h->counters[4 * b + 2] += 1;  // the first dydt_in
    if (status) status[b] = cs[b].status;
    if (cs[b].status) worst = RTRG_EODE;
  }
  for (int b = 0, v = 0; b < B; b++)  // sets executed by the deferred output stage
    for (int io = 0; io < h->cos[b].c.n_out; io++, v++) h->matvecs[b] += mv_v[v];
  if (n_active > 0) return fail(RTRG_EODE, "integration did not finish within %lld rounds", max_rounds);
  if (worst) return fail(worst, "at least one cosmology failed (see status[])");
  return RTRG_OK;
}

int rtrg_fetch_outputs(rtrg_handle *h, const double **out, size_t *out_len, const double **hdr, const double **hdr0) {
  if (!h) return fail(RTRG_EINVAL, "null handle");
  if (!h->prepared) return fail(RTRG_EINVAL, "rtrg_prepare() has not been called");
  CU(cudaSetDevice(h->cfg.device));
  const size_t B = h->S.B, n_hdr = B * MAX_OUT * 5, n_hdr0 = B * 2, need = h->out_total + n_hdr + n_hdr0;
  if (need > h->h_out_cap) {
    if (h->h_out) cudaFreeHost(h->h_out);
    h->h_out = nullptr;
    h->h_out_cap = 0;
    const size_t ncap = need + need / 8;
    if (cudaHostAlloc((void **)&h->h_out, ncap * sizeof(double), cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return fail(RTRG_ENOMEM, "pinned output buffer of %zu bytes", ncap * sizeof(double));
    }
    h->h_out_cap = ncap;
  }
  cudaStream_t st = h->stream;
  double *p_out = h->h_out, *p_hdr = p_out + h->out_total, *p_hdr0 = p_hdr + n_hdr;
  CU(cudaMemcpyAsync(p_out, h->S.out, h->out_total * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(p_hdr, h->S.hdr, n_hdr * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(p_hdr0, h->S.hdr0, n_hdr0 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (out) *out = p_out;
  if (out_len) *out_len = h->out_total;
  if (hdr) *hdr = p_hdr;
  if (hdr0) *hdr0 = p_hdr0;
  return RTRG_OK;
}

int rtrg_counters(const rtrg_handle *h, int i, long long counters[4]) {
  if (!h || !counters || i < 0 || (size_t)(4 * i + 3) >= h->counters.size()) return RTRG_EINVAL;
  for (int j = 0; j < 4; j++) counters[j] = h->counters[4 * i + j];
  return RTRG_OK;
}
int rtrg_rmax_history(rtrg_handle *h, int i, double *out, int cap) {
  if (!h || !out || !h->prepared || i < 0 || i >= h->S.B || (size_t)(4 * i + 3) >= h->counters.size())
    return fail(RTRG_EINVAL, "bad argument");
  const int n = (int)std::min<long long>(std::min<long long>(h->counters[4 * i], RMAX_HIST), cap);
  if (n > 0)
    CU(cudaMemcpy(out, h->S.rmax_hist + (size_t)i * RMAX_HIST, n * sizeof(double), cudaMemcpyDeviceToHost));
  return n;
}
long long rtrg_launch_count(const rtrg_handle *h) { return h ? h->launches : 0; }
long long rtrg_matvec_sets(const rtrg_handle *h, int i) {
  if (!h || i < 0 || (size_t)i >= h->matvecs.size()) return -1;
  return h->matvecs[i];
}

int rtrg_kshard_nccl_id(char id[128]) {
  std::string err;
  if (!id) return fail(RTRG_EINVAL, "null argument");
  if (nccl_unique_id(id, &err) != 0) return fail(RTRG_ECUDA, "%s", err.c_str());
  return RTRG_OK;
}
int rtrg_kshard_init_nccl(rtrg_handle *h, const char id[128]) {
  if (!h || !id) return fail(RTRG_EINVAL, "null argument");
  if (h->cfg.k_shards < 2) return fail(RTRG_EINVAL, "handle was created with k_shards = %d", h->cfg.k_shards);
  CU(cudaSetDevice(h->cfg.device));
  std::string err;
  h->xch = make_nccl_exchange(id, h->cfg.k_shards, h->cfg.k_rank, h->cfg.device, &err);
  if (!h->xch) return fail(RTRG_ECUDA, "%s", err.c_str());
  return RTRG_OK;
}
const char *rtrg_kshard_transport(const rtrg_handle *h) { return (h && h->xch) ? h->xch->name() : "none"; }
struct rtrg_loopback {
  std::shared_ptr<LoopbackGroup> g;
};
int rtrg_kshard_loopback_create(int nranks, rtrg_loopback **out) {
  if (nranks < 2 || !out) return fail(RTRG_EINVAL, "bad argument");
  *out = new rtrg_loopback{make_loopback_group(nranks)};
  return RTRG_OK;
}
int rtrg_kshard_init_loopback(rtrg_handle *h, rtrg_loopback *g) {
  if (!h || !g) return fail(RTRG_EINVAL, "null argument");
  if (h->cfg.k_shards < 2) return fail(RTRG_EINVAL, "handle was created with k_shards = %d", h->cfg.k_shards);
  h->xch = make_loopback_exchange(g->g, h->cfg.k_rank, h->cfg.device);
  return RTRG_OK;
}
void rtrg_kshard_loopback_free(rtrg_loopback *g) { delete g; }

int rtrg_bench_integrals(rtrg_handle *h, int reps, int groups, int identical) {
  if (!h) return fail(RTRG_EINVAL, "null handle");
  if (!h->prepared) return fail(RTRG_EINVAL, "rtrg_prepare() has not been called");
  CU(cudaSetDevice(h->cfg.device));
  const int nk = h->S.nk;
  for (int r = 0; r < reps; r++)
    h->launches += launch_integrals(h->tb, h->S, h->S.y_z1l, 3LL * nk, h->S.src, nullptr, nullptr, groups, identical,
                                    h->stream, h->prof);
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaGetLastError());
  prof_collect(h);
  return RTRG_OK;
}

int rtrg_set_profiling(rtrg_handle *h, int on) {
  if (!h) return fail(RTRG_EINVAL, "null handle");
  h->prof = on ? &h->profiler : nullptr;
  for (int i = 0; i < PC_NCAT; i++) h->prof_ms[i] = 0, h->prof_n[i] = 0;
  h->profiler.reset();
  return RTRG_OK;
}
int rtrg_profile_categories(void) { return PC_NCAT; }
const char *rtrg_profile_name(int cat) {
  static const char *names[PC_NCAT] = {"k_extrap", "k_bilinear", "k_jlo", "k_pz", "k_assemble", "k_rhs",
                                       "k_combine", "k_final", "k_ctrl", "k_accept", "k_output", "k_attempt_local", "k_stage_post", "k_xch", "k_prep_inputs", "k_beta_reduce",
                                       "k_growth_ode", "k_growth_tabs", "k_qag", "k_init_state"};
  return (cat >= 0 && cat < PC_NCAT) ? names[cat] : "";
}
int rtrg_profile_query(const rtrg_handle *h, int cat, long long *n_launches, double *total_ms) {
  if (!h || cat < 0 || cat >= PC_NCAT) return RTRG_EINVAL;
  if (n_launches) *n_launches = h->prof_n[cat];
  if (total_ms) *total_ms = h->prof_ms[cat];
  return RTRG_OK;
}

// ---------------------------------------------------------------------------- hooks
static int hook_check(rtrg_handle *h, int icosmo) {
  if (!h) return fail(RTRG_EINVAL, "null handle");
  if (!h->prepared) return fail(RTRG_EINVAL, "rtrg_prepare() has not been called");
  if (icosmo < 0 || icosmo >= h->S.B) return fail(RTRG_EINVAL, "cosmology index out of range");
  cudaError_t e = cudaSetDevice(h->cfg.device);
  if (e != cudaSuccess) return fail(RTRG_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  return RTRG_OK;
}
static int hook_mask(rtrg_handle *h, int icosmo) {
  std::vector<int> m(h->S.B, 0);
  m[icosmo] = 1;
  CU(cudaMemcpyAsync(h->d_hookmask, m.data(), m.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  return RTRG_OK;
}
static int scratch(rtrg_handle *h, size_t n) {
  if (h->scratch_len >= n) return RTRG_OK;
  double *p = nullptr;
  int rc = dev_alloc(h->batch_allocs, &p, n);
  if (rc) return rc;
  if (h->d_scratch) {  // the buffer it replaces
    auto it = std::find(h->batch_allocs.begin(), h->batch_allocs.end(), (void *)h->d_scratch);
    if (it != h->batch_allocs.end()) h->batch_allocs.erase(it);
    cudaFree(h->d_scratch);
  }
  h->d_scratch = p;
  h->scratch_len = n;
  return RTRG_OK;
}

int rtrg_extrap_P(rtrg_handle *h, int icosmo, const double *lnP3nk, double *P3np) {
  int rc = hook_check(h, icosmo);
  if (rc) return rc;
  if (!lnP3nk || !P3np) return fail(RTRG_EINVAL, "null argument");
  const int nk = h->S.nk, np = h->S.np;
  rc = hook_mask(h, icosmo);
  if (rc) return rc;
  double *yslot = h->S.ytmp + (size_t)icosmo * N_U * nk;
  CU(cudaMemcpyAsync(yslot, lnP3nk, 3 * nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  launch_extrap_only(h->tb, h->S, h->S.ytmp, (long long)N_U * nk, h->d_hookmask, h->stream), h->launches++;
  CU(cudaMemcpyAsync(P3np, h->S.P3 + (size_t)icosmo * 3 * np, 3 * np * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaGetLastError());
  return RTRG_OK;
}

static int run_integrals_hook(rtrg_handle *h, int icosmo, const double *lnP3nk, bool want_raw) {
  const int nk = h->S.nk;
  int rc = hook_mask(h, icosmo);
  if (rc) return rc;
  if (want_raw && !h->d_raw) {
    rc = dev_alloc(h->batch_allocs, &h->d_raw, (size_t)h->S.B * 190 * nk);
    if (rc) return rc;
  }
  double *yslot = h->S.ytmp + (size_t)icosmo * N_U * nk;
  CU(cudaMemcpyAsync(yslot, lnP3nk, 3 * nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  h->launches += launch_integrals(h->tb, h->S, h->S.ytmp, (long long)N_U * nk, h->S.src, want_raw ? h->d_raw : nullptr,
                                  h->d_hookmask, GRP_ALL | GRP_RAW, 0, h->stream, nullptr);
  return RTRG_OK;
}

int rtrg_integrals_full(rtrg_handle *h, int icosmo, const double *lnP3nk, double *A64, double *R24, double *PTjm9,
                        double *PMRn8) {
  int rc = hook_check(h, icosmo);
  if (rc) return rc;
  if (!lnP3nk) return fail(RTRG_EINVAL, "null argument");
  const int nk = h->S.nk;
  rc = run_integrals_hook(h, icosmo, lnP3nk, false);
  if (rc) return rc;
  std::vector<double> src((size_t)N_SRC * nk);
  CU(cudaMemcpyAsync(src.data(), h->S.src + (size_t)icosmo * N_SRC * nk, src.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaGetLastError());
  if (A64) {  // 14 unique -> 64 slots with the symmetric copies (rt:147-157, 968-978)
    static const int JU[14] = {8, 9, 10, 11, 12, 13, 14, 15, 56, 57, 59, 60, 61, 63};
    static const int cp[10][2] = {{16, 8}, {18, 9}, {17, 10}, {19, 11}, {20, 12}, {22, 13}, {21, 14}, {23, 15}, {58, 57}, {62, 61}};
    std::fill(A64, A64 + (size_t)64 * nk, 0.0);
    for (int j = 0; j < 14; j++) std::copy(&src[(size_t)j * nk], &src[(size_t)(j + 1) * nk], A64 + (size_t)JU[j] * nk);
    for (auto &q : cp) std::copy(A64 + (size_t)q[1] * nk, A64 + (size_t)(q[1] + 1) * nk, A64 + (size_t)q[0] * nk);
  }
  if (R24) std::copy(&src[(size_t)14 * nk], &src[(size_t)38 * nk], R24);
  if (PTjm9) std::copy(&src[(size_t)38 * nk], &src[(size_t)47 * nk], PTjm9);
  if (PMRn8) std::copy(&src[(size_t)47 * nk], &src[(size_t)55 * nk], PMRn8);
  return RTRG_OK;
}

int rtrg_integrals_raw(rtrg_handle *h, int icosmo, const double *lnP3nk, double *J63, double *PZ63, double *Jn063,
                       double *Jlo) {
  int rc = hook_check(h, icosmo);
  if (rc) return rc;
  if (!lnP3nk) return fail(RTRG_EINVAL, "null argument");
  const int nk = h->S.nk;
  rc = run_integrals_hook(h, icosmo, lnP3nk, true);
  if (rc) return rc;
  std::vector<double> raw((size_t)190 * nk);
  CU(cudaMemcpyAsync(raw.data(), h->d_raw + (size_t)icosmo * 190 * nk, raw.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaGetLastError());
  if (J63) std::copy(&raw[0], &raw[(size_t)63 * nk], J63);
  if (PZ63) std::copy(&raw[(size_t)63 * nk], &raw[(size_t)126 * nk], PZ63);
  if (Jn063) std::copy(&raw[(size_t)126 * nk], &raw[(size_t)189 * nk], Jn063);
  if (Jlo) *Jlo = raw[(size_t)189 * nk];
  return RTRG_OK;
}

int rtrg_derivatives(rtrg_handle *h, int icosmo, double eta, const double *y, double *dy) {
  int rc = hook_check(h, icosmo);
  if (rc) return rc;
  if (!y || !dy) return fail(RTRG_EINVAL, "null argument");
  const int nk = h->S.nk;
  const size_t n = (size_t)N_U * nk;
  rc = hook_mask(h, icosmo);
  if (rc) return rc;
  cudaStream_t st = h->stream;
  double *yslot = h->S.ytmp + (size_t)icosmo * n;
  CU(cudaMemcpyAsync(yslot, y, n * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->S.t + icosmo, &eta, sizeof(double), cudaMemcpyHostToDevice, st));
  const Cosmo &c = h->cos[icosmo].c;
  if (c.sw_nl && !c.sw_1l)
    h->launches += launch_integrals(h->tb, h->S, h->S.ytmp, (long long)n, h->S.src, nullptr, h->d_hookmask, grp_rhs(h), 0, st, nullptr);
  launch_rhs(h->S, h->d_kgrid, h->S.ytmp, h->S.ynew, -1, h->d_hookmask, st), h->launches++;
  CU(cudaMemcpyAsync(dy, h->S.ynew + (size_t)icosmo * n, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  return RTRG_OK;
}

int rtrg_D_dD(rtrg_handle *h, int icosmo, double z, const double *k, int n, double *D, double *dDda) {
  int rc = hook_check(h, icosmo);
  if (rc) return rc;
  if (n <= 0 || !k || !D || !dDda) return fail(RTRG_EINVAL, "bad argument");
  rc = scratch(h, (size_t)3 * n);
  if (rc) return rc;
  cudaStream_t st = h->stream;
  double *dk = h->d_scratch, *dD1 = dk + n, *dD2 = dk + 2 * n;
  int zero = 0, err = 0;
  CU(cudaMemcpyAsync(dk, k, n * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->d_err, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
  launch_hook_DdD(h->S, icosmo, z, dk, n, dD1, dD2, h->d_err, st), h->launches++;
  CU(cudaMemcpyAsync(D, dD1, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(dDda, dD2, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(&err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  if (err) return fail(RTRG_ERANGE, "D_dD: z=%g out of bounds (the reference aborts, hdr:646-649)", z);
  return RTRG_OK;
}

int rtrg_Beta_P(rtrg_handle *h, int icosmo, double a, const double *k, int n, double *beta) {
  int rc = hook_check(h, icosmo);
  if (rc) return rc;
  if (n <= 0 || !k || !beta) return fail(RTRG_EINVAL, "bad argument");
  rc = scratch(h, (size_t)2 * n);
  if (rc) return rc;
  cudaStream_t st = h->stream;
  double *dk = h->d_scratch, *db = dk + n;
  int zero = 0, err = 0;
  CU(cudaMemcpyAsync(dk, k, n * sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->d_err, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
  launch_hook_beta(h->S, icosmo, a, dk, n, db, h->d_err, st), h->launches++;
  CU(cudaMemcpyAsync(beta, db, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(&err, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  if (err) return fail(RTRG_ERANGE, "Beta_P: a=%g > 1.001 (the reference aborts, hdr:528-531)", a);
  return RTRG_OK;
}

int rtrg_Plin(rtrg_handle *h, int icosmo, int which, double z, const double *k, int n, double *P) {
  int rc = hook_check(h, icosmo);
  if (rc) return rc;
  if (n <= 0 || !k || !P || which < 0 || which > 2) return fail(RTRG_EINVAL, "bad argument");
  rc = scratch(h, (size_t)2 * n);
  if (rc) return rc;
  cudaStream_t st = h->stream;
  double *dk = h->d_scratch, *dp = dk + n;
  CU(cudaMemcpyAsync(dk, k, n * sizeof(double), cudaMemcpyHostToDevice, st));
  launch_hook_plin(h->S, icosmo, which, z, dk, n, dp, st), h->launches++;
  CU(cudaMemcpyAsync(P, dp, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  return RTRG_OK;
}

int rtrg_initial_state(rtrg_handle *h, int icosmo, double *y, double scal[2]) {
  int rc = hook_check(h, icosmo);
  if (rc) return rc;
  const size_t n = (size_t)N_U * h->S.nk;
  if (y) CU(cudaMemcpy(y, h->d_yinit + (size_t)icosmo * n, n * sizeof(double), cudaMemcpyDeviceToHost));
  if (scal) {
    scal[0] = h->cos[icosmo].c.Norm;
    scal[1] = h->cos[icosmo].c.sigv2_0;
  }
  return RTRG_OK;
}

}  // extern "C"
