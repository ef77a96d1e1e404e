// Batch front-end, C++ only: what replaces the sequential loop of scripts/runRedTimeBatch:91-99
// (one `cd $OUTPUT_DIR; redTime > redTime_$MODEL.dat` process per model, scripts/runRedTime:196-229)
// with ONE GPU pass over all models.
//
//   redTimeBatch_b200 <manifest> [first [stride]]
//
// <manifest>: text file, one run directory per line ('#' comments), each holding params_redTime.dat
// and the CAMB files it names.  For every directory D the table is written to
// D/redTime_<basename(D)>.dat, byte-compatible with the reference's stdout.  `first`/`stride`
// select lines first, first+stride, ... so that N processes (one per GPU, RTRG_DEVICE=g) share
// one manifest without any collective.  Environment knobs as for redTime_b200.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "../../include/redtime_b200.h"

static int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

int main(int argc, char **argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s <manifest> [first [stride]]\n", argv[0]);
    return 2;
  }
  const int first = argc > 2 ? std::atoi(argv[2]) : 0, stride = argc > 3 ? std::atoi(argv[3]) : 1;
  std::ifstream mf(argv[1]);
  if (!mf.is_open() || first < 0 || stride < 1) {
    std::fprintf(stderr, "redTimeBatch_b200: cannot open %s\n", argv[1]);
    return 2;
  }
  std::string base(argv[1]);
  base = base.find('/') == std::string::npos ? std::string(".") : base.substr(0, base.rfind('/'));
  std::vector<std::string> dirs;
  int line_no = 0;
  for (std::string line; std::getline(mf, line);) {
    line = line.substr(0, line.find('#'));
    const size_t a = line.find_first_not_of(" \t\r"), b = line.find_last_not_of(" \t\r");
    if (a == std::string::npos) continue;
    line = line.substr(a, b - a + 1);
    if (line_no >= first && (line_no - first) % stride == 0) dirs.push_back(line[0] == '/' ? line : base + "/" + line);
    line_no++;
  }
  if (dirs.empty()) return 0;

  rtrg_config cfg;
  rtrg_default_config(&cfg);
  cfg.nk = env_int("RTRG_NK", cfg.nk);
  cfg.device = env_int("RTRG_DEVICE", 0);
  cfg.print_A = env_int("RTRG_PRINTA", 0);
  cfg.print_I = env_int("RTRG_PRINTI", 0);
  cfg.print_Q = env_int("RTRG_PRINTQ", 0);
  cfg.print_bias = env_int("RTRG_PRINTBIAS", 0);
  cfg.reduce_beta = 1;  // only rtrg_run is used: send what it consumes
  if (env_int("RTRG_HIACC", 0)) cfg.beta_kmin = 1e-5, cfg.beta_kmax = 20.0, cfg.n_lnk = 1000, cfg.a_early = 1e-50;
  if (env_int("RTRG_HIGH_ACCURACY", 0)) cfg.nk = 512, cfg.eps_abs = 1e-15, cfg.eps_rel = 1e-6;

  const int n = (int)dirs.size();
  std::vector<const char *> cdirs(n);
  for (int i = 0; i < n; i++) cdirs[i] = dirs[i].c_str();
  std::vector<rtrg_run_inputs *> in(n, nullptr);
  if (rtrg_read_run_dirs(n, cdirs.data(), env_int("RTRG_CAMB_MODERN", 0), in.data()) != RTRG_OK) {
    std::fprintf(stderr, "redTimeBatch_b200: cannot read one of the run directories\n");
    return 2;
  }
  rtrg_handle *h = nullptr;
  int rc = rtrg_create(&cfg, &h);
  std::vector<const rtrg_cosmology *> cos(n);
  for (int i = 0; i < n; i++) cos[i] = rtrg_inputs_cosmology(in[i]);
  if (rc == RTRG_OK) rc = rtrg_add_cosmologies(h, n, cos.data());
  if (rc == RTRG_OK) rc = rtrg_prepare(h);
  std::vector<int> status(n, 0);
  if (rc == RTRG_OK) {
    rc = rtrg_run(h, nullptr, 0, nullptr, nullptr, status.data());
    if (rc == RTRG_EODE) rc = RTRG_OK;  // per-model failures are reported below
  }
  const double *out = nullptr, *hdr = nullptr, *hdr0 = nullptr;
  size_t len = 0;
  if (rc == RTRG_OK) rc = rtrg_fetch_outputs(h, &out, &len, &hdr, &hdr0);
  if (rc != RTRG_OK) {
    std::fprintf(stderr, "redTimeBatch_b200: %s\n", rtrg_last_error());
    return 3;
  }
  int failed = 0;
  size_t off = 0;
  for (int i = 0; i < n; i++) {
    const int ncols = rtrg_num_columns(h, i), n_out = cos[i]->n_out;
    std::string d = dirs[i];
    while (d.size() > 1 && d.back() == '/') d.pop_back();
    const std::string model = d.substr(d.rfind('/') == std::string::npos ? 0 : d.rfind('/') + 1);
    const std::string path = d + "/redTime_" + model + ".dat";
    const double *hd = hdr + (size_t)i * RTRG_MAX_OUT * 5;
    if (status[i] == 101 || status[i] == 103) {
      // the reference abort()s on these (hdr:528-531, 646-649): no table, and no stale one either
      std::remove(path.c_str());
      std::fprintf(stderr, "redTimeBatch_b200: %s: status %d, no table written\n", d.c_str(), status[i]);
    } else {
      FILE *f = std::fopen(path.c_str(), "w");
      if (!f) {
        std::fprintf(stderr, "redTimeBatch_b200: cannot write %s\n", path.c_str());
        if (!status[i]) failed++;
      } else {
        int n_done = n_out;
        if (status[i]) {  // integrator failure: the outputs reached, then the warning (rt:1631-1632)
          n_done = 0;
          while (n_done < n_out && hd[(size_t)n_done * 5 + 1] != 0.0) n_done++;
        }
        if (n_done > 0) rtrg_print_result(f, "params_redTime.dat", cfg.nk, ncols, n_done, out + off, hd, hdr0 + 2 * (size_t)i);
        if (status[i]) std::fprintf(f, "#WARNING: integrator failed, status = %d\n", status[i]);
        std::fclose(f);
      }
    }
    if (status[i]) failed++;
    off += (size_t)n_out * cfg.nk * ncols;
    rtrg_free_run_inputs(in[i]);
  }
  std::printf("redTimeBatch_b200: %d models, %d failed\n", n, failed);
  rtrg_destroy(h);
  return failed ? 1 : 0;
}
