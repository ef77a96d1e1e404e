// Host-side I/O that mirrors the reference executable's file/stdout contract:
//   * params_redTime.dat parser           (src/AU_cosmological_parameters.h:231-353)
//   * CAMB transfer-function readers      (hdr:556-622 for the interpolation set,
//                                          hdr:805-821 for the z=0 file)
//   * result printer, byte-format compatible with src/redTime.cc:1602-1603,1639-1641,
//     1670-1741 (setprecision(12), setw(20), general float format, two blank lines)
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/redtime_b200.h"

struct rtrg_run_inputs {
  rtrg_cosmology c;
  std::vector<double> z_out, k_T, Tc_T, Tb_T, z_interp, k_b, Tc_b, Tnu_b;
  std::string error;
};

namespace {

// skip '#' comment lines and empty lines (hdr:82-86)
void discard_comments(std::istream &f) {
  while (f.peek() == '#' || f.peek() == '\n') f.ignore(10000, '\n');
}

template <class T>
bool read_value(std::istream &f, T &v) {
  discard_comments(f);
  return static_cast<bool>(f >> v);
}

// ---- fast table ingestion: whole file in memory, std::from_chars, one thread
// per file.  Token semantics are those of the reference's `ifstream >> double` loops.
bool slurp(const std::string &path, std::vector<char> &buf) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? (size_t)n : 0);
  const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  buf.resize(got);
  return true;
}
inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }
// next whitespace-delimited number in [p, e); false at the end or on a malformed token
inline bool next_number(const char *&p, const char *e, double &v) {
  while (p < e && is_space(*p)) p++;
  if (p >= e) return false;
  const char *q = (*p == '+') ? p + 1 : p;
  auto r = std::from_chars(q, e, v);
  if (r.ec != std::errc()) return false;
  p = r.ptr;
  return true;
}
// the reference's discard_comments(): drop lines while the next character is '#' or '\n'
inline void skip_comment_lines(const char *&p, const char *e) {
  while (p < e && (*p == '#' || *p == '\n')) {
    while (p < e && *p != '\n') p++;
    if (p < e) p++;
  }
}

// z = 0 transfer file (hdr:805-821): comments may sit between rows
bool read_transfer_z0(const std::string &path, int nVars, int i_k, int i_dc, int i_db, std::vector<double> &k,
                      std::vector<double> &Tc, std::vector<double> &Tb) {
  std::vector<char> buf;
  if (!slurp(path, buf)) return false;
  const char *p = buf.data(), *e = p + buf.size();
  std::vector<double> row(nVars);
  k.reserve(buf.size() / (nVars * 12) + 16);
  for (;;) {
    skip_comment_lines(p, e);
    bool ok = true;
    for (int i = 0; i < nVars && ok; i++) ok = next_number(p, e, row[i]);
    if (!ok) break;
    k.push_back(row[i_k]);
    Tc.push_back(row[i_dc]);
    Tb.push_back(row[i_db]);
  }
  return true;
}
// first interpolation file (hdr:563-583): line based, '#' lines skipped, at most 30000 rows
bool read_interp_first(const std::string &path, int nVars, int i_k, int i_dc, int i_dnu, std::vector<double> &k,
                       std::vector<double> &Tc, std::vector<double> &Tnu) {
  std::vector<char> buf;
  if (!slurp(path, buf)) return false;
  const char *p = buf.data(), *e = p + buf.size();
  std::vector<double> row(nVars, 0.0);
  while (p < e && k.size() < 30000) {
    const char *eol = p;
    while (eol < e && *eol != '\n') eol++;
    if (eol > p && *p != '#') {
      const char *q = p;
      for (int i = 0; i < nVars; i++)
        if (!next_number(q, eol, row[i])) break;
      k.push_back(row[i_k]);
      Tc.push_back(row[i_dc]);
      Tnu.push_back(row[i_dnu]);
    }
    p = eol < e ? eol + 1 : e;
  }
  return true;
}
// the other interpolation files (hdr:596-622): plain token stream, n_k rows, same k list
// returns 0 ok, 1 cannot open, 2 k mismatch / short file
int read_interp_next(const std::string &path, int nVars, int i_k, int i_dc, int i_dnu, const std::vector<double> &k,
                     double *Tc, double *Tnu) {
  std::vector<char> buf;
  if (!slurp(path, buf)) return 1;
  const char *p = buf.data(), *e = p + buf.size();
  std::vector<double> row(nVars);
  for (size_t j = 0; j < k.size(); j++) {
    for (int i = 0; i < nVars; i++)
      if (!next_number(p, e, row[i])) return 2;
    const double x = k[j], y = row[i_k];
    if (2.0 * std::fabs(x - y) / (std::fabs(x) + std::fabs(y)) > 1e-5) return 2;  // hdr:605-610
    Tc[j] = row[i_dc];
    Tnu[j] = row[i_dnu];
  }
  return 0;
}

int read_run_dir_impl(const char *dir, int camb_modern, int nthreads, rtrg_run_inputs **out) {
  if (!dir || !out) return RTRG_EINVAL;
  *out = nullptr;
  const std::string base = std::string(dir) + (dir[0] && dir[std::string(dir).size() - 1] != '/' ? "/" : "");
  std::ifstream in((base + "params_redTime.dat").c_str());
  if (!in.is_open()) return RTRG_EINVAL;
  std::unique_ptr<rtrg_run_inputs> R(new rtrg_run_inputs());
  rtrg_cosmology &c = R->c;
  // column conventions (hdr:76-80)
  const int nVars = camb_modern ? 13 : 7, i_k = 0, i_dc = 1, i_db = 2, i_dnu = 5;
  bool ok = true;
  for (int i = 0; i < 9; i++) ok = ok && read_value(in, c.params[i]);
  for (int i = 0; i < 4; i++) ok = ok && read_value(in, c.switches[i]);
  ok = ok && read_value(in, c.z_in);
  int n_out = 0;
  ok = ok && read_value(in, n_out);
  if (!ok || n_out < 1 || n_out > RTRG_MAX_OUT) return RTRG_EINVAL;
  discard_comments(in);
  R->z_out.resize(n_out);
  for (int i = 0; i < n_out; i++) ok = ok && static_cast<bool>(in >> R->z_out[i]);
  std::string tc_file, tnu_root;
  int neut_interp_type = -100, n_interp = -100;
  ok = ok && read_value(in, tc_file);
  ok = ok && read_value(in, neut_interp_type);
  if (!ok || neut_interp_type != 0) return RTRG_EINVAL;  // the reference aborts (hdr:293-294)
  ok = ok && read_value(in, tnu_root);
  ok = ok && read_value(in, n_interp);
  if (!ok || n_interp < 0 || n_interp > RTRG_MAX_Z) return RTRG_EINVAL;
  std::vector<std::string> zstr(n_interp);
  discard_comments(in);
  for (int i = 0; i < n_interp; i++) {
    ok = ok && static_cast<bool>(in >> zstr[i]);
    R->z_interp.push_back(atof(zstr[i].c_str()));
  }
  if (!ok) return RTRG_EINVAL;

  // Massless neutrinos never open the interpolation files (hdr:523-525); we honour that so
  // such runs need only the z = 0 file.
  const double fn = c.params[5] / c.params[3];
  const int n_z = (fn < 1e-10) ? 0 : n_interp;
  bool ok_T = false;
  std::thread t0([&]() { ok_T = read_transfer_z0(base + tc_file, nVars, i_k, i_dc, i_db, R->k_T, R->Tc_T, R->Tb_T); });
  int n_kb = 0;
  int bad = 0;
  if (n_z > 0) {
    if (!read_interp_first(base + tnu_root + zstr[0] + ".dat", nVars, i_k, i_dc, i_dnu, R->k_b, R->Tc_b, R->Tnu_b)) {
      bad = 1;
    } else {
      n_kb = (int)R->k_b.size();
      R->Tc_b.resize((size_t)n_z * n_kb);
      R->Tnu_b.resize((size_t)n_z * n_kb);
      std::vector<int> rcs(n_z, 0);
      auto work = [&](int iz) {
        rcs[iz] = read_interp_next(base + tnu_root + zstr[iz] + ".dat", nVars, i_k, i_dc, i_dnu, R->k_b,
                                   R->Tc_b.data() + (size_t)iz * n_kb, R->Tnu_b.data() + (size_t)iz * n_kb);
      };
      if (nthreads > 1) {
        std::vector<std::thread> th;
        for (int iz = 1; iz < n_z; iz++) th.emplace_back(work, iz);
        for (auto &t : th) t.join();
      } else {
        for (int iz = 1; iz < n_z; iz++) work(iz);
      }
      for (int iz = 1; iz < n_z; iz++) bad |= rcs[iz];
    }
  }
  t0.join();
  if (!ok_T || bad) return RTRG_EINVAL;
  if (n_z == 0) R->z_interp.clear();
  c.n_kb = n_kb;
  c.n_out = n_out;
  c.z_out = R->z_out.data();
  c.n_T = (int)R->k_T.size();
  c.k_T = R->k_T.data();
  c.Tc_T = R->Tc_T.data();
  c.Tb_T = R->Tb_T.data();
  c.n_z = n_z;
  c.z_interp = R->z_interp.data();
  c.k_b = R->k_b.data();
  c.Tc_b = R->Tc_b.data();
  c.Tnu_b = R->Tnu_b.data();
  *out = R.release();
  return RTRG_OK;
}

}  // namespace

extern "C" {

int rtrg_read_run_dir(const char *dir, int camb_modern, rtrg_run_inputs **out) {
  return read_run_dir_impl(dir, camb_modern, 16, out);
}

// n run directories at once: directories are spread over the host threads
int rtrg_read_run_dirs(int n, const char *const *dirs, int camb_modern, rtrg_run_inputs **out) {
  if (n < 0 || (n > 0 && (!dirs || !out))) return RTRG_EINVAL;
  for (int i = 0; i < n; i++) out[i] = nullptr;
  int nth = (int)std::thread::hardware_concurrency();
  nth = std::max(1, std::min(std::min(nth, 32), n));
  std::vector<int> rcs(n, 0);
  std::vector<std::thread> th;
  for (int t = 0; t < nth; t++)
    th.emplace_back([&, t]() {
      for (int i = t; i < n; i += nth) rcs[i] = read_run_dir_impl(dirs[i], camb_modern, 1, &out[i]);
    });
  for (auto &t : th) t.join();
  for (int i = 0; i < n; i++)
    if (rcs[i] != RTRG_OK) {
      for (int j = 0; j < n; j++) {
        delete out[j];
        out[j] = nullptr;
      }
      return rcs[i];
    }
  return RTRG_OK;
}

const rtrg_cosmology *rtrg_inputs_cosmology(const rtrg_run_inputs *in) { return in ? &in->c : nullptr; }
void rtrg_free_run_inputs(rtrg_run_inputs *in) { delete in; }

int rtrg_print_result(void *cfile, const char *paramfile_name, int nk, int ncols, int n_out, const double *out,
                      const double *hdr, const double *hdr0) {
  FILE *f = cfile ? (FILE *)cfile : stdout;
  if (!out || !hdr || !hdr0 || nk <= 0 || ncols <= 0 || n_out <= 0) return RTRG_EINVAL;
  if (paramfile_name) std::fprintf(f, "#cosmological_parameters: opening parameter file: %s\n", paramfile_name);
  std::fprintf(f, "###main: eta_fin = %.12g, sigmaV2(z=0) = %.12g\n", hdr0[0], hdr0[1]);
  for (int io = 0; io < n_out; io++) {
    const double *h = hdr + 5 * io;
    std::fprintf(f, "### main: output at eta=%.12g, a=%.12g, z=%.12g, H=%.12g, sigma_v^2=%.12g\n", h[0], h[1], h[2],
                 h[3], h[4]);
    for (int i = 0; i < nk; i++) {
      const double *row = out + ((size_t)io * nk + i) * ncols;
      for (int cidx = 0; cidx < ncols; cidx++) std::fprintf(f, "%20.12g", row[cidx]);
      std::fputc('\n', f);
    }
    std::fputs("\n\n", f);
  }
  std::fflush(f);
  return RTRG_OK;
}

}  // extern "C"
