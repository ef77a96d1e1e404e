// Host-side I/O that mirrors the reference executable's file/stdout contract:
//   * params_redTime.dat parser           (src/AU_cosmological_parameters.h:231-353)
//   * CAMB transfer-function readers      (hdr:556-622 for the interpolation set,
//                                          hdr:805-821 for the z=0 file)
//   * result printer, byte-format compatible with src/redTime.cc:1602-1603,1639-1641,
//     1670-1741 (setprecision(12), setw(20), general float format, two blank lines)
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <cstring>
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

// ---- fast table ingestion: the file is mapped, only the columns the run consumes
// are converted, and the conversion takes Clinger's exact fast path -- a decimal mantissa below
// 2^53 times or divided by a power of ten up to 10^22 is ONE correctly rounded IEEE operation,
// i.e. the very double strtod / `ifstream >> double` return -- with std::from_chars for everything
// else (more than 15-19 digits, huge exponents, inf/nan).  CAMB writes 5-6 significant digits, so
// its files never leave the fast path: ~6 ms per example-1-sized directory (13 files x 15 447
// rows x 7 columns) and thread instead of 40 ms with from_chars on every token and 210 ms with
// iostreams.  Token semantics are those of the reference's `ifstream >> double` loops.
struct MappedFile {
  const char *data = nullptr;
  size_t size = 0;
  bool ok = false;
  explicit MappedFile(const std::string &path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return;
    struct stat st;
    if (::fstat(fd, &st) == 0) {
      size = (size_t)st.st_size;
      ok = true;
      if (size) {
        void *p = ::mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
        if (p == MAP_FAILED) ok = false, size = 0;
        else data = (const char *)p;
      }
    }
    ::close(fd);
  }
  ~MappedFile() {
    if (data) ::munmap((void *)data, size);
  }
  MappedFile(const MappedFile &) = delete;
  MappedFile &operator=(const MappedFile &) = delete;
};
// blanks, tabs, line ends and the other control characters (none of which occurs inside a number)
inline bool is_space(char c) { return (unsigned char)c <= (unsigned char)' '; }
const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
// number starting at p (no leading blanks); advances p past it.  false: malformed token
inline bool parse_double(const char *&p, const char *e, double &v) {
  const char *s = p;
  bool neg = false;
  if (s < e && (*s == '-' || *s == '+')) neg = (*s == '-'), s++;
  unsigned long long m = 0;
  int nd = 0, dropped = 0, frac = 0;
  bool any = false;
  for (; s < e && (unsigned)(*s - '0') < 10u; s++) {
    any = true;
    if (nd < 19) {
      m = m * 10 + (unsigned)(*s - '0');
      if (m) nd++;
    } else {
      dropped++;
    }
  }
  if (s < e && *s == '.') {
    s++;
    for (; s < e && (unsigned)(*s - '0') < 10u; s++) {
      any = true;
      if (nd < 19) {
        m = m * 10 + (unsigned)(*s - '0');
        if (m) nd++;
        frac++;
      } else {
        dropped++;  // digits beyond the 19th: not exact any more, handled by the fallback
      }
    }
  }
  if (any) {
    int ex = 0;
    bool exact = dropped == 0;
    if (s < e && (*s == 'e' || *s == 'E' || *s == 'd' || *s == 'D')) {
      if (*s == 'd' || *s == 'D') exact = false;  // not a C++ stream format: let from_chars decide
      const char *t = s + 1;
      bool eneg = false;
      if (t < e && (*t == '-' || *t == '+')) eneg = (*t == '-'), t++;
      if (t < e && (unsigned)(*t - '0') < 10u) {
        int x = 0;
        for (; t < e && (unsigned)(*t - '0') < 10u; t++)
          if (x < 100000) x = x * 10 + (*t - '0');
        ex = eneg ? -x : x;
        s = t;
      }
    }
    const int e10 = ex - frac;
    if (exact && m < (1ULL << 53) && e10 >= -22 && e10 <= 22) {
      double r = (double)m;
      r = e10 < 0 ? r / kPow10[-e10] : r * kPow10[e10];
      v = neg ? -r : r;
      p = s;
      return true;
    }
  }
  const char *q = (p < e && *p == '+') ? p + 1 : p;
  auto r = std::from_chars(q, e, v);
  if (r.ec != std::errc()) return false;
  p = r.ptr;
  return true;
}
// next whitespace-delimited number in [p, e); false at the end or on a malformed token
inline bool next_number(const char *&p, const char *e, double &v) {
  while (p < e && is_space(*p)) p++;
  if (p >= e) return false;
  return parse_double(p, e, v);
}
// step over the next token without converting it (a column the run does not consume)
inline bool skip_token(const char *&p, const char *e) {
  while (p < e && is_space(*p)) p++;
  if (p >= e) return false;
  while (p < e && !is_space(*p)) p++;
  return true;
}
// one table row of nVars tokens: columns c0 < c1 < c2 are converted, the others skipped.
// tok (optional) receives the start of the three converted tokens.
inline bool read_row3(const char *&p, const char *e, int nVars, int c0, int c1, int c2, double out[3],
                      const char **tok = nullptr) {
  for (int i = 0; i < nVars; i++) {
    if (i == c0 || i == c1 || i == c2) {
      const int j = i == c0 ? 0 : i == c1 ? 1 : 2;
      while (p < e && is_space(*p)) p++;
      if (tok) tok[j] = p;
      if (p >= e || !parse_double(p, e, out[j])) return false;
    } else if (!skip_token(p, e)) {
      return false;
    }
  }
  return true;
}
// CAMB writes fixed-width rows.  Once one row has been read token by token, its length and the
// offsets of the three wanted tokens are tried on the following rows first: three conversions per
// row and no scan over the columns nobody reads (the scan is what dominates: 20 MB of text per
// directory).  Every assumption is checked on every row (line end where expected, blank before
// and after each token); a row that does not fit falls back to the token scan.
struct RowLayout {
  int len = 0;  // bytes per row including the '\n'; 0 = not learnt (yet)
  int off[3] = {0, 0, 0};
  void learn(const char *row, const char *after_last, const char *e, const char *const tok[3]) {
    len = 0;
    const char *nl = (const char *)std::memchr(row, '\n', (size_t)(e - row));
    if (!nl || nl < after_last) return;  // the row spans lines: token stream semantics only
    for (const char *q = after_last; q < nl; q++)
      if (!is_space(*q)) return;         // more tokens on the line than the row has
    for (int j = 0; j < 3; j++) off[j] = (int)(tok[j] - row);
    len = (int)(nl + 1 - row);
  }
  bool read(const char *&p, const char *e, double out[3]) const {
    if (len == 0 || e - p < len || p[len - 1] != '\n') return false;
    const char *end = p + len;
    double v[3];
    for (int j = 0; j < 3; j++) {
      const char *q = p + off[j];
      if ((off[j] > 0 && !is_space(q[-1])) || is_space(*q)) return false;
      if (!parse_double(q, end, v[j]) || !is_space(*q)) return false;
    }
    out[0] = v[0], out[1] = v[1], out[2] = v[2];
    p = end;
    return true;
  }
};
// token-stream reader of rows (the semantics of the reference's `ifstream >> double` loops) with
// the fixed-width shortcut
struct RowReader {
  const char *b, *p, *e;  // buffer begin, cursor, end
  int nVars, c0, c1, c2;
  RowLayout layout;
  bool next(double out[3]) {
    if (layout.read(p, e, out)) return true;  // p is at a line start whenever the layout is known
    while (p < e && is_space(*p)) p++;
    const char *ls = p, *tok[3];
    while (ls > b && ls[-1] != '\n') ls--;     // start of the line the row begins on
    if (!read_row3(p, e, nVars, c0, c1, c2, out, tok)) return false;
    if (layout.len == 0) layout.learn(ls, p, e, tok);
    // leave the cursor at the start of the next line when nothing else is on this one
    const char *nl = (const char *)std::memchr(p, '\n', (size_t)(e - p));
    if (nl) {
      const char *q = p;
      while (q < nl && is_space(*q)) q++;
      if (q == nl) p = nl + 1;
    }
    return true;
  }
};
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
  MappedFile mf(path);
  if (!mf.ok) return false;
  const char *p = mf.data, *e = p + mf.size;
  const size_t guess = mf.size / (nVars * 12) + 16;
  k.reserve(guess), Tc.reserve(guess), Tb.reserve(guess);
  double row[3];
  RowReader rr{mf.data, p, e, nVars, i_k, i_dc, i_db, RowLayout()};
  for (;;) {
    skip_comment_lines(rr.p, e);
    if (!rr.next(row)) break;
    k.push_back(row[0]);
    Tc.push_back(row[1]);
    Tb.push_back(row[2]);
  }
  return true;
}
// first interpolation file (hdr:563-583): line based, '#' lines skipped, at most 30000 rows
bool read_interp_first(const std::string &path, int nVars, int i_k, int i_dc, int i_dnu, std::vector<double> &k,
                       std::vector<double> &Tc, std::vector<double> &Tnu) {
  MappedFile mf(path);
  if (!mf.ok) return false;
  const char *p = mf.data, *e = p + mf.size;
  const size_t guess = mf.size / (nVars * 12) + 16;
  k.reserve(guess), Tc.reserve(guess), Tnu.reserve(guess);
  double row[3] = {0.0, 0.0, 0.0};
  RowLayout layout;
  while (p < e && k.size() < 30000) {
    if (*p != '#' && layout.read(p, e, row)) {  // fixed-width row: p is already at the next line
      k.push_back(row[0]);
      Tc.push_back(row[1]);
      Tnu.push_back(row[2]);
      continue;
    }
    const char *eol = (const char *)std::memchr(p, '\n', (size_t)(e - p));
    if (!eol) eol = e;
    if (eol > p && *p != '#') {
      const char *q = p, *tok[3] = {p, p, p};
      // a short line keeps the previous values (sscanf-like)
      if (read_row3(q, eol, nVars, i_k, i_dc, i_dnu, row, tok) && layout.len == 0 && eol < e) layout.learn(p, q, e, tok);
      k.push_back(row[0]);
      Tc.push_back(row[1]);
      Tnu.push_back(row[2]);
    }
    p = eol < e ? eol + 1 : e;
  }
  return true;
}
// the other interpolation files (hdr:596-622): plain token stream, n_k rows, same k list
// returns 0 ok, 1 cannot open, 2 k mismatch / short file
int read_interp_next(const std::string &path, int nVars, int i_k, int i_dc, int i_dnu, const std::vector<double> &k,
                     double *Tc, double *Tnu) {
  MappedFile mf(path);
  if (!mf.ok) return 1;
  const char *p = mf.data, *e = p + mf.size;
  double row[3];
  RowReader rr{mf.data, p, e, nVars, i_k, i_dc, i_dnu, RowLayout()};
  for (size_t j = 0; j < k.size(); j++) {
    if (!rr.next(row)) return 2;
    const double x = k[j], y = row[0];
    if (2.0 * std::fabs(x - y) / (std::fabs(x) + std::fabs(y)) > 1e-5) return 2;  // hdr:605-610
    Tc[j] = row[1];
    Tnu[j] = row[2];
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

// "%20.12g" of v appended to buf: std::to_chars in general format with precision 12 is specified
// to produce what printf("%.12g") produces in the C locale (same digits, same choice between
// fixed and scientific notation, same exponent form), at a fraction of the cost; right-aligned
// in 20 columns like setw(20).  tests/test_host_io.py compares it with printf on random values.
static inline void put20(std::vector<char> &buf, double v) {
  char tmp[40];
  const auto r = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::general, 12);
  const size_t n = (size_t)(r.ptr - tmp);
  const size_t at = buf.size();
  if (n < 20) {
    buf.resize(at + 20, ' ');
    std::memcpy(buf.data() + at + 20 - n, tmp, n);
  } else {
    buf.insert(buf.end(), tmp, tmp + n);
  }
}

int rtrg_print_result(void *cfile, const char *paramfile_name, int nk, int ncols, int n_out, const double *out,
                      const double *hdr, const double *hdr0) {
  FILE *f = cfile ? (FILE *)cfile : stdout;
  if (!out || !hdr || !hdr0 || nk <= 0 || ncols <= 0 || n_out <= 0) return RTRG_EINVAL;
  if (paramfile_name) std::fprintf(f, "#cosmological_parameters: opening parameter file: %s\n", paramfile_name);
  std::fprintf(f, "###main: eta_fin = %.12g, sigmaV2(z=0) = %.12g\n", hdr0[0], hdr0[1]);
  std::vector<char> buf;
  buf.reserve((size_t)nk * ((size_t)ncols * 20 + 1) + 8);
  for (int io = 0; io < n_out; io++) {
    const double *h = hdr + 5 * io;
    std::fprintf(f, "### main: output at eta=%.12g, a=%.12g, z=%.12g, H=%.12g, sigma_v^2=%.12g\n", h[0], h[1], h[2],
                 h[3], h[4]);
    buf.clear();
    for (int i = 0; i < nk; i++) {
      const double *row = out + ((size_t)io * nk + i) * ncols;
      for (int cidx = 0; cidx < ncols; cidx++) put20(buf, row[cidx]);
      buf.push_back('\n');
    }
    buf.push_back('\n');
    buf.push_back('\n');
    std::fwrite(buf.data(), 1, buf.size(), f);
  }
  std::fflush(f);
  return RTRG_OK;
}

// "%20.12g" formatting of n values into dst (20 n + 1 bytes, NUL terminated): the table printer's
// number format, exposed for the format test
int rtrg_format_g12(const double *v, int n, char *dst) {
  if (!v || !dst || n < 0) return RTRG_EINVAL;
  std::vector<char> buf;
  for (int i = 0; i < n; i++) {
    const size_t at = buf.size();
    put20(buf, v[i]);
    if (buf.size() - at != 20) return RTRG_EINVAL;  // wider than the column: never for %.12g (<= 19 chars)
  }
  std::memcpy(dst, buf.data(), buf.size());
  dst[buf.size()] = 0;
  return RTRG_OK;
}

}  // extern "C"
