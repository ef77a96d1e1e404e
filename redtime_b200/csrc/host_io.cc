// Host-side I/O that mirrors the reference executable's file/stdout contract:
//   * params_redTime.dat parser           (src/AU_cosmological_parameters.h:231-353)
//   * CAMB transfer-function readers      (hdr:556-622 for the interpolation set,
//                                          hdr:805-821 for the z=0 file)
//   * result printer, byte-format compatible with src/redTime.cc:1602-1603,1639-1641,
//     1670-1741 (setprecision(12), setw(20), general float format, two blank lines)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
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

}  // namespace

extern "C" {

int rtrg_read_run_dir(const char *dir, int camb_modern, rtrg_run_inputs **out) {
  if (!dir || !out) return RTRG_EINVAL;
  *out = nullptr;
  const std::string base = std::string(dir) + (dir[0] && dir[std::string(dir).size() - 1] != '/' ? "/" : "");
  std::ifstream in((base + "params_redTime.dat").c_str());
  if (!in.is_open()) return RTRG_EINVAL;
  rtrg_run_inputs *R = new rtrg_run_inputs();
  rtrg_cosmology &c = R->c;
  // column conventions (hdr:76-80)
  const int nVars = camb_modern ? 13 : 7, i_k = 0, i_dc = 1, i_db = 2, i_dnu = 5;
  bool ok = true;
  for (int i = 0; i < 9; i++) ok = ok && read_value(in, c.params[i]);
  for (int i = 0; i < 4; i++) ok = ok && read_value(in, c.switches[i]);
  ok = ok && read_value(in, c.z_in);
  int n_out = 0;
  ok = ok && read_value(in, n_out);
  if (!ok || n_out < 1 || n_out > RTRG_MAX_OUT) {
    delete R;
    return RTRG_EINVAL;
  }
  discard_comments(in);
  R->z_out.resize(n_out);
  for (int i = 0; i < n_out; i++) ok = ok && static_cast<bool>(in >> R->z_out[i]);
  std::string tc_file, tnu_root;
  int neut_interp_type = -100, n_interp = -100;
  ok = ok && read_value(in, tc_file);
  ok = ok && read_value(in, neut_interp_type);
  if (!ok || neut_interp_type != 0) {  // the reference aborts (hdr:293-294)
    delete R;
    return RTRG_EINVAL;
  }
  ok = ok && read_value(in, tnu_root);
  ok = ok && read_value(in, n_interp);
  if (!ok || n_interp < 0 || n_interp > RTRG_MAX_Z) {
    delete R;
    return RTRG_EINVAL;
  }
  std::vector<std::string> zstr(n_interp);
  discard_comments(in);
  for (int i = 0; i < n_interp; i++) {
    ok = ok && static_cast<bool>(in >> zstr[i]);
    R->z_interp.push_back(atof(zstr[i].c_str()));
  }
  if (!ok) {
    delete R;
    return RTRG_EINVAL;
  }

  // ---- z = 0 transfer file (hdr:805-821)
  {
    std::ifstream tf((base + tc_file).c_str());
    if (!tf.is_open()) {
      delete R;
      return RTRG_EINVAL;
    }
    std::vector<double> temp(nVars);
    bool st = true;
    discard_comments(tf);
    for (int i = 0; i < nVars; i++) st = st && static_cast<bool>(tf >> temp[i]);
    while (st) {
      discard_comments(tf);
      R->k_T.push_back(temp[i_k]);
      R->Tc_T.push_back(temp[i_dc]);
      R->Tb_T.push_back(temp[i_db]);
      for (int i = 0; i < nVars; i++) st = st && static_cast<bool>(tf >> temp[i]);
    }
  }
  // ---- interpolation set (hdr:556-622).  Massless neutrinos never open the files
  // (hdr:523-525); we honour that so such runs need only the z=0 file.
  const double fn = c.params[5] / c.params[3];
  int n_z = n_interp;
  if (fn < 1e-10) n_z = 0;
  if (n_z > 0) {
    size_t n_k = 0;
    for (int iz = 0; iz < n_z; iz++) {
      std::ifstream tf((base + tnu_root + zstr[iz] + ".dat").c_str());
      if (!tf.is_open()) {
        delete R;
        return RTRG_EINVAL;
      }
      std::vector<double> temp(nVars);
      if (iz == 0) {
        std::string line;
        while (std::getline(tf, line) && R->k_b.size() < 30000) {
          if (line.empty() || line[0] == '#' || line[0] == '\n') continue;
          std::istringstream ls(line);
          for (int j = 0; j < nVars; j++) ls >> temp[j];
          R->k_b.push_back(temp[i_k]);
          R->Tc_b.push_back(temp[i_dc]);
          R->Tnu_b.push_back(temp[i_dnu]);
        }
        n_k = R->k_b.size();
      } else {
        size_t j = 0;
        bool st = true;
        do {
          for (int q = 0; q < nVars; q++) st = st && static_cast<bool>(tf >> temp[q]);
          if (st) {
            const double x = R->k_b[j], y = temp[i_k];
            if (2.0 * std::fabs(x - y) / (std::fabs(x) + std::fabs(y)) > 1e-5) {  // hdr:605-610
              delete R;
              return RTRG_EINVAL;
            }
            R->Tc_b.push_back(temp[i_dc]);
            R->Tnu_b.push_back(temp[i_dnu]);
          }
        } while (st && ++j < n_k);
        if (R->Tc_b.size() != (size_t)(iz + 1) * n_k) {
          delete R;
          return RTRG_EINVAL;
        }
      }
    }
    c.n_kb = (int)n_k;
  } else {
    c.n_kb = 0;
    R->z_interp.clear();
  }
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
  *out = R;
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
