// Drop-in for the reference's post-processing tool convertPt (src/convert_pt.c, built by
// src/Makefile:19-20), which turns one output redshift of redTime_M<nnn>.dat into the k / P(k)
// files the emulator comparison reads:
//
//   convertPt_b200 <n_models> <step_no> <nk_pt> <models file> <redTime output folder>
//
// Same argv, same files (<folder>/STEP<step>/k_M%03d_no_interp_test.dat and pk_M%03d_...), same
// "%lf " formatting, so the outputs are byte-identical (tests/test_convert_pt.py pins this against
// the reference source compiled with a spline shim).  What the reference computes:
//   * h and f_cb = (Om - Omnu)/Om of model m from row m of the models file, after five header
//     lines (convert_pt.c:73-122; columns name Om Omb s8 h n_s w0 wa Omnu -- lower-case omegas
//     in the design files, the ratio is the same);
//   * the 17-column table with '#' lines removed (it shells out to sed, :125-131); k -> k h,
//     P_nl (column 8) -> P / h^3, P_nu (column 7) -> P / h^3 (:160-165);
//   * the redshift block of the analysis step: steps {163,...,499} -> blocks {9,...,32} of the 33
//     HACC outputs (:147-153); D (column 2) divided by its value at the LAST wavenumber of that
//     block (:173);
//   * P_nl f_cb^2 (:51-55).  A natural cubic spline of D(k) is initialised and freed without being
//     evaluated (:46-49,58-59: the files are named *_no_interp_test): gsl_spline_init only
//     requires strictly increasing k, which is checked here as well.
// Differences: no `sed`, no junk.dat in the CWD, no mkdir through system(); a table with too few
// redshift blocks is an error instead of a read of uninitialised memory.
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static bool model_row(const char *path, int model_no, double *h, double *f_cb) {
  FILE *fp = std::fopen(path, "r");
  if (!fp) {
    std::fprintf(stderr, "Couldn't read: %s\n", path);
    return false;
  }
  char tmp[300], name[64];
  for (int i = 0; i < 5; i++)
    if (!std::fgets(tmp, sizeof tmp, fp)) {
      std::fclose(fp);
      return false;
    }
  double Om = 0, Omb, s8, hh = 0, ns, w0, wa, Omnu = 0;
  for (int i = 0; i < model_no; i++)
    if (std::fscanf(fp, "%63s %lf %lf %lf %lf %lf %lf %lf %lf", name, &Om, &Omb, &s8, &hh, &ns, &w0, &wa, &Omnu) != 9) {
      std::fclose(fp);
      return false;
    }
  std::fclose(fp);
  *h = hh;
  *f_cb = (Om - Omnu) / Om;
  return true;
}

int main(int argc, char **argv) {
  if (argc != 6) {
    std::fprintf(stderr, "Invalid arguments\n");
    return -1;
  }
  const int n_models = std::atoi(argv[1]), step_no = std::atoi(argv[2]), nk_pt = std::atoi(argv[3]);
  const char *params_file = argv[4], *red_dir = argv[5];
  if (nk_pt < 3) return -1;
  const std::string out_dir = std::string(red_dir) + "/STEP" + std::to_string(step_no);
  if (::mkdir(out_dir.c_str(), 0777) != 0 && ::access(out_dir.c_str(), W_OK) != 0) {
    std::fprintf(stderr, "cannot create %s\n", out_dir.c_str());
    return 1;
  }
  static const int steps[8] = {163, 189, 247, 300, 347, 401, 453, 499};
  static const int output_z[8] = {9, 11, 14, 18, 24, 28, 31, 32};
  int z_no = 0;
  for (int i = 0; i < 8; i++)
    if (step_no == steps[i]) z_no = i;
  const int blk = output_z[z_no];

  for (int mn = 1; mn <= n_models; mn++) {
    double h, f_cb;
    if (!model_row(params_file, mn, &h, &f_cb)) {
      std::fprintf(stderr, "model %d: cannot read its row of %s\n", mn, params_file);
      return 1;
    }
    char path[512];
    std::snprintf(path, sizeof path, "%s/redTime_M%03d.dat", red_dir, mn);
    FILE *fp = std::fopen(path, "r");
    if (!fp) {
      std::fprintf(stderr, "Couldn't read: %s\n", path);
      return 1;
    }
    // rows of the 17-column table, '#' lines skipped; only block `blk` is kept
    std::vector<double> k(nk_pt), D(nk_pt), Pk(nk_pt), Pnu(nk_pt);
    std::vector<char> line(1 << 16);
    long row = 0;
    int have = 0;
    while (std::fgets(line.data(), (int)line.size(), fp)) {
      const char *p = line.data();
      while (*p == ' ' || *p == '\t') p++;
      if (*p == '#' || *p == '\n' || *p == '\r' || *p == '\0') continue;
      double c[17];
      int n = 0;
      char *end = nullptr;
      for (; n < 17; n++) {
        c[n] = std::strtod(p, &end);
        if (end == p) break;
        p = end;
      }
      if (n < 17) {
        std::fprintf(stderr, "%s: row %ld has %d columns, 17 expected\n", path, row, n);
        std::fclose(fp);
        return 1;
      }
      if (row / nk_pt == blk) {
        const int j = (int)(row % nk_pt);
        k[j] = c[0] * h;
        D[j] = c[1];
        Pnu[j] = c[6] / std::pow(h, 3.);
        Pk[j] = c[7] / std::pow(h, 3.);
        have++;
      }
      row++;
    }
    std::fclose(fp);
    if (have != nk_pt) {
      std::fprintf(stderr, "%s: redshift block %d (analysis step %d) is missing: %ld rows of %d wavenumbers\n", path,
                   blk, step_no, row, nk_pt);
      return 1;
    }
    for (int j = 1; j < nk_pt; j++)
      if (!(k[j] > k[j - 1])) {  // what gsl_spline_init(sp, k_pt, D, nk_pt) insists on
        std::fprintf(stderr, "%s: wavenumbers are not strictly increasing\n", path);
        return 1;
      }
    const double D0 = D[nk_pt - 1];
    for (int j = 0; j < nk_pt; j++) D[j] /= D0;
    char kf[512], pf[512];
    std::snprintf(kf, sizeof kf, "%s/k_M%03d_no_interp_test.dat", out_dir.c_str(), mn);
    std::snprintf(pf, sizeof pf, "%s/pk_M%03d_no_interp_test.dat", out_dir.c_str(), mn);
    FILE *fk = std::fopen(kf, "w"), *fpk = std::fopen(pf, "w");
    if (!fk || !fpk) {
      std::fprintf(stderr, "cannot write %s\n", kf);
      return 1;
    }
    for (int j = 0; j < nk_pt; j++) {
      std::fprintf(fpk, "%lf ", Pk[j] * f_cb * f_cb);
      std::fprintf(fk, "%lf ", k[j]);
    }
    std::fclose(fk);
    std::fclose(fpk);
  }
  return 0;
}
