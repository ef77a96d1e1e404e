// Host-side staging arithmetic of rtrg_add_cosmologies (reduce_beta = 1): the row beta(a = 1, k_b) of
// the neutrino table beta = f_nu T_nu / T_c (hdr:556-623), interpolated in a column by column with the
// rule of tab:437-471.  One call handles the n_kb (15 447 for the shipped CAMB tables) columns of one
// cosmology.  Written so that the compiler vectorises it: the interpolation weights depend on the a
// nodes only and are formed once (in the operation order of cub4 / lin2, rtrg_math.h), the four
// divisions per column run as packed divisions.  Same bits as the scalar expression, 8x faster
// (54 us instead of 430 us per cosmology on the build host) -- with 8 ranks sharing one host's cores
// the staging of a 1024-cosmology batch was what the end-to-end pipeline waited for.
#include <cstddef>

namespace rtrg {

// row1[i] = cub4(x, beta(0, i), ..., beta(3, i), xq) with beta(m, i) = fn * tn[m * n + i] / tc[m * n + i]
__attribute__((target_clones("avx2", "default"), optimize("O3")))
void beta_row_cubic(const double *tn, const double *tc, size_t n, double fn, const double *x, double xq, double *row1) {
  const double w0 = (xq - x[1]) * (xq - x[2]) * (xq - x[3]) / (x[0] - x[1]) / (x[0] - x[2]) / (x[0] - x[3]);
  const double w1 = (xq - x[0]) * (xq - x[2]) * (xq - x[3]) / (x[1] - x[0]) / (x[1] - x[2]) / (x[1] - x[3]);
  const double w2 = (xq - x[0]) * (xq - x[1]) * (xq - x[3]) / (x[2] - x[0]) / (x[2] - x[1]) / (x[2] - x[3]);
  const double w3 = (xq - x[0]) * (xq - x[1]) * (xq - x[2]) / (x[3] - x[0]) / (x[3] - x[1]) / (x[3] - x[2]);
  const double *n0 = tn, *n1 = tn + n, *n2 = tn + 2 * n, *n3 = tn + 3 * n;
  const double *c0 = tc, *c1 = tc + n, *c2 = tc + 2 * n, *c3 = tc + 3 * n;
  for (size_t i = 0; i < n; i++)
    row1[i] = w0 * (fn * n0[i] / c0[i]) + w1 * (fn * n1[i] / c1[i]) + w2 * (fn * n2[i] / c2[i]) + w3 * (fn * n3[i] / c3[i]);
}

// row1[i] = lin2(x0, x1, beta(0, i), beta(1, i), xq)
__attribute__((target_clones("avx2", "default"), optimize("O3")))
void beta_row_linear(const double *tn, const double *tc, size_t n, double fn, double x0, double x1, double xq, double *row1) {
  const double *n0 = tn, *n1 = tn + n, *c0 = tc, *c1 = tc + n;
  for (size_t i = 0; i < n; i++) {
    const double f0 = fn * n0[i] / c0[i], f1 = fn * n1[i] / c1[i];
    row1[i] = f0 + (f1 - f0) / (x1 - x0) * (xq - x0);
  }
}

}  // namespace rtrg
