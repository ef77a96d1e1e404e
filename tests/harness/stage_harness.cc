// TEST INFRASTRUCTURE: the vectorised host staging of the a = 1 beta row (redtime_b200/csrc/host_stage.cc)
// next to the scalar rule it replaces, cub4 / lin2 of rtrg_math.h applied column by column.
#include <cstddef>

#include "../../redtime_b200/csrc/rtrg_math.h"
#include "../../redtime_b200/csrc/host_stage.cc"

extern "C" {
void sh_row_cubic(const double *tn, const double *tc, size_t n, double fn, const double *x, double xq, double *fast,
                  double *scalar) {
  rtrg::beta_row_cubic(tn, tc, n, fn, x, xq, fast);
  for (size_t i = 0; i < n; i++)
    scalar[i] = rtrg::cub4(x, fn * tn[i] / tc[i], fn * tn[n + i] / tc[n + i], fn * tn[2 * n + i] / tc[2 * n + i],
                           fn * tn[3 * n + i] / tc[3 * n + i], xq);
}
void sh_row_linear(const double *tn, const double *tc, size_t n, double fn, double x0, double x1, double xq, double *fast,
                   double *scalar) {
  rtrg::beta_row_linear(tn, tc, n, fn, x0, x1, xq, fast);
  for (size_t i = 0; i < n; i++) scalar[i] = rtrg::lin2(x0, x1, fn * tn[i] / tc[i], fn * tn[n + i] / tc[n + i], xq);
}
}
