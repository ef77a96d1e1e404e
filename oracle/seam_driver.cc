// In-process seam test driver -- TEST INFRASTRUCTURE ONLY.
//
// The reference's UNMODIFIED main() (src/redTime.cc:1551-1745) and its GSL driver loop
// (gsl_odeiv_evolve_apply over the mini-GSL shim, :1588-1632) run as they are; only the GSL
// callback `derivatives` (:1416-1547, registered at :1596) is bound to the C-ABI of
// libredtime_b200.so, exactly as INTEGRATION.md section 2 tells a maintainer to do it.
//
// How the swap is done without touching the reference file: `derivatives` is made a FUNCTION-LIKE
// macro while the reference TU is included.  Its definition `int derivatives(double eta, ...)` is
// followed by '(' and becomes `reference_derivatives`; the bare name in main()'s
// `gsl_odeiv_system sys = {derivatives, dummy_jacobian, N_EQ, &mu}` is not, so it resolves to the
// function declared here.  No reference source is copied: REF_MAIN_TU is a path.
#include <cstdio>
#include <cstdlib>

#include "../include/redtime_b200.h"

static rtrg_handle *H = nullptr;
static long n_calls = 0;
int derivatives(double eta, const double y[], double dy[], void *params);

#define main redTime_reference_main
#define derivatives(...) reference_derivatives(__VA_ARGS__)
#include REF_MAIN_TU
#undef derivatives
#undef main

// the replacement body of INTEGRATION.md section 2
int derivatives(double eta, const double y[], double dy[], void *) {
  n_calls++;
  return rtrg_derivatives(H, /*icosmo=*/0, eta, y, dy) == RTRG_OK ? GSL_SUCCESS : GSL_FAILURE;
}

int main() {
  rtrg_config cfg;
  rtrg_default_config(&cfg);  // nk = 128, tolerances, z1l ... as the reference is compiled
  cfg.nk = nk;
  rtrg_run_inputs *in = nullptr;
  if (rtrg_create(&cfg, &H) != RTRG_OK || rtrg_read_run_dir(".", 0, &in) != RTRG_OK ||
      rtrg_add_cosmology(H, rtrg_inputs_cosmology(in)) != RTRG_OK || rtrg_prepare(H) != RTRG_OK) {
    std::fprintf(stderr, "seam driver: %s\n", rtrg_last_error());
    return 3;
  }
  const int rc = redTime_reference_main();
  std::fprintf(stderr, "seam driver: %ld derivatives() calls served by rtrg_derivatives\n", n_calls);
  rtrg_destroy(H);
  rtrg_free_run_inputs(in);
  return rc;
}
