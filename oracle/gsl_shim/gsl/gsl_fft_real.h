/* Mini-GSL shim -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product).
 * GSL is an un-vendored, unpinned dependency of the reference (CMakeLists.txt:9,
 * src/Makefile:5) and is absent from this image; this header declares exactly the
 * symbols src/redTime.cc and src/AU_cosmological_parameters.h use so that the
 * UNMODIFIED reference sources compile.  Algorithms restated in ../gsl_shim.cc. */
#ifndef SHIM_GSL_FFT_REAL_H
#define SHIM_GSL_FFT_REAL_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
/* in-place real -> halfcomplex, forward sign exp(-2 pi i jk/n); redTime.cc:360 */
int gsl_fft_real_radix2_transform(double data[], size_t stride, size_t n);
#ifdef __cplusplus
}
#endif
#endif
