/* Mini-GSL shim -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product).
 * GSL is an un-vendored, unpinned dependency of the reference (CMakeLists.txt:9,
 * src/Makefile:5) and is absent from this image; this header declares exactly the
 * symbols src/redTime.cc and src/AU_cosmological_parameters.h use so that the
 * UNMODIFIED reference sources compile.  Algorithms restated in ../gsl_shim.cc. */
#ifndef SHIM_GSL_FFT_HALFCOMPLEX_H
#define SHIM_GSL_FFT_HALFCOMPLEX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
/* halfcomplex -> real, sign exp(+...); inverse scales by 1/n; redTime.cc:363-370 */
int gsl_fft_halfcomplex_radix2_backward(double data[], size_t stride, size_t n);
int gsl_fft_halfcomplex_radix2_inverse(double data[], size_t stride, size_t n);
#ifdef __cplusplus
}
#endif
#endif
