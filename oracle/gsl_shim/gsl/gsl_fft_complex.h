/* Mini-GSL shim -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product).
 * GSL is an un-vendored, unpinned dependency of the reference (CMakeLists.txt:9,
 * src/Makefile:5) and is absent from this image; this header declares exactly the
 * symbols src/redTime.cc and src/AU_cosmological_parameters.h use so that the
 * UNMODIFIED reference sources compile.  Algorithms restated in ../gsl_shim.cc. */
#ifndef SHIM_GSL_FFT_COMPLEX_H
#define SHIM_GSL_FFT_COMPLEX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef double *gsl_complex_packed_array;
/* packed (re,im) arrays; redTime.cc:386-392 */
int gsl_fft_complex_radix2_forward(gsl_complex_packed_array data, size_t stride, size_t n);
int gsl_fft_complex_radix2_backward(gsl_complex_packed_array data, size_t stride, size_t n);
int gsl_fft_complex_radix2_inverse(gsl_complex_packed_array data, size_t stride, size_t n);
#ifdef __cplusplus
}
#endif
#endif
