/* Mini-GSL shim -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product).
 * GSL is an un-vendored, unpinned dependency of the reference (CMakeLists.txt:9,
 * src/Makefile:5) and is absent from this image; this header declares exactly the
 * symbols src/redTime.cc and src/AU_cosmological_parameters.h use so that the
 * UNMODIFIED reference sources compile.  Algorithms restated in ../gsl_shim.cc. */
#ifndef SHIM_GSL_MATRIX_H
#define SHIM_GSL_MATRIX_H
/* included by redTime.cc:34 but no symbol of it is used */
#endif
