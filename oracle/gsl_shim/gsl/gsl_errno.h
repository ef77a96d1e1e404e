/* Mini-GSL shim -- ORACLE / TEST INFRASTRUCTURE ONLY (never linked into the product).
 * GSL is an un-vendored, unpinned dependency of the reference (CMakeLists.txt:9,
 * src/Makefile:5) and is absent from this image; this header declares exactly the
 * symbols src/redTime.cc and src/AU_cosmological_parameters.h use so that the
 * UNMODIFIED reference sources compile.  Algorithms restated in ../gsl_shim.cc. */
#ifndef SHIM_GSL_ERRNO_H
#define SHIM_GSL_ERRNO_H
enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EDOM = 1, GSL_ERANGE = 2,
       GSL_EFAULT = 3, GSL_EINVAL = 4, GSL_EFAILED = 5, GSL_ESANITY = 7, GSL_ENOMEM = 8,
       GSL_EBADFUNC = 9, GSL_EMAXITER = 11, GSL_EROUND = 18, GSL_ESING = 21, GSL_EDIVERGE = 22 };
#endif
