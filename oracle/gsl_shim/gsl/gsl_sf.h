/* Stand-in for <gsl/gsl_sf.h>.  TEST INFRASTRUCTURE ONLY, see gsl_rng.h. */
#ifndef SHIM_GSL_SF_H
#define SHIM_GSL_SF_H
#include <math.h>
#include "gsl_sf_psi.h"
#ifdef __cplusplus
extern "C" {
#endif
static inline double gsl_sf_lngamma(double x) { return lgamma(x); }
static inline double gsl_sf_gamma(double x) { return tgamma(x); }
#ifdef __cplusplus
}
#endif
#endif
