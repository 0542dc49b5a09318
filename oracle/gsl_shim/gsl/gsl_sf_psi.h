/* Stand-in for <gsl/gsl_sf_psi.h>: FP64 digamma for x > 0 (recurrence up to x >= 10,
 * then the asymptotic series).  TEST INFRASTRUCTURE ONLY, see gsl_rng.h. */
#ifndef SHIM_GSL_SF_PSI_H
#define SHIM_GSL_SF_PSI_H
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif
static inline double gsl_sf_psi(double x) {
  double acc = 0.0;
  if (!(x > 0.0)) {
    if (x == 0.0 || x != x) return NAN;
    /* reflection for negative non-integers (not used on the link-sampling path) */
    return gsl_sf_psi(1.0 - x) - M_PI / tan(M_PI * x);
  }
  while (x < 10.0) { acc -= 1.0 / x; x += 1.0; }
  double inv = 1.0 / x, inv2 = inv * inv;
  /* ln x - 1/(2x) - sum B_2n / (2n x^2n) */
  double series = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 - inv2 * (1.0 / 240.0
                  - inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 - inv2 * (1.0 / 12.0)))))));
  return acc + log(x) - 0.5 * inv - series;
}
#ifdef __cplusplus
}
#endif
#endif
