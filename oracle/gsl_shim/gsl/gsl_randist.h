/* Stand-in for <gsl/gsl_randist.h>.  TEST INFRASTRUCTURE ONLY, see gsl_rng.h.
 * Only gsl_ran_bernoulli_pdf is reached on the -link-sampling path; the samplers
 * below are valid draws from the named distributions (Marsaglia-Tsang gamma with a
 * polar-method normal) but do NOT replay upstream GSL's exact variate streams. */
#ifndef SHIM_GSL_RANDIST_H
#define SHIM_GSL_RANDIST_H
#include <math.h>
#include <string.h>
#include "gsl_rng.h"
#ifdef __cplusplus
extern "C" {
#endif

static inline double gsl_ran_bernoulli_pdf(const unsigned int k, double p) {
  if (k == 0) return 1 - p;
  if (k == 1) return p;
  return 0;
}

static inline unsigned int gsl_ran_bernoulli(const gsl_rng *r, double p) {
  return gsl_rng_uniform(r) < p ? 1u : 0u;
}

static inline double shim_ran_gaussian(const gsl_rng *r) {
  double x, y, r2;
  do {
    x = -1 + 2 * gsl_rng_uniform_pos(r);
    y = -1 + 2 * gsl_rng_uniform_pos(r);
    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0);
  return y * sqrt(-2.0 * log(r2) / r2);
}

static inline double gsl_ran_gamma(const gsl_rng *r, const double a, const double b) {
  if (a < 1) {
    double u = gsl_rng_uniform_pos(r);
    return gsl_ran_gamma(r, 1.0 + a, b) * pow(u, 1.0 / a);
  }
  double d = a - 1.0 / 3.0, c = (1.0 / 3.0) / sqrt(d), x, v, u;
  for (;;) {
    do { x = shim_ran_gaussian(r); v = 1.0 + c * x; } while (v <= 0);
    v = v * v * v;
    u = gsl_rng_uniform_pos(r);
    if (u < 1 - 0.0331 * x * x * x * x) break;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
  }
  return b * d * v;
}

static inline double gsl_ran_beta(const gsl_rng *r, const double a, const double b) {
  double x1 = gsl_ran_gamma(r, a, 1.0), x2 = gsl_ran_gamma(r, b, 1.0);
  return x1 / (x1 + x2);
}

static inline void gsl_ran_dirichlet(const gsl_rng *r, const size_t K, const double alpha[], double theta[]) {
  double norm = 0.0;
  for (size_t i = 0; i < K; i++) { theta[i] = gsl_ran_gamma(r, alpha[i], 1.0); norm += theta[i]; }
  for (size_t i = 0; i < K; i++) theta[i] /= norm;
}

static inline unsigned int shim_ran_binomial(const gsl_rng *r, double p, unsigned int n) {
  unsigned int k = 0;
  for (unsigned int i = 0; i < n; i++) if (gsl_rng_uniform(r) < p) k++;
  return k;
}

static inline void gsl_ran_multinomial(const gsl_rng *r, const size_t K, const unsigned int N,
                                       const double p[], unsigned int n[]) {
  double norm = 0.0, sum_p = 0.0;
  unsigned int sum_n = 0;
  for (size_t k = 0; k < K; k++) norm += p[k];
  for (size_t k = 0; k < K; k++) {
    n[k] = p[k] > 0.0 ? shim_ran_binomial(r, p[k] / (norm - sum_p), N - sum_n) : 0;
    sum_p += p[k];
    sum_n += n[k];
  }
}

static inline void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t n, size_t size) {
  char *b = (char *)base;
  char tmp[256];
  for (size_t i = n - 1; i > 0; i--) {
    size_t j = gsl_rng_uniform_int(r, i + 1);
    if (i == j || size > sizeof(tmp)) continue;
    memcpy(tmp, b + i * size, size);
    memcpy(b + i * size, b + j * size, size);
    memcpy(b + j * size, tmp, size);
  }
}
#ifdef __cplusplus
}
#endif
#endif
