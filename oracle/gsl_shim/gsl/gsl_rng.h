/* Minimal header-only stand-in for the parts of GSL's <gsl/gsl_rng.h> that the
 * svinet reference sources call.  TEST INFRASTRUCTURE ONLY: it exists so that the
 * unmodified reference under /root/reference/src can be compiled into
 * oracle/_ref/svinet_ref (GSL itself is not installed and there is no network).
 *
 * Generator: MT19937 (Matsumoto & Nishimura 2002 initialisation), which is GSL's
 * gsl_rng_default.  Seeding rule restated from GSL's documented behaviour: seed 0 is
 * replaced by 4357; gsl_rng_uniform = next32 / 2^32; gsl_rng_uniform_int(n) rejects
 * k = next32 / (0xffffffff / n) until k < n.  Upstream GSL is absent from
 * /root/reference, so this restatement is "parity unpinned" at the GSL boundary
 * (see DESIGN.md). */
#ifndef SHIM_GSL_RNG_H
#define SHIM_GSL_RNG_H
#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { const char *name; } gsl_rng_type;

typedef struct {
  const gsl_rng_type *type;
  uint32_t mt[624];
  int mti;
} gsl_rng;

static const gsl_rng_type shim_gsl_rng_mt19937_type = { "mt19937" };
static const gsl_rng_type *gsl_rng_mt19937 = &shim_gsl_rng_mt19937_type;
static const gsl_rng_type *gsl_rng_default = &shim_gsl_rng_mt19937_type;
static unsigned long int gsl_rng_default_seed = 0;

static inline const gsl_rng_type *gsl_rng_env_setup(void) { return gsl_rng_default; }

static inline void gsl_rng_set(gsl_rng *r, unsigned long int s) {
  if (s == 0) s = 4357;
  r->mt[0] = (uint32_t)(s & 0xffffffffUL);
  for (int i = 1; i < 624; ++i)
    r->mt[i] = 1812433253U * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
  r->mti = 624;
}

static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *T) {
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  r->type = T;
  gsl_rng_set(r, gsl_rng_default_seed);
  return r;
}

static inline void gsl_rng_free(gsl_rng *r) { free(r); }

static inline unsigned long int gsl_rng_get(const gsl_rng *cr) {
  gsl_rng *r = (gsl_rng *)cr;
  uint32_t *mt = r->mt;
  if (r->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
      uint32_t y = (mt[kk] & 0x80000000U) | (mt[kk + 1] & 0x7fffffffU);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    for (; kk < 623; kk++) {
      uint32_t y = (mt[kk] & 0x80000000U) | (mt[kk + 1] & 0x7fffffffU);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    {
      uint32_t y = (mt[623] & 0x80000000U) | (mt[0] & 0x7fffffffU);
      mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    r->mti = 0;
  }
  uint32_t k = mt[r->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680U;
  k ^= (k << 15) & 0xefc60000U;
  k ^= (k >> 18);
  return k;
}

static inline double gsl_rng_uniform(const gsl_rng *r) {
  return gsl_rng_get(r) / 4294967296.0;
}

static inline double gsl_rng_uniform_pos(const gsl_rng *r) {
  double x;
  do { x = gsl_rng_uniform(r); } while (x == 0);
  return x;
}

static inline unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n) {
  unsigned long int scale = 0xffffffffUL / n;
  unsigned long int k;
  do { k = gsl_rng_get(r) / scale; } while (k >= n);
  return k;
}

#ifdef __cplusplus
}
#endif
#endif
