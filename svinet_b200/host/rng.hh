// rng.hh -- the random stream the reference consumes through GSL (src/linksampling.cc:71-75,392;
// src/linksampling.hh:336-344): MT19937 with GSL's conventions (seed 0 -> 4357, uniform = x / 2^32,
// uniform_int by rejection with scale = 0xffffffff / n).  Host-side only: on this path the generator is
// used at start-up (held-out draw + gamma initialisation), never inside the iteration.  The
// -rnode -stratified path (fastamm2.cc) also draws its minibatches from it, one Bernoulli and one or two
// uniform integers per iteration; the gamma variates follow the published Marsaglia-Tsang method with a
// polar normal (upstream GSL's exact variate stream is not reproducible here: GSL is not vendored).
#ifndef SVINET_B200_RNG_HH
#define SVINET_B200_RNG_HH
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>

class Mt19937 {
 public:
  explicit Mt19937(unsigned long seed = 0) { set(seed); }
  void set(unsigned long s) {
    if (s == 0) s = 4357;
    x_[0] = (uint32_t)(s & 0xffffffffUL);
    for (int i = 1; i < N; ++i) x_[i] = 1812433253U * (x_[i - 1] ^ (x_[i - 1] >> 30)) + (uint32_t)i;
    at_ = N;
  }
  uint32_t next() {
    if (at_ >= N) refill();
    uint32_t y = x_[at_++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680U;
    y ^= (y << 15) & 0xefc60000U;
    y ^= y >> 18;
    return y;
  }
  double uniform() { return next() / 4294967296.0; }
  // `count` consecutive uniform() values; same stream, tempered a block at a time (vectorisable)
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
  __attribute__((target_clones("avx2", "default")))   // runtime dispatch: the build host is not the run host
#endif
  void uniform_fill(double *out, size_t count) {
    size_t i = 0;
    while (i < count) {
      if (at_ >= N) refill();
      const size_t take = std::min<size_t>(count - i, (size_t)(N - at_));
      const uint32_t *src = x_ + at_;
      for (size_t j = 0; j < take; ++j) {
        uint32_t y = src[j];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680U;
        y ^= (y << 15) & 0xefc60000U;
        y ^= y >> 18;
        out[i + j] = y / 4294967296.0;
      }
      at_ += (int)take;
      i += take;
    }
  }
  // ---- hooks for mt_jump.hh (several producers generating disjoint pieces of this one stream) ----
  size_t remaining_in_block() const { return (size_t)(N - at_); }   // words left before the next refill
  // the last 624 words produced, oldest first: only meaningful when remaining_in_block() == 0
  void history(uint32_t out[624]) const { std::copy(x_, x_ + N, out); }
  // continue from a history window: the next uniform() is the word that follows it
  void set_history(const uint32_t in[624]) { std::copy(in, in + N, x_); at_ = N; }

  unsigned long uniform_int(unsigned long n) {
    const unsigned long scale = 0xffffffffUL / n;
    unsigned long k;
    do { k = next() / scale; } while (k >= n);
    return k;
  }
  // ---- variates the -rnode -stratified start-up consumes (src/fastamm2.cc:493,510,528,574) ----
  double uniform_pos() {                       // gsl_rng_uniform_pos
    double x;
    do { x = uniform(); } while (x == 0);
    return x;
  }
  unsigned bernoulli(double p) { return uniform() < p ? 1u : 0u; }   // gsl_ran_bernoulli
  double gaussian() {                          // polar (Box-Muller) method, unit variance
    double x, y, r2;
    do {
      x = -1 + 2 * uniform_pos();
      y = -1 + 2 * uniform_pos();
      r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0);
    return y * std::sqrt(-2.0 * std::log(r2) / r2);
  }
  double gamma(double a, double b) {           // gsl_ran_gamma: Marsaglia & Tsang (2000)
    if (a < 1) {
      const double u = uniform_pos();
      return gamma(1.0 + a, b) * std::pow(u, 1.0 / a);
    }
    const double d = a - 1.0 / 3.0, c = (1.0 / 3.0) / std::sqrt(d);
    double x, v, u;
    for (;;) {
      do { x = gaussian(); v = 1.0 + c * x; } while (v <= 0);
      v = v * v * v;
      u = uniform_pos();
      if (u < 1 - 0.0331 * x * x * x * x) break;
      if (std::log(u) < 0.5 * x * x + d * (1 - v + std::log(v))) break;
    }
    return b * d * v;
  }
  template <class T>
  void shuffle(T *base, size_t n) {            // gsl_ran_shuffle: Fisher-Yates from the top
    for (size_t i = n - 1; n && i > 0; i--) {
      const size_t j = uniform_int(i + 1);
      T t = base[i]; base[i] = base[j]; base[j] = t;
    }
  }

 private:
  static const int N = 624, M = 397;
  void refill() {   // three modulo-free runs of the recurrence x[i] = x[i+M] ^ twist(x[i], x[i+1])
    auto tw = [](uint32_t a, uint32_t b) {
      const uint32_t y = (a & 0x80000000U) | (b & 0x7fffffffU);
      return (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    };
    int i = 0;
    for (; i < N - M; ++i) x_[i] = x_[i + M] ^ tw(x_[i], x_[i + 1]);
    for (; i < N - 1; ++i) x_[i] = x_[i + M - N] ^ tw(x_[i], x_[i + 1]);
    x_[N - 1] = x_[M - 1] ^ tw(x_[N - 1], x_[0]);
    at_ = 0;
  }
  uint32_t x_[N];
  int at_;
};
#endif
