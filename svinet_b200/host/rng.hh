// rng.hh -- the random stream the reference consumes through GSL (src/linksampling.cc:71-75,392;
// src/linksampling.hh:336-344): MT19937 with GSL's conventions (seed 0 -> 4357, uniform = x / 2^32,
// uniform_int by rejection with scale = 0xffffffff / n).  Host-side only: on this path the generator is
// used at start-up (held-out draw + gamma initialisation), never inside the iteration.
#ifndef SVINET_B200_RNG_HH
#define SVINET_B200_RNG_HH
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
  unsigned long uniform_int(unsigned long n) {
    const unsigned long scale = 0xffffffffUL / n;
    unsigned long k;
    do { k = next() / scale; } while (k >= n);
    return k;
  }

 private:
  static const int N = 624, M = 397;
  void refill() {
    for (int i = 0; i < N; ++i) {
      const uint32_t y = (x_[i] & 0x80000000U) | (x_[(i + 1) % N] & 0x7fffffffU);
      x_[i] = x_[(i + M) % N] ^ (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
    }
    at_ = 0;
  }
  uint32_t x_[N];
  int at_;
};
#endif
