// fixed_fmt.hh -- "%.Nf" for the model writers (gamma.txt is 2e8 numbers at n=1e6, k=200).
//
// Same bytes as printf("%.Nf") -- the reference's writers use fprintf (src/linksampling.cc:805-837,
// :1453-1476; src/fastamm2.cc:705-739): the value is scaled by 10^N and rounded in double arithmetic, which
// decides the correctly rounded result whenever the scaled value is not within a few ulps of a rounding
// boundary (x.5); the rare boundary cases, non-finite values and huge magnitudes go through snprintf.
#ifndef SVINET_B200_FIXED_FMT_HH
#define SVINET_B200_FIXED_FMT_HH

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>

// appends printf("%.<decimals>f", v) (+ `tail` if non-zero) to s; decimals in 0..9
inline void append_fixed(std::string &s, double v, int decimals, char tail = 0) {
  static const double p10[10] = {1, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9};
  static const uint64_t ip10[10] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000};
  char b[48];
  const double a = std::fabs(v);
  if (std::isfinite(v) && a < 1e15 / p10[decimals]) {
    const double scaled = a * p10[decimals];          // |error| <= scaled * 2^-53
    const double r = std::nearbyint(scaled);
    const double dist = std::fabs(std::fabs(scaled - r) - 0.5);   // distance of scaled from the nearest x.5
    if (dist > scaled * 4.5e-16 + 1e-300) {
      // the exact product rounds to r as well
      uint64_t q = (uint64_t)r;
      const uint64_t ipart = q / ip10[decimals];
      uint64_t fpart = q % ip10[decimals];
      char *e = b + sizeof b;
      char *p = e;
      if (tail) *--p = tail;
      for (int i = 0; i < decimals; ++i) { *--p = (char)('0' + fpart % 10); fpart /= 10; }
      if (decimals) *--p = '.';
      uint64_t ip = ipart;
      do { *--p = (char)('0' + ip % 10); ip /= 10; } while (ip);
      if (std::signbit(v)) *--p = '-';
      s.append(p, (size_t)(e - p));
      return;
    }
  }
  // slow path: boundary cases, non-finite values, huge magnitudes (DBL_MAX prints 309 digits before the point;
  // gamma can get there in the annealing phase when a starved community's sum[k] is ~0, src/linksampling.cc:541-542)
  char big[352];
  int len = snprintf(big, sizeof big - 1, "%.*f", decimals, v);
  if (len < 0) len = 0;
  if (len > (int)sizeof big - 2) {                 // cannot happen for a double; never index past the buffer
    std::string wide((size_t)len + 1, '\0');
    snprintf(&wide[0], wide.size(), "%.*f", decimals, v);
    wide.resize((size_t)len);
    s += wide;
    if (tail) s.push_back(tail);
    return;
  }
  if (tail) big[len++] = tail;
  s.append(big, (size_t)len);
}

#endif
