// Checks svinet_b200/host/fixed_fmt.hh against printf("%.Nf") on random values, on values at and next to the
// rounding boundaries, and on the special cases.  Built and run by tests/test_fixed_fmt.py.
#include "fixed_fmt.hh"

#include <cmath>
#include <cstdio>
#include <random>
#include <string>

int main() {
  std::mt19937_64 g(1);
  std::uniform_real_distribution<double> u(0, 1);
  long bad = 0, n = 0;
  auto check = [&](double v, int d) {
    std::string s;
    append_fixed(s, v, d, '\t');
    char b[400];
    snprintf(b, sizeof b, "%.*f\t", d, v);
    if (s != b) {
      if (bad < 10) printf("MISMATCH d=%d v=%.17g ours=%s ref=%s\n", d, v, s.c_str(), b);
      bad++;
    }
    n++;
  };
  for (int d : {3, 5, 9, 0}) {
    for (int i = 0; i < 400000; ++i) {
      const double mag = std::pow(10.0, std::floor(u(g) * 22) - 9);
      check(u(g) * mag * (u(g) < 0.2 ? -1 : 1), d);
      const double t = (std::floor(u(g) * 1e6) + 0.5) / std::pow(10.0, d);   // on / next to a rounding boundary
      check(t, d);
      check(std::nextafter(t, 0), d);
      check(std::nextafter(t, 1e300), d);
    }
    for (double v : {0.0, -0.0, 1e-300, -1e-300, 0.5, 1.5, 2.5, 0.125, 0.0625, 1e14, 9.99999e14, 1e15, 1e22, -1e15, 0.000005,
                     0.0000049999999, 1e41, -1e41, 3.3e47, 1e100, 1e300, -1e300, 1.7976931348623157e308,
                     -1.7976931348623157e308, 1.0 / 0.0, -1.0 / 0.0, std::nan("")})
      check(v, d);
  }
  printf("%ld checks, %ld mismatches\n", n, bad);
  return bad != 0;
}
