// mt_jump.hh against the generator itself: a stream entered by jump-ahead must continue with exactly the numbers the
// sequential generator produces there.  Built and run by tests/test_mt_jump.py.
#include "mt_jump.hh"
#include "rng.hh"

#include <cstdio>
#include <vector>

int main() {
  if (!mtjump::ready()) { printf("characteristic polynomial not recovered\n"); return 2; }
  int bad = 0;
  const unsigned long seeds[] = {0, 7, 123456789};
  const uint64_t jumps[] = {1, 623, 624, 625, 100000, 3276800, 26214400};
  for (unsigned long seed : seeds) {
    for (uint64_t skip_first : {0u, 5u, 311u}) {           // start somewhere inside a block
      Mt19937 base(seed);
      for (uint64_t i = 0; i < skip_first; ++i) base.next();
      // move a copy to the next block boundary: its array is then the history window
      Mt19937 probe = base;
      const size_t r = probe.remaining_in_block();
      for (size_t i = 0; i < r; ++i) probe.next();
      uint32_t h0[624];
      probe.history(h0);
      for (uint64_t J : jumps) {
        uint32_t h[624];
        std::copy(h0, h0 + 624, h);
        mtjump::apply(mtjump::power_of_t(J), h);
        Mt19937 jumped(1);
        jumped.set_history(h);
        Mt19937 seq = probe;
        for (uint64_t i = 0; i < J; ++i) seq.next();
        for (int i = 0; i < 2000; ++i)
          if (jumped.next() != seq.next()) { ++bad; break; }
      }
    }
  }
  // composition: two jumps equal one
  {
    Mt19937 g(42);
    for (int i = 0; i < 624; ++i) g.next();
    uint32_t a[624], b[624];
    g.history(a);
    g.history(b);
    mtjump::apply(mtjump::power_of_t(1000), a);
    mtjump::apply(mtjump::power_of_t(2345), a);
    mtjump::apply(mtjump::power_of_t(3345), b);
    for (int i = 1; i < 624; ++i) if (a[i] != b[i]) { ++bad; break; }     // (word 0 contributes its top bit only)
    if ((a[0] ^ b[0]) & 0x80000000U) ++bad;
  }
  printf("%d mismatches\n", bad);
  return bad != 0;
}
