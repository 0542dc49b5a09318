// mt_jump.hh -- jump-ahead for MT19937: the state J outputs further on, without generating them.
//
// Why: LinkSampling::init_gamma2 (src/linksampling.cc:374-401) consumes K uniforms per link from ONE mt19937 stream --
// 2e10 numbers at n = 1e6, K = 200, 1e8 links, ~20 s of a single core however fast the consumers are.  The stream is
// a linear recurrence over GF(2), so the state after J steps is g(F) applied to the state, F the one-word transition
// and g(t) = t^J mod phi(t), phi the characteristic polynomial (degree 19937) -- Haramoto, Matsumoto, Nishimura,
// Panneton, L'Ecuyer, "Efficient jump ahead for F2-linear random number generators", INFORMS J. Comput. 20 (2008).
// Several producer threads then generate disjoint pieces of the SAME stream: identical numbers, in parallel.
//
//   * phi(t) is not tabulated: it is recovered once with Berlekamp-Massey from 2 x 19937 output bits.
//   * t^J mod phi by square-and-multiply on bit-packed polynomials (~0.1 s per exponent).
//   * g(F) s by Horner's rule on the HISTORY window (the last 624 words produced; the next word is
//     x[n] = x[n-227] ^ twist(x[n-624], x[n-623])): ~0.5 ms per jump.
#ifndef SVINET_B200_MT_JUMP_HH
#define SVINET_B200_MT_JUMP_HH

#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

namespace mtjump {

constexpr int kDeg = 19937;
constexpr int kWords = (kDeg + 64) / 64;   // 312 x 64 bits = 19968 >= 19938 coefficients
constexpr int kN = 624, kM = 397;

using Poly = std::vector<uint64_t>;         // bit i of word i/64 = coefficient of t^i

inline uint32_t twist(uint32_t a, uint32_t b) {
  const uint32_t y = (a & 0x80000000U) | (b & 0x7fffffffU);
  return (y >> 1) ^ ((y & 1U) ? 0x9908b0dfU : 0U);
}

// history window in a circular buffer: logical word j (0 = oldest of the last 624) lives at w[(head + j) % 624]
struct Window {
  uint32_t w[kN];
  int head = 0;
  void step() {   // produce the next word; it becomes the newest, the oldest drops out
    const uint32_t nw = w[(head + kM) % kN] ^ twist(w[head], w[(head + 1) % kN]);
    w[head] = nw;
    head = (head + 1) % kN;
  }
};

// Berlekamp-Massey over GF(2): connection polynomial C (C[0] = 1) of the bit sequence s, and its length L
inline void berlekamp_massey(const std::vector<uint8_t> &s, Poly &c_out, int &l_out) {
  const int nw = (int)(s.size() / 64) + 2;
  Poly c(nw, 0), b(nw, 0), t(nw, 0), rev(nw, 0);   // rev: bit i = s[N - i]
  c[0] = b[0] = 1;
  int L = 0, m = 1;
  for (int N = 0; N < (int)s.size(); ++N) {
    // rev <<= 1; rev bit 0 = s[N]
    uint64_t carry = s[N];
    for (int i = 0; i < nw; ++i) {
      const uint64_t nc = rev[i] >> 63;
      rev[i] = (rev[i] << 1) | carry;
      carry = nc;
    }
    uint64_t acc = 0;
    const int lw = L / 64 + 1;
    for (int i = 0; i < lw && i < nw; ++i) acc ^= c[i] & rev[i];
    const int d = __builtin_parityll(acc);
    if (!d) { ++m; continue; }
    const bool grow = 2 * L <= N;
    if (grow) t = c;
    // c ^= b << m
    const int ws = m / 64, bs = m % 64;
    for (int i = nw - 1; i >= ws; --i) {
      uint64_t v = b[i - ws] << bs;
      if (bs && i - ws - 1 >= 0) v |= b[i - ws - 1] >> (64 - bs);
      c[i] ^= v;
    }
    if (grow) { L = N + 1 - L; b = t; m = 1; } else { ++m; }
  }
  c_out = c;
  l_out = L;
}

// phi(t), degree 19937, from the generator's own output (any seed; the recurrence is the same)
inline const Poly &charpoly() {
  static Poly phi;
  static std::once_flag once;
  std::call_once(once, [] {
    Window h;
    uint32_t x = 19650218U;
    for (int i = 0; i < kN; ++i) { h.w[i] = x; x = 1812433253U * (x ^ (x >> 30)) + (uint32_t)(i + 1); }
    std::vector<uint8_t> bits(2 * kDeg + 64);
    for (size_t i = 0; i < bits.size(); ++i) {
      h.step();
      bits[i] = (uint8_t)(h.w[(h.head + kN - 1) % kN] & 1U);   // a linear functional of the state: bit 0 of the new word
    }
    Poly c;
    int L = 0;
    berlekamp_massey(bits, c, L);
    phi.assign(kWords, 0);
    if (L != kDeg) return;   // (cannot happen: the sequence has linear complexity 19937) -- callers check degree()
    for (int j = 0; j <= kDeg; ++j)   // phi_j = C_{L-j}
      if ((c[(L - j) / 64] >> ((L - j) % 64)) & 1ULL) phi[j / 64] |= 1ULL << (j % 64);
  });
  return phi;
}

inline bool ready() {
  const Poly &phi = charpoly();
  return (phi[kDeg / 64] >> (kDeg % 64)) & 1ULL;
}

// r (degree < 2*19937, 2*kWords words) reduced mod phi, in place; result in the low kWords words
inline void reduce(std::vector<uint64_t> &r) {
  const Poly &phi = charpoly();
  for (int i = 2 * kDeg - 1; i >= kDeg; --i) {
    if (!((r[i / 64] >> (i % 64)) & 1ULL)) continue;
    const int sh = i - kDeg, ws = sh / 64, bs = sh % 64;
    for (int j = 0; j < kWords; ++j) {
      r[j + ws] ^= phi[j] << bs;
      if (bs) r[j + ws + 1] ^= phi[j] >> (64 - bs);
    }
  }
}

// t^e mod phi
inline Poly power_of_t(uint64_t e) {
  std::vector<uint64_t> r(2 * kWords + 2, 0), sq(2 * kWords + 2, 0);
  r[0] = 1;
  for (int bit = 63; bit >= 0; --bit) {
    // square: coefficient i -> 2i
    std::fill(sq.begin(), sq.end(), 0);
    for (int i = 0; i < kWords; ++i) {
      uint64_t v = r[i];
      while (v) {
        const int b = __builtin_ctzll(v);
        v &= v - 1;
        const int p = 2 * (i * 64 + b);
        sq[p / 64] |= 1ULL << (p % 64);
      }
    }
    reduce(sq);
    std::copy(sq.begin(), sq.begin() + kWords, r.begin());
    std::fill(r.begin() + kWords, r.end(), 0);
    if ((e >> bit) & 1ULL) {   // times t
      uint64_t carry = 0;
      for (int i = 0; i < kWords + 1; ++i) {
        const uint64_t nc = r[i] >> 63;
        r[i] = (r[i] << 1) | carry;
        carry = nc;
      }
      reduce(r);
      std::fill(r.begin() + kWords, r.end(), 0);
    }
  }
  return Poly(r.begin(), r.begin() + kWords);
}

// hist (the last 624 words produced, oldest first) -> the last 624 words produced J steps later, g = t^J mod phi
inline void apply(const Poly &g, uint32_t hist[kN]) {
  Window h;
  std::memset(h.w, 0, sizeof h.w);
  for (int i = kDeg - 1; i >= 0; --i) {
    h.step();   // (F applied to the zero window is the zero window: the leading steps cost nothing but time)
    if ((g[i / 64] >> (i % 64)) & 1ULL) {
      const int first = kN - h.head;   // logical j = 0 .. first-1 are w[head ..], the rest wrap to w[0 ..]
      for (int j = 0; j < first; ++j) h.w[h.head + j] ^= hist[j];
      for (int j = first; j < kN; ++j) h.w[j - first] ^= hist[j];
    }
  }
  for (int j = 0; j < kN; ++j) hist[j] = h.w[(h.head + j) % kN];
}

}  // namespace mtjump
#endif
