// nmi.hh -- normalised mutual information of two covers (overlapping communities), Lancichinetti, Fortunato &
// Kertesz, New J. Phys. 11 (2009) 033015, appendix B.
//
// The reference does not compute this itself: with -nmi it pipes communities.txt and ground_truth.txt through an
// external binary, `/usr/local/bin/mutual ... >> mutual.txt` (src/linksampling.cc:843-851), which is not part of its
// tree.  That binary is the authors' implementation of the measure above; this is the measure, written from the paper,
// so that -nmi works without the external tool.  One line "mutual3:\t<value>" per report, as in the reference's
// example output (example/n1000-k28-LFR-linksampling.tgz: mutual.txt).
#ifndef SVINET_B200_NMI_HH
#define SVINET_B200_NMI_HH

#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace nmi {

inline double h(double p) { return p > 0 ? -p * std::log2(p) : 0.0; }

// (1/|A|) sum_k H(A_k | B) / H(A_k); n11[k][l] = |A_k & B_l|
inline double conditional(uint32_t n, const std::vector<uint32_t> &sa, const std::vector<uint32_t> &sb,
                          const std::vector<std::vector<uint32_t>> &n11, bool transposed) {
  double acc = 0;
  uint32_t used = 0;
  for (size_t k = 0; k < sa.size(); ++k) {
    const double pa = (double)sa[k] / n, ha = h(pa) + h(1 - pa);
    if (!(ha > 0)) continue;
    double best = std::numeric_limits<double>::infinity();
    for (size_t l = 0; l < sb.size(); ++l) {
      const double c11 = transposed ? n11[l][k] : n11[k][l];
      const double p11 = c11 / n, p10 = pa - p11, p01 = (double)sb[l] / n - p11, p00 = 1 - p11 - p10 - p01;
      if (!(h(p11) + h(p00) > h(p01) + h(p10))) continue;          // eq. (B.14): a community is not matched to a complement
      const double pb = (double)sb[l] / n;
      const double cond = h(p11) + h(p10) + h(p01) + h(p00) - (h(pb) + h(1 - pb));
      if (cond < best) best = cond;
    }
    acc += (std::isfinite(best) ? best : ha) / ha;
    ++used;
  }
  return used ? acc / used : 0.0;
}

// covers as lists of node indices in [0, n), one list per community (empty communities are ignored)
inline double lfk(uint32_t n, const std::vector<std::vector<uint32_t>> &a, const std::vector<std::vector<uint32_t>> &b) {
  std::vector<std::vector<uint32_t>> of_b(n);
  std::vector<uint32_t> sa, sb;
  std::vector<const std::vector<uint32_t> *> aa, bb;
  for (const auto &c : a) if (!c.empty()) { aa.push_back(&c); sa.push_back((uint32_t)c.size()); }
  for (const auto &c : b) if (!c.empty()) { bb.push_back(&c); sb.push_back((uint32_t)c.size()); }
  if (aa.empty() || bb.empty() || !n) return 0.0;
  for (size_t l = 0; l < bb.size(); ++l)
    for (uint32_t v : *bb[l]) of_b[v].push_back((uint32_t)l);
  std::vector<std::vector<uint32_t>> n11(aa.size(), std::vector<uint32_t>(bb.size(), 0));
  for (size_t k = 0; k < aa.size(); ++k)
    for (uint32_t v : *aa[k])
      for (uint32_t l : of_b[v]) n11[k][l]++;
  return 1.0 - 0.5 * (conditional(n, sa, sb, n11, false) + conditional(n, sb, sa, n11, true));
}

}  // namespace nmi
#endif
