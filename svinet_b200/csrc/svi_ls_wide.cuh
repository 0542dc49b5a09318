// svi_ls_wide.cuh -- the link-sampling kernels for K > 1024 (the reference's type limit is 65 535, src/env.hh:37).
//
// The register-tiled kernels (svi_ls_kernels.cuh, svi_ls_ring.cuh) keep a whole K-row in the registers of one lane
// group and stop at 32 lanes x 16 double2.  Here one BLOCK of kWideT threads owns a work item (a segment, a node
// row, a held-out pair); thread t holds the columns t, t + kWideT, ... and walks them in a loop, so K is bounded by
// memory only.  Same formulation as the register tiles (pull form over the CSR of half-edges, one arg-max per link,
// deferred annealing rescale), always in the log domain: `b` holds Elogpi, `eb` holds
// Elogbeta[:,0], and a neighbour costs three passes over its row (max, sum of exps, accumulate).  These kernels are the
// correctness path of a range the benchmarked configurations never reach; they are not tuned beyond coalescing.
//
// Accumulators that do not fit registers live in the kernel's own output rows (part[seg], the block's kpart slot):
// every column of such a row is read and written by ONE thread of ONE block, so there are no atomics and the
// summation order is fixed.  Block-wide reductions go through a shared-memory tree in a fixed order.
//
// The kernels use nothing but threadIdx/blockIdx, __syncthreads and shared arrays, so the same source also runs as
// host code under tests/cc/wide_emul.cc (a block = kWideT host threads, __syncthreads = a barrier): that is how their
// arithmetic and their synchronisation are checked against the oracle (and under ThreadSanitizer) without a GPU.
#pragma once
#include "svi_ls_kernels.cuh"
#include "svi_wide_reduce.cuh"

namespace svi {

// ------------------------------------------------------------------------------------------
// phi sweep (src/linksampling.cc:605-725), one block per segment; the launch covers two ranges of the segment table
// like k_phi.  part[seg] = sum of phi over the segment's neighbours.
//   full branch :685-701 (dense) / :634-664 (SPARSE: phi over the union of the endpoints' active communities)
//   shortcut    :619-631 (exactly one endpoint converged: one-hot phi)
//   tally       :704-717 (COMM): the first maximum of phi; `publish` sets the neighbour's bit too (one arg-max per link)
template <bool SPARSE, bool COMM>
__global__ void __launch_bounds__(SVI_WIDE_T) k_phi_wide(const Params P, const uint32_t seg_first, const uint32_t seg_end,
                                                         const uint32_t seg_first2, const uint32_t seg_end2,
                                                         const uint32_t publish) {
  SVI_BLOCK_SHARED double red[kWideT];
  SVI_BLOCK_SHARED uint32_t redk[kWideT];
  const uint32_t t = threadIdx.x, sidx = blockIdx.x, nseg1 = seg_end - seg_first;
  if (sidx >= nseg1 + (seg_end2 - seg_first2)) return;   // (the whole block)
  const uint32_t seg = sidx < nseg1 ? seg_first + sidx : seg_first2 + (sidx - nseg1);
  const uint32_t p = P.seg_node[seg], beg = P.seg_beg[seg], cnt = P.seg_cnt[seg];
  const uint32_t pc = P.conv[p];
  const uint32_t pa = SPARSE ? P.active[p] : 0u;
  const double *ep = P.b + (size_t)p * P.ld;
  const uint32_t *ap = P.abits + (size_t)p * P.words;
  double *out = P.part + (size_t)seg * P.ld;
  for (uint32_t c = t; c < P.ld; c += kWideT) out[c] = 0.0;   // (thread t owns the columns t, t + kWideT, ...)

  for (uint32_t j = 0; j < cnt; ++j) {
    const uint32_t q = P.col[beg + j], qc = P.conv[q];
    if ((pc != 0u) != (qc != 0u)) {   // one-hot phi on the converged endpoint's community
      const uint32_t c = (pc ? pc : qc) - 1u;
      if (c % kWideT == t) out[c] += 1.0;
      continue;
    }
    const double *eq = P.b + (size_t)q * P.ld;
    const uint32_t *aq = P.abits + (size_t)q * P.words;
    const bool sparse = SPARSE && pa < P.k_div10 && P.active[q] < P.k_div10;
    // pass 1: the largest exponent (over the active union when the branch applies)
    double m = -CUDART_INF;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      if (SPARSE && sparse && !(((ap[c >> 5] | aq[c >> 5]) >> (c & 31u)) & 1u)) continue;
      m = fmax(m, (ep[c] + eq[c]) + P.eb[c]);
    }
    m = wide_max(m, red);
    if (!(m > -CUDART_INF)) continue;   // empty active union: phi stays all-zero (the same for every thread)
    // pass 2: the normaliser
    double s = 0.0;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      if (SPARSE && sparse && !(((ap[c >> 5] | aq[c >> 5]) >> (c & 31u)) & 1u)) continue;
      s += exp(((ep[c] + eq[c]) + P.eb[c]) - m);
    }
    s = wide_sum(s, red);
    if (!(s > 0.0)) continue;
    const double inv = 1.0 / s;
    // pass 3: accumulate phi, and its first maximum when the tally is on
    double best = 0.0;
    uint32_t bestk = 0xffffffffu;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      if (SPARSE && sparse && !(((ap[c >> 5] | aq[c >> 5]) >> (c & 31u)) & 1u)) continue;
      const double ph = exp(((ep[c] + eq[c]) + P.eb[c]) - m) * inv;
      out[c] += ph;
      if (COMM && ph > best) { best = ph; bestk = c; }
    }
    if (COMM) {
      wide_argmax(best, bestk, red, redk);
      if (t == 0 && best > 0.0) {
        atomicOr(P.mbits + (size_t)p * P.words + (bestk >> 5), 1u << (bestk & 31u));
        if (publish) atomicOr(P.mbits + (size_t)q * P.words + (bestk >> 5), 1u << (bestk & 31u));
      }
    }
  }
}

// compute_mean_indicators (src/linksampling.cc:526-545) for the nodes [node_begin, node_end), fed by the partial rows
// of the sweep; persistent blocks, a block's column sums (sum, s1, s2) accumulate in its kpart slot [3][cap].
static __global__ void __launch_bounds__(SVI_WIDE_T) k_node_wide(const Params P, const uint32_t cap) {
  const uint32_t t = threadIdx.x;
  double *out = P.kpart + (size_t)blockIdx.x * 3 * cap;
  for (uint32_t c = t; c < P.ld; c += kWideT) out[c] = out[cap + c] = out[2 * (size_t)cap + c] = 0.0;
  for (uint32_t p = P.node_begin + blockIdx.x; p < P.node_end; p += gridDim.x) {
    const double tlp = P.tl[p], rest = (double)P.n - tlp - 1.0;
    const uint32_t i = p - P.shard_begin;
    const uint32_t lo0 = P.node_seg_lo[i], lo1 = P.node_seg_lo[i + 1], up0 = P.node_seg_up[i], up1 = P.node_seg_up[i + 1];
    double *grow = P.gacc + (size_t)p * P.ld, *mrow = P.mphi + (size_t)p * P.ld;
    for (uint32_t c = t; c < P.ld; c += kWideT) {
      if (tlp == 0.0) {   // :532-533 -- the row stays at alpha, mphi keeps its previous value
        grow[c] = c < P.k ? P.alpha : 0.0;
        continue;
      }
      double acc = 0.0;   // the node's "lo" segments, then its "up" segments: a fixed order
      for (uint32_t s = lo0; s < lo1; ++s) acc += P.part[(size_t)s * P.ld + c];
      for (uint32_t s = up0; s < up1; ++s) acc += P.part[(size_t)s * P.ld + c];
      double g = P.alpha + acc;
      const double m = c < P.k ? (g - P.alpha) / tlp : 0.0;
      out[c] += acc;
      out[cap + c] += m;
      out[2 * (size_t)cap + c] += m * m;
      g = c < P.k ? g + rest * m : 0.0;
      mrow[c] = m;
      grow[c] = g;
    }
  }
}

// s3 sweep (src/linksampling.cc:731-746) over the half-edges the shard's nodes own; persistent blocks, the block's
// column sums accumulate in its kpart slot [cap].  Shortcut: column pc, not pc-1; column K reads as 0.
static __global__ void __launch_bounds__(SVI_WIDE_T) k_s3_wide(const Params P, const uint32_t cap) {
  const uint32_t t = threadIdx.x;
  double *out = P.kpart + (size_t)blockIdx.x * cap;
  for (uint32_t c = t; c < P.ld; c += kWideT) out[c] = 0.0;
  for (uint32_t seg = P.nseg_lo + blockIdx.x; seg < P.nseg; seg += gridDim.x) {
    const uint32_t p = P.seg_node[seg], beg = P.seg_beg[seg], cnt = P.seg_cnt[seg];
    const uint32_t pc = P.conv[p];
    const double *mp = P.mphi + (size_t)p * P.ld;
    for (uint32_t j = 0; j < cnt; ++j) {
      const uint32_t q = P.col[beg + j], qc = P.conv[q];
      const double *mq = P.mphi + (size_t)q * P.ld;
      if ((pc != 0u) == (qc != 0u)) {
        for (uint32_t c = t; c < P.k; c += kWideT) out[c] += mp[c] * mq[c];
      } else if (pc) {
        if ((pc - 1u) % kWideT == t) out[pc - 1u] += pc < P.k ? mq[pc] : 0.0;
      } else {
        if ((qc - 1u) % kWideT == t) out[qc - 1u] += qc < P.k ? mp[qc] : 0.0;
      }
    }
  }
}

// gamma <- gammanext with the deferred annealing rescale (:541-542), set_dir_exp(gamma) (src/linksampling.hh:171-187)
// -> b = Elogpi (log domain), prune / check_and_set_converged (src/linksampling.cc:456-491).  One block per node row.
//   FROM_GACC = false: initial refresh from an uploaded gamma (no rescale, no prune; :123,:561).
template <bool FROM_GACC>
__global__ void __launch_bounds__(SVI_WIDE_T) k_refresh_wide(const Params P) {
  SVI_BLOCK_SHARED double red[kWideT];
  SVI_BLOCK_SHARED uint32_t redc[kWideT];
  SVI_BLOCK_SHARED uint32_t redm[kWideT];
  const uint32_t t = threadIdx.x, p = P.node_begin + blockIdx.x;
  if (p >= P.node_end) return;   // (the whole block)
  double *grow = P.gamma + (size_t)p * P.ld, *brow = P.b + (size_t)p * P.ld;
  const double *arow = P.gacc + (size_t)p * P.ld;
  const bool rescale = FROM_GACC && P.tl[p] != 0.0;
  double rs = 0.0;
  for (uint32_t c = t; c < P.ld; c += kWideT) {
    double g;
    if (FROM_GACC) {
      g = arow[c];
      if (rescale) g *= P.scale[c];
      if (c >= P.k) g = 0.0;
      grow[c] = g;
    } else {
      g = grow[c];
    }
    if (c < P.k) rs += g;
  }
  rs = wide_sum(rs, red);
  const double psi_sum = digamma_pos(rs);
  uint32_t cnt = 0, lastk = 0;
  for (uint32_t c = t; c < P.ld; c += kWideT) {
    const double g = grow[c];   // (this thread's own store above)
    brow[c] = c < P.k ? digamma_pos(g) - psi_sum : 0.0;
    if (FROM_GACC && c < P.k && g - P.alpha >= 1.0) {   // prune: communities with gamma - alpha >= 1 (:462-468)
      cnt++;
      lastk = c;
    }
  }
  if (FROM_GACC) {
    wide_count_max(cnt, lastk, redc, redm);
    if (t == 0) {   // sticky: never cleared (:472-473)
      const uint32_t was = P.conv[p];
      if (cnt == 1u && was == 0u) *P.conv_dirty = 1u;
      P.conv_next[p] = cnt == 1u ? lastk + 1u : was;
      P.active[p] = cnt;
    }
    // the active mask, word by word from the row the block has just written (wide_count_max ended in a barrier)
    for (uint32_t w = t; w < P.words; w += kWideT) {
      uint32_t word = 0;
      for (uint32_t i = 0; i < 32u; ++i) {
        const uint32_t c = 32u * w + i;
        if (c < P.k && grow[c] - P.alpha >= 1.0) word |= 1u << i;
      }
      P.abits[(size_t)p * P.words + w] = word;
    }
  }
}

// held-out log-likelihood, LinkSampling::edge_likelihood (src/linksampling.hh:259-292) in the O(K) non-link form of
// k_heldout; one block per pair, a grid-stride loop over the pairs.
template <class ROWS>
__global__ void __launch_bounds__(SVI_WIDE_T) k_heldout_wide(const Params P, const ROWS rows, uint64_t npairs,
                                                             const uint32_t *pp, const uint32_t *qq, const uint8_t *yy,
                                                             double epsilon, double *out, unsigned long long *bad) {
  SVI_BLOCK_SHARED double red[kWideT];
  const uint32_t t = threadIdx.x;
  for (uint64_t i = blockIdx.x; i < npairs; i += gridDim.x) {
    const uint32_t p = pp[i], q = qq[i];
    if (p >= P.n || q >= P.n) {   // *bad = 1 + index of one bad pair
      if (t == 0) {
        atomicMax(bad, (unsigned long long)i + 1ull);
        out[i] = CUDART_NAN;
      }
      continue;
    }
    const int y = yy[i];
    const double *rp = rows.gamma_row(p, P.ld), *rq = rows.gamma_row(q, P.ld);
    double sp = 0.0, sq = 0.0;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      sp += rp[c];
      sq += rq[c];
    }
    sp = wide_sum(sp, red);
    sq = wide_sum(sq, red);
    double piq_sum = 0.0;
    for (uint32_t c = t; c < P.k; c += kWideT) piq_sum += rq[c] / sq;
    piq_sum = wide_sum(piq_sum, red);
    double s = 0.0;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      const double l0 = P.lambda[2 * c], l1 = P.lambda[2 * c + 1], rate = l0 / (l0 + l1);
      const double px = rp[c] / sp, gq = rq[c] / sq;
      if (y) s += px * gq * rate;
      else s += px * (gq * (1.0 - rate) + (piq_sum - gq) * (1.0 - epsilon));   // diagonal (1 - beta_z) + off-diagonal
    }
    s = wide_sum(s, red);
    if (s < 1e-30) s = 1e-30;
    if (t == 0) out[i] = log(s);
  }
}

}  // namespace svi
