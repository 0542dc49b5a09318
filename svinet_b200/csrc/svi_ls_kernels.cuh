// svi_ls_kernels.cuh -- sm_100a kernels of the link-sampling iteration.
//
// Formulation.  The reference sweeps the undirected link list once and
// scatters phi into both endpoint rows (push form, src/linksampling.cc:605-725).  Here every
// node PULLS from its neighbours over a CSR of half-edges, so a gamma row is owned by exactly
// one group of lanes: no atomics, deterministic summation order, half the HBM bytes of the
// push form.  The per-edge softmax
//     phi[k] = exp(Elogpi[p][k] + Elogpi[q][k] + Elogbeta[k][0] - logsumexp)      (:685-694)
// is evaluated in factorised form: the refresh kernel stores  b[i][k] = exp(Elogpi[i][k] -
// max_k Elogpi[i][:])  once per node per iteration (N*K exps instead of E*K), so an edge costs
// two multiplies, one group reduction and one FMA per community:
//     w[k] = (b[p][k] * b[q][k]) * eb[k];   phi[k] = w[k] / sum_k w[k]
// (k_phi multiplies b[p]*b[q] first: IEEE multiplication commutes, so both directions of a link see bit-identical
// phi; the ring sweeps fold eb into the self factor, (b[p]*eb)*b[q], and their two directions may differ in the last
// bit -- which is why the link-community arg-max is taken ONCE per link, on the owner's side, for both endpoints.)
// For K > 256 the factors can underflow, and the same kernels run in the log domain (LOGDOM: rows hold Elogpi, one
// exp per element).
//
// Thread mapping: a GROUP of G lanes (G = 2..32, a power of two) owns one work item (a segment
// of <= seg_len neighbours of one node, or one node row); each lane holds V double2 = 2V
// columns, column c = 2*(lane + G*j) + {0,1}, so one row load is V coalesced 16-byte loads per
// lane (LDG.E.128).  Rows are padded to `ld` (multiple of 4 doubles, 32-byte sectors).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace svi {

struct Params {
  // problem
  uint32_t n, k, ld, words;        // words = (k+31)/32
  uint32_t node_begin, node_end;   // node range a launch works on (the shard, or one chunk of it)
  uint32_t shard_begin;            // first node of the handle's block: base of node_seg_lo / node_seg_up
  double alpha, eta0, eta1, ones_d;
  uint32_t k_div10;
  // graph: half-edges of the shard's nodes.  A node's list is [neighbours it does not own | neighbours it OWNS]
  // (s3_owner, svi_ls.cu); each part is cut into work segments of <= seg_len neighbours.  Segment table: all "lo"
  // segments (node order) first, [0, nseg_lo), then all "up" (owned) segments, [nseg_lo, nseg).  The phi sweep runs
  // over all of them, the s3 sweep and the link-community tally over the "up" ones (one visit per LINK).
  const uint32_t *col;                                 // reordered inside a segment by k_partition
  const uint32_t *seg_node, *seg_beg, *seg_cnt;
  uint32_t *seg_nnc;                                   // [nseg] leading not-converged neighbours of each segment
  uint32_t nseg, nseg_lo;
  const uint32_t *node_seg_lo, *node_seg_up;           // [nlocal+1] each: segment ranges of every local node
  uint32_t *conv_dirty;                                // [1] some node newly converged since the last k_partition
  const double *tl;                                    // [n]
  // state
  double *b;        // [n*ld] exp(Elogpi - rowmax)   (LOGDOM: Elogpi)
  double *mphi;     // [n*ld]
  double *gamma;    // [n*ld]
  double *gacc;     // [n*ld] unscaled gammanext of the sweep
  double *part;     // [nseg*ld] per-segment partial rows
  double *kvec;     // [4*ld] sum, s1, s2, s3
  double *kpart;    // [kpart_blocks * 3 * cap] block partials of column sums
  double *lambda;   // [k*2]
  double *eb;       // [ld] exp(Elogbeta[:,0] - max)  (LOGDOM: Elogbeta[:,0])
  double *scale;    // [ld] annealing rescale ones/sum[k] (1 when not annealing)
  uint32_t *conv;   // [n] converged (0 or c+1) as this iteration's sweeps see it
  uint32_t *conv_next;  // [n] ... as prune leaves it for the next iteration (double-buffered: the s3 sweep of this
                        //     iteration may run after the refresh and must still see the pre-prune flags, :731-761)
  uint32_t *active; // [n] active_comms
  uint32_t *abits;  // [n*words] active-community mask
  uint32_t *mbits;  // [n*words] link-community membership
};

// ------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (G == 32) {
    return 0xffffffffu;
  } else {
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << (lane & ~(unsigned)(G - 1));
  }
}

// xor butterfly: every lane ends with the bit-identical total (a+b == b+a in IEEE)
template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ double group_max(double v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o));
  return v;
}

__device__ __forceinline__ double2 ld_row2(const double *row, uint32_t c, uint32_t ld) {
  // 16-byte read-only load; columns beyond the padded row read as zero
  if (c < ld) return __ldg(reinterpret_cast<const double2 *>(row + c));
  return make_double2(0.0, 0.0);
}
__device__ __forceinline__ void st_row2(double *row, uint32_t c, uint32_t ld, double2 v) {
  if (c < ld) *reinterpret_cast<double2 *>(row + c) = v;
}

// gsl_sf_psi for x > 0 (call sites src/linksampling.hh:181,184): recurrence to x >= 10, then the
// asymptotic series ln x - 1/(2x) - sum B_2n/(2n x^2n) (truncation error < 1e-17 at x = 10).
// (The reciprocals stay separate divisions on purpose: they are independent and pipeline, whereas carrying the
// sum as one fraction num/den -- two FMAs per term, one division -- is a serial chain and measured SLOWER in the
// latency-bound refresh kernel: 4.1 ms vs 3.4 ms at config 4.)
__device__ __forceinline__ double digamma_pos(double x) {
  double acc = 0.0;
  if (!(x > 0.0)) return CUDART_NAN;
  // recurrence psi(x) = psi(x+1) - 1/x up to x >= 10.  Late in a run most of gamma sits near alpha << 1, i.e. ten
  // steps per element, and an FP64 division is ~12 instructions: four steps are taken at once as ONE division,
  //   1/a + 1/b + 1/c + 1/d = ((a+b)cd + (c+d)ab) / (ab cd)      (a..d = x..x+3; products < 1e4, no range issue)
  // -- 3 divisions instead of 10 for x < 1, the same end point x + n as the step-by-step loop (the refresh kernel at
  // config 4, once the state has concentrated: 4.6 -> 4.1 ms with the four-step form alone).
  while (x < 7.0) {
    const double a = x, b = x + 1.0, c = x + 2.0, d = x + 3.0;
    const double ab = a * b, cd = c * d;
    acc -= fma(a + b, cd, (c + d) * ab) / (ab * cd);
    x += 4.0;
  }
  if (x < 9.0) {   // two steps as one division
    acc -= (x + (x + 1.0)) / (x * (x + 1.0));
    x += 2.0;
  }
  if (x < 10.0) { acc -= 1.0 / x; x += 1.0; }
  const double inv = 1.0 / x, inv2 = inv * inv;
  const double series = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 - inv2 * (1.0 / 240.0
                        - inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 - inv2 * (1.0 / 12.0)))))));
  return acc + log(x) - 0.5 * inv - series;
}

// ------------------------------------------------------------------------------------------
// K1: phi sweep.  One group per segment of one node's neighbour list.
//   full branch      src/linksampling.cc:685-701 (dense) / :634-664 (active-set, SPARSE)
//   shortcut branch  :619-631 (exactly one endpoint converged -> one-hot phi)
//   tally            :668-681,704-717 (COMM): arg-max community of the link
// Output: part[seg] = sum of phi over the segment's neighbours (a K-row).
template <int G, int V, bool LOGDOM, bool SPARSE, bool COMM>
__global__ void __launch_bounds__(256) k_phi(const Params P, const uint32_t seg_first, const uint32_t seg_end,
                                             const uint32_t seg_first2, const uint32_t seg_end2, const uint32_t publish) {
  constexpr int U = V == 1 ? 4 : (V <= 4) ? 2 : 1;   // neighbour rows in flight per group (small rows: latency, not bytes)
  const unsigned mask = group_mask<G>();
  const uint32_t lane = threadIdx.x & (G - 1);
  // two ranges of the segment table in one launch (a chunk of nodes: its "lo" segments and its "up" segments)
  const uint32_t sidx = (blockIdx.x * blockDim.x + threadIdx.x) / G, nseg1 = seg_end - seg_first;
  if (sidx >= nseg1 + (seg_end2 - seg_first2)) return;
  const uint32_t seg = sidx < nseg1 ? seg_first + sidx : seg_first2 + (sidx - nseg1);
  const uint32_t p = P.seg_node[seg], beg = P.seg_beg[seg], cnt = P.seg_cnt[seg];
  const uint32_t pc = P.conv[p];
  const uint32_t pa = SPARSE ? P.active[p] : 0u;
  const double *brow_p = P.b + (size_t)p * P.ld;

  double2 bs[V], eb[V], acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c = 2u * (lane + G * j);
    bs[j] = ld_row2(brow_p, c, P.ld);
    eb[j] = ld_row2(P.eb, c, P.ld);
    acc[j] = make_double2(0.0, 0.0);
  }
  uint32_t mb = 0;  // membership word `lane` of node p

  for (uint32_t j0 = 0; j0 < cnt; j0 += U) {
    uint32_t q[U], qc[U];
    bool live[U], full[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      live[u] = j0 + u < cnt;
      q[u] = live[u] ? __ldg(P.col + beg + j0 + u) : p;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      qc[u] = live[u] ? P.conv[q[u]] : 0u;
      full[u] = live[u] && !((pc != 0u) != (qc[u] != 0u));
    }
    double2 row[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double *brow_q = P.b + (size_t)q[u] * P.ld;
#pragma unroll
      for (int j = 0; j < V; ++j)
        row[u][j] = full[u] ? ld_row2(brow_q, 2u * (lane + G * j), P.ld) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!live[u]) continue;
      if (!full[u]) {
        // one-hot phi on the converged endpoint's community
        const uint32_t c = (pc ? pc : qc[u]) - 1u;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const uint32_t c0 = 2u * (lane + G * j);
          if (c == c0) acc[j].x += 1.0;
          if (c == c0 + 1u) acc[j].y += 1.0;
        }
        continue;
      }
      bool sparse = false;
      if (SPARSE) sparse = pa < P.k_div10 && P.active[q[u]] < P.k_div10;
      double2 w[V];
      if (!LOGDOM) {
#pragma unroll
        for (int j = 0; j < V; ++j) {
          w[j].x = (bs[j].x * row[u][j].x) * eb[j].x;
          w[j].y = (bs[j].y * row[u][j].y) * eb[j].y;
        }
      } else {
        double m = -CUDART_INF;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const uint32_t c0 = 2u * (lane + G * j);
          w[j].x = (bs[j].x + row[u][j].x) + eb[j].x;
          w[j].y = (bs[j].y + row[u][j].y) + eb[j].y;
          if (c0 < P.k) m = fmax(m, w[j].x);
          if (c0 + 1u < P.k) m = fmax(m, w[j].y);
        }
        m = group_max<G>(m, mask);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const uint32_t c0 = 2u * (lane + G * j);
          w[j].x = c0 < P.k ? exp(w[j].x - m) : 0.0;
          w[j].y = c0 + 1u < P.k ? exp(w[j].y - m) : 0.0;
        }
      }
      if (SPARSE && sparse) {
        // restrict phi to the union of the endpoints' active communities (:634-664)
        const uint32_t *ap = P.abits + (size_t)p * P.words, *aq = P.abits + (size_t)q[u] * P.words;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const uint32_t c0 = 2u * (lane + G * j);
          uint32_t bits = 0;
          if (c0 < P.k) bits = (ap[c0 >> 5] | aq[c0 >> 5]) >> (c0 & 31u);
          if (!(bits & 1u)) w[j].x = 0.0;
          if (!(bits & 2u)) w[j].y = 0.0;
        }
      }
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < V; ++j) s += w[j].x + w[j].y;
      s = group_sum<G>(s, mask);
      if (!(s > 0.0)) continue;  // empty active union: phi stays all-zero, nothing accumulates
      const double inv = 1.0 / s;
      double best = 0.0;
      uint32_t bestk = 0xffffffffu;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const double px = w[j].x * inv, py = w[j].y * inv;
        acc[j].x += px;
        acc[j].y += py;
        if (COMM) {  // D1Array::max (src/matrix.hh:521-532): first strictly larger value wins
          const uint32_t c0 = 2u * (lane + G * j);
          if (px > best) { best = px; bestk = c0; }
          if (py > best) { best = py; bestk = c0 + 1u; }
        }
      }
      if (COMM) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(mask, best, o);
          const uint32_t ok = __shfl_xor_sync(mask, bestk, o);
          if (ob > best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
        }
        if (best > 0.0 && lane == (bestk >> 5)) {
          mb |= 1u << (bestk & 31u);
          if (publish) atomicOr(P.mbits + (size_t)q[u] * P.words + lane, 1u << (bestk & 31u));   // one arg-max per link
        }
      }
    }
  }
  double *out = P.part + (size_t)seg * P.ld;
#pragma unroll
  for (int j = 0; j < V; ++j) st_row2(out, 2u * (lane + G * j), P.ld, acc[j]);
  if (COMM && mb) atomicOr(P.mbits + (size_t)p * P.words + lane, mb);
}

// ------------------------------------------------------------------------------------------
// block-level, fixed-order reduction of per-group column accumulators into kpart
template <int G, int V>
__device__ __forceinline__ void block_reduce_columns(const double2 (&v)[V], double *smem, double *out) {
  constexpr int CAP = 2 * G * V;
  const uint32_t lane = threadIdx.x & (G - 1), grp = threadIdx.x / G, ngrp = blockDim.x / G;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c = 2u * (lane + G * j);
    smem[grp * CAP + c] = v[j].x;
    smem[grp * CAP + c + 1] = v[j].y;
  }
  __syncthreads();
  for (uint32_t c = threadIdx.x; c < (uint32_t)CAP; c += blockDim.x) {
    double s = 0.0;
    for (uint32_t g = 0; g < ngrp; ++g) s += smem[g * CAP + c];
    out[c] = s;
  }
}

// K2: compute_mean_indicators (src/linksampling.cc:526-545) for the shard's rows, fed by the
// partial rows of K1.  gacc keeps gammanext WITHOUT the annealing rescale (it needs the global
// sum[k]); the rescale is applied by k_refresh.  Column sums (sum, s1, s2) leave as block partials.
template <int G, int V>
__global__ void __launch_bounds__(256) k_node(const Params P) {
  extern __shared__ double smem[];
  constexpr int CAP = 2 * G * V;
  const uint32_t lane = threadIdx.x & (G - 1);
  const uint32_t ggid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const uint32_t ngroups = gridDim.x * blockDim.x / G;
  double2 csum[V], cs1[V], cs2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) csum[j] = cs1[j] = cs2[j] = make_double2(0.0, 0.0);

  for (uint32_t p = P.node_begin + ggid; p < P.node_end; p += ngroups) {
    double2 acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {   // the node's "lo" segments, then its "up" segments: a fixed order
      const uint32_t *off = half ? P.node_seg_up : P.node_seg_lo;
      const uint32_t s0 = off[p - P.shard_begin], s1 = off[p - P.shard_begin + 1];
      for (uint32_t s = s0; s < s1; ++s) {
        const double *row = P.part + (size_t)s * P.ld;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const double2 r = ld_row2(row, 2u * (lane + G * j), P.ld);
          acc[j].x += r.x;
          acc[j].y += r.y;
        }
      }
    }
    const double tlp = P.tl[p];
    double *grow = P.gacc + (size_t)p * P.ld, *mrow = P.mphi + (size_t)p * P.ld;
    if (tlp == 0.0) {  // :532-533 -- row stays at alpha, mphi keeps its previous value
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const uint32_t c0 = 2u * (lane + G * j);
        st_row2(grow, c0, P.ld, make_double2(c0 < P.k ? P.alpha : 0.0, c0 + 1u < P.k ? P.alpha : 0.0));
      }
      continue;
    }
    const double rest = (double)P.n - tlp - 1.0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t c0 = 2u * (lane + G * j);
      double2 g, m;
      g.x = P.alpha + acc[j].x;
      g.y = P.alpha + acc[j].y;
      m.x = c0 < P.k ? (g.x - P.alpha) / tlp : 0.0;
      m.y = c0 + 1u < P.k ? (g.y - P.alpha) / tlp : 0.0;
      csum[j].x += acc[j].x;
      csum[j].y += acc[j].y;
      cs1[j].x += m.x;
      cs1[j].y += m.y;
      cs2[j].x += m.x * m.x;
      cs2[j].y += m.y * m.y;
      g.x = c0 < P.k ? g.x + rest * m.x : 0.0;
      g.y = c0 + 1u < P.k ? g.y + rest * m.y : 0.0;
      st_row2(mrow, c0, P.ld, m);
      st_row2(grow, c0, P.ld, g);
    }
  }
  double *out = P.kpart + (size_t)blockIdx.x * 3 * CAP;
  block_reduce_columns<G, V>(csum, smem, out);
  block_reduce_columns<G, V>(cs1, smem, out + CAP);
  block_reduce_columns<G, V>(cs2, smem, out + 2 * CAP);
}

// fixed-order final reduction of `nvec` column vectors over `nblocks` block partials: one warp per (vector, column);
// lane l sums the blocks l, l+32, ... in order, then a fixed butterfly -- the same bits on every run, and a serial
// chain of nblocks/32 instead of nblocks (the block count is ~1000 at small sizes, where this kernel used to cost as
// much as the sweep it follows)
static __global__ void __launch_bounds__(256) k_reduce_kpart(const double *kpart, uint32_t nblocks, uint32_t nvec,
                                                             uint32_t cap, double *kvec, uint32_t ld) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (w >= nvec * ld) return;
  const uint32_t v = w / ld, c = w % ld;
  double s = 0.0;
  if (c < cap)
    for (uint32_t b = lane; b < nblocks; b += 32u) s += kpart[((size_t)b * nvec + v) * cap + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) kvec[(size_t)v * ld + c] = s;
}
static inline uint32_t reduce_kpart_blocks(uint32_t nvec, uint32_t ld) { return (nvec * ld * 32u + 255u) / 256u; }

// K3: s3 sweep (src/linksampling.cc:731-746) over the half-edges the shard's nodes own (p = owner, q = other):
//   s3[k] += mphi[p][k]*mphi[q][k]              both or neither endpoint converged
//   s3[pc-1] += mphi[q][pc]  /  s3[qc-1] += mphi[p][qc]   exactly one converged (sic: column
//   pc, not pc-1; column K reads the never-written slack after the row == 0)
template <int G, int V>
__global__ void __launch_bounds__(256) k_s3(const Params P) {
  extern __shared__ double smem[];
  constexpr int CAP = 2 * G * V;
  constexpr int U = V == 1 ? 4 : (V <= 4) ? 2 : 1;
  const unsigned mask = group_mask<G>();
  const uint32_t lane = threadIdx.x & (G - 1);
  const uint32_t ggid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const uint32_t ngroups = gridDim.x * blockDim.x / G;
  double2 s3[V];
#pragma unroll
  for (int j = 0; j < V; ++j) s3[j] = make_double2(0.0, 0.0);

  for (uint32_t seg = P.nseg_lo + ggid; seg < P.nseg; seg += ngroups) {
    const uint32_t p = P.seg_node[seg], beg = P.seg_beg[seg], cnt = P.seg_cnt[seg];
    const uint32_t pc = P.conv[p];
    const double *mrow_p = P.mphi + (size_t)p * P.ld;
    double2 t[V];
#pragma unroll
    for (int j = 0; j < V; ++j) t[j] = make_double2(0.0, 0.0);
    double one_hot = 0.0;       // shortcut mass for column pc-1 (p converged)
    // ids and converged flags of G neighbours at a time, one per lane (coalesced: two dependent loads per G
    // neighbours instead of two per U), handed out by shuffles.  (The same change in k_phi measured slower.)
    for (uint32_t c0 = 0; c0 < cnt; c0 += G) {
     const uint32_t rem = min((uint32_t)G, cnt - c0);
     const uint32_t qv = lane < rem ? __ldg(P.col + beg + c0 + lane) : p;
     const uint32_t qcv = lane < rem ? P.conv[qv] : 0u;
     for (uint32_t j1 = 0; j1 < rem; j1 += U) {
      uint32_t q[U], qc[U];
      bool live[U], full[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t idx = j1 + u;
        live[u] = idx < rem;
        q[u] = __shfl_sync(mask, qv, min(idx, (uint32_t)G - 1u), G);
        qc[u] = __shfl_sync(mask, qcv, min(idx, (uint32_t)G - 1u), G);
        if (!live[u]) { q[u] = p; qc[u] = 0u; }
        full[u] = live[u] && !((pc != 0u) != (qc[u] != 0u));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;
        const double *mrow_q = P.mphi + (size_t)q[u] * P.ld;
        if (full[u]) {
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const double2 r = ld_row2(mrow_q, 2u * (lane + G * j), P.ld);
            t[j].x += r.x;
            t[j].y += r.y;
          }
        } else if (pc) {
          one_hot += pc < P.k ? mrow_q[pc] : 0.0;
        } else {
          const uint32_t c = qc[u] - 1u;  // s3[qc-1] += mphi[p][qc]
          const double v = qc[u] < P.k ? mrow_p[qc[u]] : 0.0;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const uint32_t c0 = 2u * (lane + G * j);
            if (c == c0) s3[j].x += v;
            if (c == c0 + 1u) s3[j].y += v;
          }
        }
      }
     }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t c0 = 2u * (lane + G * j);
      const double2 mp = ld_row2(mrow_p, c0, P.ld);
      s3[j].x += mp.x * t[j].x;
      s3[j].y += mp.y * t[j].y;
      if (pc && pc - 1u == c0) s3[j].x += one_hot;
      if (pc && pc - 1u == c0 + 1u) s3[j].y += one_hot;
    }
  }
  block_reduce_columns<G, V>(s3, smem, P.kpart + (size_t)blockIdx.x * CAP);
}

// K4: lambda finish (src/linksampling.cc:748-755) + set_dir_exp(lambda) (:758-759) + the factors
// the next sweep reads.  One block.  lambda[k][0] = eta0 + sum[k]  (lnext[:,0] and _sum receive
// identical increments, :625-626,699-700);  lambda[k][1] = eta1 + s1^2 - s2 - s3.
template <bool LOGDOM>
__global__ void k_lambda(const Params P, int annealing, int update_lambda) {
  __shared__ double red[32];
  const double *sum = P.kvec, *s1 = P.kvec + P.ld, *s2 = P.kvec + 2 * P.ld, *s3 = P.kvec + 3 * P.ld;
  double mymax = -CUDART_INF;
  for (uint32_t c = threadIdx.x; c < P.k; c += blockDim.x) {
    double l0, l1;
    if (update_lambda) {
      l0 = P.eta0 + sum[c];
      l1 = P.eta1 + (s1[c] * s1[c] - s2[c] - s3[c]);
      P.lambda[2 * c] = l0;
      P.lambda[2 * c + 1] = l1;
      P.scale[c] = annealing ? P.ones_d / sum[c] : 1.0;
    } else {
      l0 = P.lambda[2 * c];
      l1 = P.lambda[2 * c + 1];
    }
    const double e0 = digamma_pos(l0) - digamma_pos(l0 + l1);
    P.eb[c] = e0;
    mymax = fmax(mymax, e0);
  }
  if constexpr (!LOGDOM) {
    // block max, then eb = exp(Elogbeta0 - max)
    for (int o = 16; o > 0; o >>= 1) mymax = fmax(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mymax;
    __syncthreads();
    double m = red[0];
    for (uint32_t w = 1; w < (blockDim.x + 31) / 32; ++w) m = fmax(m, red[w]);
    for (uint32_t c = threadIdx.x; c < P.k; c += blockDim.x) P.eb[c] = exp(P.eb[c] - m);
  }
}

// annealing rescale ones/sum[k] (:541-542) on its own, for drivers that run the refresh before the s3 sweep
static __global__ void k_scale(const Params P, int annealing) {
  for (uint32_t c = threadIdx.x; c < P.k; c += blockDim.x) P.scale[c] = annealing ? P.ones_d / P.kvec[c] : 1.0;
}

// K5: gamma <- gammanext (with the deferred annealing rescale, :541-542), set_dir_exp(gamma)
// (src/linksampling.hh:171-187), the sweep factor b = exp(Elogpi - rowmax), and prune /
// check_and_set_converged (src/linksampling.cc:456-491).  One group per node row.
//   FROM_GACC = false: initial refresh from an uploaded gamma (no rescale, no prune; :123,:561).
template <int G, int V, bool LOGDOM, bool FROM_GACC>
__global__ void __launch_bounds__(256) k_refresh(const Params P) {
  const unsigned mask = group_mask<G>();
  const uint32_t lane = threadIdx.x & (G - 1);
  const uint32_t p = P.node_begin + (blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (p >= P.node_end) return;
  double *grow = P.gamma + (size_t)p * P.ld;
  double2 g[V];
  const bool rescale = FROM_GACC && P.tl[p] != 0.0;
  double rs = 0.0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c0 = 2u * (lane + G * j);
    if (FROM_GACC) {
      g[j] = ld_row2(P.gacc + (size_t)p * P.ld, c0, P.ld);
      if (rescale) {
        const double2 sc = ld_row2(P.scale, c0, P.ld);
        g[j].x *= sc.x;
        g[j].y *= sc.y;
      }
      if (c0 >= P.k) g[j].x = 0.0;
      if (c0 + 1u >= P.k) g[j].y = 0.0;
      st_row2(grow, c0, P.ld, g[j]);
    } else {
      g[j] = ld_row2(grow, c0, P.ld);
    }
    if (c0 < P.k) rs += g[j].x;
    if (c0 + 1u < P.k) rs += g[j].y;
  }
  rs = group_sum<G>(rs, mask);
  const double psi_sum = digamma_pos(rs);
  double2 e[V];
  double m = -CUDART_INF;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c0 = 2u * (lane + G * j);
    e[j].x = c0 < P.k ? digamma_pos(g[j].x) - psi_sum : -CUDART_INF;
    e[j].y = c0 + 1u < P.k ? digamma_pos(g[j].y) - psi_sum : -CUDART_INF;
    m = fmax(m, fmax(e[j].x, e[j].y));
  }
  double *brow = P.b + (size_t)p * P.ld;
  if (!LOGDOM) {
    m = group_max<G>(m, mask);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t c0 = 2u * (lane + G * j);
      double2 o;
      o.x = c0 < P.k ? exp(e[j].x - m) : 0.0;
      o.y = c0 + 1u < P.k ? exp(e[j].y - m) : 0.0;
      st_row2(brow, c0, P.ld, o);
    }
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t c0 = 2u * (lane + G * j);
      double2 o;
      o.x = c0 < P.k ? e[j].x : 0.0;
      o.y = c0 + 1u < P.k ? e[j].y : 0.0;
      st_row2(brow, c0, P.ld, o);
    }
  }
  if constexpr (FROM_GACC) {
  // prune: count communities with gamma - alpha >= 1 (:462-468)
  uint32_t cnt = 0, lastk = 0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c0 = 2u * (lane + G * j);
    if (c0 < P.k && g[j].x - P.alpha >= 1.0) { cnt++; lastk = c0; }
    if (c0 + 1u < P.k && g[j].y - P.alpha >= 1.0) { cnt++; lastk = c0 + 1u; }
  }
  // the active mask: word w is assembled by lane w from everybody's columns
  uint32_t total = cnt, maxk = cnt ? lastk : 0u;
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    total += __shfl_xor_sync(mask, total, o);
    maxk = max(maxk, __shfl_xor_sync(mask, maxk, o));
  }
  if (lane == 0) {   // sticky: never cleared (:472-473)
    const uint32_t was = P.conv[p];
    if (total == 1u && was == 0u) *P.conv_dirty = 1u;   // the neighbour lists that hold p are re-partitioned (k_partition)
    P.conv_next[p] = total == 1u ? maxk + 1u : was;
  }
  if (lane == 0) P.active[p] = total;
  // active bits (only meaningful for the iter > 1000 branch; cheap enough to keep current)
  for (uint32_t w = 0; w < P.words; ++w) {
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t c0 = 2u * (lane + G * j);
      if ((c0 >> 5) == w) {
        if (c0 < P.k && g[j].x - P.alpha >= 1.0) word |= 1u << (c0 & 31u);
        if (c0 + 1u < P.k && g[j].y - P.alpha >= 1.0) word |= 1u << ((c0 + 1u) & 31u);
      }
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) word |= __shfl_xor_sync(mask, word, o);
    if (lane == 0) P.abits[(size_t)p * P.words + w] = word;
  }
  }
}

// K6: held-out log-likelihood, LinkSampling::edge_likelihood (src/linksampling.hh:259-292).
// The reference's non-link form is an O(K^2) double sum of pi_p[zp]*pi_q[zq]*(1 - rate(zp,zq))
// with rate = beta_z on the diagonal and epsilon = 1e-30 off it; 1 - 1e-30 == 1 in FP64, so the
// sum equals  sum_z pi_p[z] * (S_q - pi_q[z]*beta_z)  with S_q = sum_z pi_q[z]: O(K).
// ROWS: where a node's gamma row lives -- LocalRows (the handle's own matrix) or svi::Peers (the arena of the shard
// that owns the node, svi_ls_mg.cuh).
struct LocalRows {
  const double *gamma;
  __device__ __forceinline__ const double *gamma_row(uint32_t node, uint32_t ld) const { return gamma + (size_t)node * ld; }
};
template <int G, int V, class ROWS>
__global__ void __launch_bounds__(256) k_heldout(const Params P, const ROWS rows, uint64_t npairs, const uint32_t *pp,
                                                 const uint32_t *qq, const uint8_t *yy, double epsilon,
                                                 double *out, unsigned long long *bad) {
  const unsigned mask = group_mask<G>();
  const uint32_t lane = threadIdx.x & (G - 1);
  const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (i >= npairs) return;
  const uint32_t p = pp[i], q = qq[i];
  if (p >= P.n || q >= P.n) {   // the pair list is validated here, not in a host loop: *bad = 1 + index of one bad pair
    if (lane == 0) {
      atomicMax(bad, (unsigned long long)i + 1ull);
      out[i] = CUDART_NAN;
    }
    return;
  }
  const int y = yy[i];
  double2 gp[V], gq[V];
  double sp = 0.0, sq = 0.0;
  const double *rp = rows.gamma_row(p, P.ld), *rq = rows.gamma_row(q, P.ld);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c0 = 2u * (lane + G * j);
    gp[j] = ld_row2(rp, c0, P.ld);
    gq[j] = ld_row2(rq, c0, P.ld);
    sp += gp[j].x + gp[j].y;
    sq += gq[j].x + gq[j].y;
  }
  sp = group_sum<G>(sp, mask);
  sq = group_sum<G>(sq, mask);
  double s = 0.0, piq_sum = 0.0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    gq[j].x /= sq; gq[j].y /= sq;
    piq_sum += gq[j].x + gq[j].y;
  }
  piq_sum = group_sum<G>(piq_sum, mask);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c0 = 2u * (lane + G * j);
    double2 rate = make_double2(0.0, 0.0);
    if (c0 < P.k) rate.x = P.lambda[2 * c0] / (P.lambda[2 * c0] + P.lambda[2 * c0 + 1]);
    if (c0 + 1u < P.k) rate.y = P.lambda[2 * c0 + 2] / (P.lambda[2 * c0 + 2] + P.lambda[2 * c0 + 3]);
    const double px = gp[j].x / sp, py = gp[j].y / sp;
    if (y) {
      s += px * gq[j].x * rate.x + py * gq[j].y * rate.y;
    } else {
      // diagonal term (1 - beta_z) plus off-diagonal terms (1 - epsilon)
      const double offx = (piq_sum - gq[j].x) * (1.0 - epsilon), offy = (piq_sum - gq[j].y) * (1.0 - epsilon);
      s += px * (gq[j].x * (1.0 - rate.x) + offx) + py * (gq[j].y * (1.0 - rate.y) + offy);
    }
  }
  s = group_sum<G>(s, mask);
  if (s < 1e-30) s = 1e-30;
  if (lane == 0) out[i] = log(s);
}

static __global__ void k_fill(double *p, size_t n, double v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// repack between the caller's dense [n*k] layout and the padded [n*ld] device layout
static __global__ void k_pad_rows(const double *src, double *dst, uint32_t n, uint32_t k, uint32_t ld) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n * ld;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = i % ld;
    dst[i] = c < k ? src[(i / ld) * k + c] : 0.0;
  }
}
static __global__ void k_unpad_rows(const double *src, double *dst, uint32_t n, uint32_t k, uint32_t ld) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n * k;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[(i / k) * ld + i % k];
}

}  // namespace svi
