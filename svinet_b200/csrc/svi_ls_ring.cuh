// svi_ls_ring.cuh -- the two edge sweeps (phi, s3) for 32 < K <= 256, second generation.
//
// What the first-generation kernels (svi_ls_kernels.cuh: k_phi, k_s3) left on the table
// (profiles/r01_v1_k_phi_c4_ncu_full.txt: 246 warp-instructions per neighbour, 12 resident warps/SM,
// long_scoreboard on a serialised col -> converged -> row load chain, 2.78 TB/s):
//   * rows now travel global -> shared through the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP)
//     into a per-group ring of R slots, each completed on its own mbarrier; the ring keeps R rows in
//     flight per group without spending registers on them
//   * neighbour ids and their converged flags are fetched one CHUNK (G neighbours, coalesced) ahead,
//     so the dependent-load chain is paid once per G neighbours instead of once per neighbour
//   * a group is G = 8 or 16 lanes (not a full warp): the per-neighbour reductions, reciprocal and
//     arg-max butterflies are issued once per WARP instruction and serve 32/G neighbours; the groups
//     of a warp run in lockstep over a common trip count, and when every group of the warp is on the
//     full-phi branch (the common case) the body runs warp-converged with full-mask shuffles
//   * w[k] = (b[p][k]*eb[k]) * b[q][k]: the self factor is folded once per segment, the accumulate is
//     one DFMA; ring slots are padded to the tile width and the pad is zeroed once, so row reads from
//     shared memory are unpredicated LDS.128
#pragma once
#include "svi_ls_kernels.cuh"

namespace svi {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (complete_tx)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ double2 lds2(uint32_t addr) {
  double2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
  return r;
}

template <int G>
__device__ __forceinline__ uint32_t warp_max_over_groups(uint32_t v) {
#pragma unroll
  for (int o = G; o < 32; o <<= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Per-group view of the ring.  Slot s of this group: rows + s*slot_bytes, barrier bars + 8*s.
template <int R>
struct Ring {
  uint32_t rows, bars, row_bytes, slot_bytes;
  uint32_t phase;  // bit s = parity the next wait on slot s must see
  __device__ __forceinline__ void issue(uint32_t slot, const double *src) const {
    mbar_expect_tx(bars + 8u * slot, row_bytes);
    bulk_g2s(rows + slot * slot_bytes, src, row_bytes, bars + 8u * slot);
  }
  __device__ __forceinline__ void wait(uint32_t slot) {
    const uint32_t par = (phase >> slot) & 1u;
    while (!mbar_try_wait(bars + 8u * slot, par)) {
    }
    phase ^= 1u << slot;
  }
};

enum class Sweep { Phi, S3 };

__device__ __forceinline__ uint32_t grp_of(uint32_t tid, uint32_t g) { return tid / g; }

// phi of one neighbour row sitting in shared memory at `rbase`, accumulated into acc (and the arg-max
// community into mb).  `mask` is the shuffle mask: the full warp when every group of the warp is here,
// the group's own lanes otherwise.
//
// Two passes over the shared-memory row instead of a register copy of the weights: pass 1 forms
// s = sum_k be[k]*r[k] (and, for the tally, the arg-max of the products), pass 2 re-reads the row and
// accumulates phi[k] = be[k] * (r[k] / s).  The row costs two LDS.128 per 16 bytes, but the 2V weight
// registers are gone, which is what lets four blocks (16 warps) share an SM.
template <int G, int V, bool SPARSE, bool COMM>
__device__ __forceinline__ void phi_row(const Params &P, uint32_t rbase, uint32_t lane, unsigned mask,
                                        const double2 (&be)[V], double2 (&acc)[V], uint32_t &mb, bool sparse,
                                        uint32_t p, uint32_t q, uint32_t publish) {
  // SPARSE: restrict to the union of the endpoints' active communities (:634-664); bit e of `keep`
  // is element e = 2j + {0,1} of this lane
  uint32_t keep = 0xffffffffu;
  if (SPARSE && sparse) {
    const uint32_t *ap = P.abits + (size_t)p * P.words, *aq = P.abits + (size_t)q * P.words;
    keep = 0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t c = 2u * (lane + G * j);
      uint32_t bits = 0;
      if (c < P.k) bits = (ap[c >> 5] | aq[c >> 5]) >> (c & 31u);
      keep |= (bits & 3u) << (2 * j);
    }
  }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  // arg-max of the unnormalised weights (same arg-max as phi = w/s); the FIRST maximum wins
  // (D1Array::max, src/matrix.hh:521-532).  Two independent chains (even / odd j) halve the exposed
  // compare latency; the merge keeps the lower element on ties.
  double best = 0.0, bestb = 0.0;
  uint32_t beste = 0xffffffffu, besteb = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    double2 r = lds2(rbase + 16u * (lane + G * j));
    if (SPARSE) {
      if (!((keep >> (2 * j)) & 1u)) r.x = 0.0;
      if (!((keep >> (2 * j)) & 2u)) r.y = 0.0;
    }
    if (COMM) {
      const double wx = be[j].x * r.x, wy = be[j].y * r.y;
      const bool yy = wy > wx;
      const double wm = yy ? wy : wx;
      const uint32_t we = yy ? 2 * j + 1 : 2 * j;
      if (j & 1) {
        s2 += wx; s3 += wy;
        if (wm > bestb) { bestb = wm; besteb = we; }
      } else {
        s0 += wx; s1 += wy;
        if (wm > best) { best = wm; beste = we; }
      }
    } else if (j & 1) {
      s2 = fma(be[j].x, r.x, s2);
      s3 = fma(be[j].y, r.y, s3);
    } else {
      s0 = fma(be[j].x, r.x, s0);
      s1 = fma(be[j].y, r.y, s1);
    }
  }
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(mask, s, o);
  // s == 0 only for an empty active union: phi stays all-zero (:634-664 with an empty list).  No early
  // return: other groups of the warp may share the shuffles below.
  const double inv = s > 0.0 ? 1.0 / s : 0.0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    double2 r = lds2(rbase + 16u * (lane + G * j));
    if (SPARSE) {
      if (!((keep >> (2 * j)) & 1u)) r.x = 0.0;
      if (!((keep >> (2 * j)) & 2u)) r.y = 0.0;
    }
    acc[j].x = fma(be[j].x, r.x * inv, acc[j].x);
    acc[j].y = fma(be[j].y, r.y * inv, acc[j].y);
  }
  if (COMM) {
    if (bestb > best || (bestb == best && besteb < beste)) { best = bestb; beste = besteb; }
    uint32_t bestk = beste == 0xffffffffu ? beste : 2u * (lane + G * (beste >> 1)) + (beste & 1u);
    // (Merging across the group with three redux.sync on the bit patterns -- max high word, max low word among
    // the ties, min column among the ties -- is ~10 instructions instead of ~35, but redux.sync with one mask
    // per 8-lane group serialises the groups of a warp: measured 59.3 ms against 53.0 ms for this butterfly.)
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(mask, best, o);
      const uint32_t ok = __shfl_xor_sync(mask, bestk, o);
      if (ob > best || (ob == best && ok < bestk)) { best = ob; bestk = ok; }
    }
    if (best > 0.0 && lane == (bestk >> 5)) {
      mb |= 1u << (bestk & 31u);
      // one arg-max per link: the neighbour's bit too (a 4-byte reduction in L2; mbits is 28 MB at config 4)
      if (publish) atomicOr(P.mbits + (size_t)q * P.words + lane, 1u << (bestk & 31u));
    }
  }
}


// One group per segment, 32/G segments per warp in lockstep.
//   MODE == Phi : part[seg] = sum over the segment's neighbours of phi        (K1, src/linksampling.cc:605-725)
//   MODE == S3  : block partials of s3[k] = sum mphi[p][k]*mphi[q][k]          (K3, :731-746)
// Segments [seg_first, seg_end) of the segment table.  Every segment's neighbour list is PARTITIONED by
// k_partition (below): its first seg_nnc[seg] neighbours are the not-converged ones, the rest the converged ones.
// For a not-converged node p the full-phi links (:632-718: both or neither endpoint converged) are therefore the
// first part and the one-hot shortcut links (:619-631) the second; for a converged p it is the other way round.
// The ring loop runs over the full part only -- no per-neighbour flag load, no lock-stepped pass for a shortcut link
// -- and the shortcut part is tallied G neighbours per instruction.
//   publish (COMM only; launched over the "up" segments): the link-community bit is also set for the neighbour q (one arg-max per LINK, computed on
//   the s3-owner's side only, src/linksampling.cc:704-717 sets fmap[p] and fmap[q] from one max_k)
template <int G, int V, int R, int T, int MINB, Sweep MODE, bool SPARSE, bool COMM>
__global__ void __launch_bounds__(T, MINB) k_sweep_ring(const Params P, const uint32_t seg_first, const uint32_t seg_end,
                                                        const uint32_t seg_first2, const uint32_t seg_end2,
                                                        const uint32_t publish) {
  static_assert(R <= G && (R & (R - 1)) == 0, "ring depth: power of two, at most one chunk");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int GPB = T / G;        // groups per block
  constexpr int CAP = 2 * G * V;    // columns covered by the tile (>= ld)
  const unsigned gmask = group_mask<G>();
  const uint32_t lane = threadIdx.x & (G - 1);
  const uint32_t grp = threadIdx.x / G;
  // smem: [GPB][R] slots of CAP doubles (the tail beyond ld stays zero), then [GPB][R] mbarriers, then
  // [GPB][CAP] u32 tallies of the one-hot (shortcut) links per column
  const uint32_t rows0 = smem_u32(smem_raw);
  const uint32_t bars0 = rows0 + GPB * R * CAP * 8u;
  uint32_t *hist = reinterpret_cast<uint32_t *>(smem_raw + (size_t)GPB * R * (CAP * 8u + 8u)) + grp_of(threadIdx.x, G) * CAP;
  for (uint32_t i = threadIdx.x; i < (uint32_t)(GPB * CAP); i += T)
    reinterpret_cast<uint32_t *>(smem_raw + (size_t)GPB * R * (CAP * 8u + 8u))[i] = 0u;
  Ring<R> ring;
  ring.row_bytes = P.ld * 8u;
  ring.slot_bytes = CAP * 8u;
  ring.rows = rows0 + grp * R * CAP * 8u;
  ring.bars = bars0 + grp * R * 8u;
  ring.phase = 0;
  // zero the pad columns [ld, CAP) of every slot once; the bulk copies only ever write [0, ld)
  for (uint32_t i = threadIdx.x; i < (uint32_t)(GPB * R * CAP); i += T)
    if (i % CAP >= P.ld) reinterpret_cast<double *>(smem_raw)[i] = 0.0;
  if (lane < R) mbar_init(ring.bars + 8u * lane, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const double *src_rows = MODE == Sweep::Phi ? P.b : P.mphi;
  // two ranges of the segment table in one launch (a chunk of nodes: its "lo" segments and its "up" segments)
  const uint32_t nseg1 = seg_end - seg_first, nseg = nseg1 + (seg_end2 - seg_first2);

  double2 s3acc[V];  // S3 only: per-lane column sums across this group's segments
#pragma unroll
  for (int j = 0; j < V; ++j) s3acc[j] = make_double2(0.0, 0.0);

  const uint32_t ggid = (blockIdx.x * T + threadIdx.x) / G;
  const uint32_t ngroups = gridDim.x * T / G;
  // Phi: one segment per group (grid covers nseg).  S3: persistent, grid-stride over segments.
  const uint32_t warp_first = ggid - grp % (32 / G);   // first group id of this warp
  for (uint32_t base = warp_first; base < nseg; base += ngroups) {
    const uint32_t sidx = base + grp % (32 / G);
    const bool have = sidx < nseg;
    const uint32_t seg = sidx < nseg1 ? seg_first + sidx : seg_first2 + (sidx - nseg1);
    const uint32_t p = have ? P.seg_node[seg] : 0u;
    const uint32_t beg = have ? P.seg_beg[seg] : 0u;
    const uint32_t cnt = have ? P.seg_cnt[seg] : 0u;
    const uint32_t nnc = have ? P.seg_nnc[seg] : 0u;
    const uint32_t pc = have ? P.conv[p] : 0u;
    // full-phi part [fbeg, fbeg + fcnt), shortcut part [sbeg, sbeg + scnt)
    const uint32_t fbeg = pc ? beg + nnc : beg, fcnt = pc ? cnt - nnc : nnc;
    const uint32_t sbeg = pc ? beg : beg + nnc, scnt = cnt - fcnt;
    const uint32_t cntmax = warp_max_over_groups<G>(fcnt);
    uint32_t pa = 0;
    if (SPARSE) pa = have ? P.active[p] : 0u;
    const double *self_row = src_rows + (size_t)p * P.ld;

    double2 be[V], acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t c = 2u * (lane + G * j);
      acc[j] = make_double2(0.0, 0.0);
      if (MODE == Sweep::Phi) {
        const double2 bs = have ? ld_row2(self_row, c, P.ld) : make_double2(0.0, 0.0);
        const double2 eb = ld_row2(P.eb, c, P.ld);
        be[j] = make_double2(bs.x * eb.x, bs.y * eb.y);
      }
    }
    uint32_t mb = 0;        // Phi+COMM: membership word `lane`
    double one_hot = 0.0;   // S3: shortcut mass for column pc-1

    // chunk 0 ids, then the ring prologue
    uint32_t q_cur = p, q_nxt = p;
    if (lane < fcnt) q_cur = __ldg(P.col + fbeg + lane);
    if (lane < R && lane < fcnt) ring.issue(lane, src_rows + (size_t)q_cur * P.ld);

    // ---- shortcut links (exactly one endpoint converged), while the first rows are in flight ----
    if (scnt) {
      if (MODE == Sweep::Phi) {
        // one-hot phi on the converged endpoint's community (:622-631), counted per column in shared memory
        if (pc) {
          if (lane == 0) hist[pc - 1u] += scnt;
        } else {
          for (uint32_t i = lane; i < scnt; i += G) {
            const uint32_t qc = P.conv[__ldg(P.col + sbeg + i)];
            if (qc) atomicAdd(hist + (qc - 1u), 1u);
          }
        }
      } else if (pc) {
        // s3[pc-1] += mphi[q][pc]  (sic, :739-740; column K reads as 0): per-lane partial sums, fixed-order merge
        double oh = 0.0;
        if (pc < P.k)
          for (uint32_t i = lane; i < scnt; i += G) oh += P.mphi[(size_t)__ldg(P.col + sbeg + i) * P.ld + pc];
        one_hot = group_sum<G>(oh, gmask);
      } else {
        // s3[qc-1] += mphi[p][qc]  (:741-742): the same addend for every neighbour converged to qc, so only
        // counted here and applied once per segment
        for (uint32_t i = lane; i < scnt; i += G) {
          const uint32_t qc = P.conv[__ldg(P.col + sbeg + i)];
          if (qc) atomicAdd(hist + (qc - 1u), 1u);
        }
      }
    }
    __syncwarp();

    for (uint32_t c0 = 0; c0 < cntmax; c0 += G) {
      if (c0 > 0) q_cur = q_nxt;
      q_nxt = c0 + G + lane < fcnt ? __ldg(P.col + fbeg + c0 + G + lane) : p;
      const uint32_t ulim = min((uint32_t)G, cntmax - c0);
      // one copy of the body: unrolled G times with two phi_row instances each, the sweep is ~150 KB of
      // SASS and runs out of the instruction cache
#pragma unroll 1
      for (uint32_t u = 0; u < ulim; ++u) {
        const uint32_t i = c0 + u;
        // control flow is warp-uniform here: full-mask shuffles, each confined to its G-lane segment
        const uint32_t q = __shfl_sync(0xffffffffu, q_cur, u, G);
        const bool live = i < fcnt;
        const uint32_t slot = i & (R - 1);
        const bool alllive = __all_sync(0xffffffffu, live);
        const uint32_t rbase = ring.rows + slot * ring.slot_bytes;
        if (live) {
          ring.wait(slot);
          if (MODE == Sweep::S3) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
              const double2 r = lds2(rbase + 16u * (lane + G * j));
              acc[j].x += r.x;
              acc[j].y += r.y;
            }
          } else {
            bool sparse = false;
            if (SPARSE) sparse = pa < P.k_div10 && P.active[q] < P.k_div10;
            phi_row<G, V, SPARSE, COMM>(P, rbase, lane, alllive ? 0xffffffffu : gmask, be, acc, mb, sparse, p, q,
                                        publish);
          }
        }
        // refill this slot with neighbour i+R (its id sits in the current or the next chunk).  Ordering of the
        // slot's reuse: every lane's generic-proxy reads of the slot (lds2 above, results consumed) precede
        // __syncwarp; lane 0 issues the async-proxy write after it.  DESIGN.md section 4, "ring slot reuse".
        __syncwarp();
        const uint32_t ua = u + R;
        const uint32_t qa = __shfl_sync(0xffffffffu, ua < (uint32_t)G ? q_cur : q_nxt, ua & (G - 1), G);
        if (lane == 0 && i + R < fcnt) ring.issue(slot, src_rows + (size_t)qa * P.ld);
      }
    }

    __syncwarp();   // the shared-memory tallies -> every lane
    if (MODE == Sweep::Phi) {
      if (have) {
        double *out = P.part + (size_t)seg * P.ld;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const uint2 h = *reinterpret_cast<const uint2 *>(hist + 2u * (lane + G * j));
          acc[j].x += (double)h.x;     // exact: integers, same as the reference's repeated += 1
          acc[j].y += (double)h.y;
          st_row2(out, 2u * (lane + G * j), P.ld, acc[j]);
        }
        if (COMM && mb) atomicOr(P.mbits + (size_t)p * P.words + lane, mb);
      }
    } else if (have) {
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const uint32_t c = 2u * (lane + G * j);
        const double2 mp = ld_row2(self_row, c, P.ld);
        s3acc[j].x = fma(mp.x, acc[j].x, s3acc[j].x);
        s3acc[j].y = fma(mp.y, acc[j].y, s3acc[j].y);
        if (pc && pc - 1u == c) s3acc[j].x += one_hot;
        if (pc && pc - 1u == c + 1u) s3acc[j].y += one_hot;
        // neighbours converged to community c (resp. c+1): count x mphi[p][c+1] (resp. [c+2]), sic (Q4);
        // column K reads as 0
        uint2 *hp = reinterpret_cast<uint2 *>(hist + c);
        const uint2 h = *hp;
        if (h.x | h.y) {
          const double vx = c + 1u < P.k ? self_row[c + 1u] : 0.0, vy = c + 2u < P.k ? self_row[c + 2u] : 0.0;
          s3acc[j].x = fma((double)h.x, vx, s3acc[j].x);
          s3acc[j].y = fma((double)h.y, vy, s3acc[j].y);
          *hp = make_uint2(0u, 0u);   // this group's next segment starts from zero
        }
      }
      __syncwarp(gmask);   // (`have` differs between the groups of a warp: group mask, not the full warp)
    }
  }
  if (MODE == Sweep::S3) {
    // all rings are drained here (every issued copy was waited on), so the rows area can be reused
    block_reduce_columns<G, V>(s3acc, reinterpret_cast<double *>(smem_raw),
                               P.kpart + (size_t)blockIdx.x * CAP);
  }
}

// Stable partition of every segment's neighbour list by the neighbours' converged flags: not-converged first,
// seg_nnc[seg] = how many.  `converged` is sticky (src/linksampling.cc:472-473), so the lists change only in
// sweeps where some node newly converged; *P.conv_dirty says so (set by k_refresh / svi_ls_set_converged, cleared
// by the caller after this kernel) and lets the launch return at once otherwise.  One warp per segment, ids staged
// in shared memory (segments hold at most kMaxSegLen neighbours), in place.
constexpr uint32_t kMaxSegLen = 1024;
static __global__ void __launch_bounds__(256) k_partition(const Params P, const uint32_t force) {
  __shared__ uint32_t stage[8][kMaxSegLen];
  if (!force && *P.conv_dirty == 0u) return;
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint32_t nwarps = gridDim.x * 8u;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t *col = const_cast<uint32_t *>(P.col);
  uint32_t *st = stage[w];
  for (uint32_t seg = blockIdx.x * 8u + w; seg < P.nseg; seg += nwarps) {
    const uint32_t beg = P.seg_beg[seg], cnt = P.seg_cnt[seg];
    // three passes with INDEPENDENT loads inside each (the first version chained id -> flag -> ballot per 32
    // neighbours and was latency-bound: ~2.5 ms at config 4 whenever a node had converged): ids, then their flags
    // (kept in bit 31 of the staged id: node ids are < 2^31), then the scatter
    for (uint32_t i = lane; i < cnt; i += 32u) st[i] = col[beg + i];
    __syncwarp();
    uint32_t mine = 0;
    for (uint32_t i = lane; i < cnt; i += 32u) {
      const uint32_t q = st[i];
      const uint32_t nc = P.conv[q] == 0u ? 1u : 0u;
      st[i] = q | (nc << 31);
      mine += nc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    const uint32_t nnc = mine;
    __syncwarp();
    if (nnc != P.seg_nnc[seg] || force) {   // (a segment none of whose neighbours changed keeps its order)
      uint32_t a = 0, b = nnc;
      for (uint32_t i0 = 0; i0 < cnt; i0 += 32u) {
        const bool in = i0 + lane < cnt;
        const uint32_t v = in ? st[i0 + lane] : 0u;
        const bool nc = in && (v >> 31);
        const uint32_t mn = __ballot_sync(0xffffffffu, nc), mc = __ballot_sync(0xffffffffu, in && !nc);
        if (in) col[beg + (nc ? a + __popc(mn & lt) : b + __popc(mc & lt))] = v & 0x7fffffffu;
        a += __popc(mn);
        b += __popc(mc);
      }
      if (lane == 0) P.seg_nnc[seg] = nnc;
    }
    __syncwarp();
  }
}

}  // namespace svi
