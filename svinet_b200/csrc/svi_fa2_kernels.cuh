// svi_fa2_kernels.cuh -- sm_100a kernels of the `-rnode -stratified` iteration (class FastAMM2).
//
// What one iteration is (src/fastamm2.cc:566-640): a minibatch of node pairs that all contain one
// start node (its links, or a "non-informative" set of n/10 non-links); every pair runs a two-phi
// coordinate ascent (PhiCompute::update_phis_until_conv, src/fastamm2.hh:151-209, <= 50 rounds of two
// K-wide softmaxes); phi of the far endpoint becomes that node's gammat row, phi of the start node and
// phi1*phi2 are summed over the minibatch; then EVERY gamma row is blended (touched rows towards
// alpha + scale*gammat, all other rows decay towards alpha) and lambda is blended.
//
// Device formulation:
//   k_fa2_prep   one block: Elogbeta = psi(lambda) - psi(sum), Elogf = column `type` of it, Elogpi of
//                the start node (the far endpoints' Elogpi rows are formed by the pair kernel from
//                their gamma rows: no N x K Elogpi matrix exists)
//   k_fa2_pairs  persistent; a GROUP of G lanes owns one pair, the whole fixed point runs in registers;
//                the far endpoint is touched by exactly one pair, so its Robbins-Monro blend is done
//                right there; start-node phi and phi1*phi2 go to per-group shared-memory accumulators
//                and leave as per-block partial rows (fixed order: run-to-run deterministic)
//   k_fa2_blend  every other row: decay towards alpha (the bandwidth-bound pass, 2 x N x K x 8 bytes);
//                the start row gets the reduced partials.  EAGER mode only.  In the default LAZY mode the
//                decay (gamma - alpha) *= (1 - rho) of the untouched rows -- the same factor for every row,
//                because the reference bumps every node's counter every iteration -- is carried as ONE scalar:
//                rows store u with gamma = alpha + c*u, an iteration multiplies c by (1 - rho), touched rows
//                get u += rho*scale*gammat / c, and the O(N*K) pass disappears (k_fa2_fold re-bases u <- c*u
//                every few 1e4 iterations, before c can underflow)
//   k_fa2_lambda one block: reduce the phi1*phi2 partials, blend lambda, advance the node counter
//   k_fa2_draw_* the minibatch itself, from a Philox4x32-10 stream keyed by (seed, iter)
#pragma once
#include "svi_ls_kernels.cuh"

namespace svi {

struct Fa2Ctrl {            // one iteration's control block, device resident
  uint32_t type, start, npairs, iter;
  uint64_t sampled_inc;     // the reference's _total_pairs_sampled increment
  double rho_node, rho_t, scale;
  double nodec;             // _nodec[i] (identical for every node: each is bumped every iteration)
  double cscale;            // lazy decay: gamma = alpha + cscale * stored (1 in eager mode)
  uint64_t total_sampled, total_rounds, last_rounds;
};

struct Fa2Params {
  uint32_t n, k, ld;
  double alpha, eta0, eta1, logeps, thresh, epsilon;
  double tau0, kappa, nodetau0, nodekappa, inf_epsilon;
  uint32_t online_iters, m_sets, nolambda;
  uint32_t lazy;            // 1: rows hold u with gamma = alpha + cscale*u, the decay of untouched rows is the scalar
                            //    update cscale *= (1 - rho) and no O(N*K) pass runs; 0: rows hold gamma, k_fa2_blend runs
  double *gamma;            // [n*ld]  (lazy: u)
  double *lambda;           // [k*2]
  double *elogbeta;         // [2*ld]  column 0 then column 1
  double *elogf;            // [ld]    -inf beyond k
  double *epi_start;        // [ld]    -inf beyond k
  uint32_t *pairs;          // [2*cap_pairs]
  uint32_t cap_pairs;
  uint8_t *touched;         // [n]
  Fa2Ctrl *ctrl;
  double *partS, *partL;    // [pair_blocks * cap] block partials
  double *wide;             // K > 512 (svi_fa2_wide.cuh): scratch rows of the pair blocks, else null
  uint32_t pair_blocks;
  // graph for device-side draws
  const uint64_t *adj_off;  // [n+1]
  const uint32_t *adj;      // sorted neighbour lists
  const uint64_t *heldout;  // sorted (p<<32|q)
  uint64_t nheldout;
  const uint32_t *shuffled; // [n]
};

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (c0,c1,c2,c3), key = (k0,k1) -------------------
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// ---- K-wide softmax of t[] over the group, in place: t <- exp(t - logsumexp(t)) -------------------
// (D1Array::lognormalize, src/matrix.hh:296-318, evaluated as max / sum-of-exp instead of the running
// pairwise log-sum; pad columns hold -inf and come out as 0)
template <int G, int V>
__device__ __forceinline__ void group_softmax(double2 (&t)[V], unsigned mask) {
  double m = -CUDART_INF;
#pragma unroll
  for (int j = 0; j < V; ++j) m = fmax(m, fmax(t[j].x, t[j].y));
  m = group_max<G>(m, mask);
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    t[j].x = exp(t[j].x - m);
    t[j].y = exp(t[j].y - m);
    s += t[j].x + t[j].y;
  }
  s = group_sum<G>(s, mask);
  const double inv = 1.0 / s;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    t[j].x *= inv;
    t[j].y *= inv;
  }
}

// stored row element -> gamma: eager (cs, a0) = (1, 0), exact; lazy (cs, a0) = (cscale, alpha)
struct Fa2Map {
  double cs, a0;
  __device__ __forceinline__ double operator()(double u) const { return fma(cs, u, a0); }
};
__device__ __forceinline__ Fa2Map fa2_map(const Fa2Params &P, const Fa2Ctrl &c) {
  Fa2Map m;
  m.cs = P.lazy ? c.cscale : 1.0;
  m.a0 = P.lazy ? P.alpha : 0.0;
  return m;
}

// Elogpi row of node a from its gamma row (FastAMM2::set_dir_exp(a,..), src/fastamm2.hh:424-435);
// pad columns -> -inf
template <int G, int V>
__device__ __forceinline__ void elogpi_row(const Fa2Params &P, const Fa2Map &gm, uint32_t a, uint32_t lane,
                                           unsigned mask, double2 (&e)[V]) {
  const double *row = P.gamma + (size_t)a * P.ld;
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c = 2u * (lane + G * j);
    e[j] = ld_row2(row, c, P.ld);
    e[j].x = c < P.k ? gm(e[j].x) : 0.0;
    e[j].y = c + 1u < P.k ? gm(e[j].y) : 0.0;
    s += e[j].x + e[j].y;
  }
  s = group_sum<G>(s, mask);
  const double psi_sum = digamma_pos(s);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c = 2u * (lane + G * j);
    e[j].x = c < P.k ? digamma_pos(e[j].x) - psi_sum : -CUDART_INF;
    e[j].y = c + 1u < P.k ? digamma_pos(e[j].y) - psi_sum : -CUDART_INF;
  }
}

// The coordinate ascent of one pair.  ep/eq: Elogpi rows of p and q; ef: Elogf; on return phi1/phi2.
// Returns the number of rounds.  Both sides of a round read the OLD phis (update_phis does not write
// _phi1/_phi2 when phifix is false, src/fastamm2.hh:129-135); convergence is tested on odd rounds against
// the phis of two rounds earlier (:170-199).
template <int G, int V>
__device__ __forceinline__ uint32_t fa2_pair_core(const Fa2Params &P, unsigned mask, int y,
                                                  const double2 (&ep)[V], const double2 (&eq)[V],
                                                  const double2 (&ef)[V], uint32_t lane,
                                                  double2 (&phi1)[V], double2 (&phi2)[V]) {
  const double u0 = 1.0 / (double)P.k;
  double2 old1[V], old2[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c = 2u * (lane + G * j);
    phi1[j] = make_double2(c < P.k ? u0 : 0.0, c + 1u < P.k ? u0 : 0.0);
    phi2[j] = phi1[j];
    old1[j] = old2[j] = make_double2(0.0, 0.0);
  }
  const double inv_k = 1.0 / (double)P.k;   // mean() divides by k; x/k vs x*(1/k) only matters at the threshold's last ulp
  uint32_t rounds = 0;
  for (uint32_t i = 0; i < P.online_iters; ++i) {
    if ((i & 1u) == 0u) {
#pragma unroll
      for (int j = 0; j < V; ++j) { old1[j] = phi1[j]; old2[j] = phi2[j]; }
    }
    double2 t1[V], t2[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      // anext[k] = Elogpi[c][k] + Elogf[k]*b[k] + [y=1](1-b[k])*log(eps)      (src/fastamm2.hh:112-118)
      double ux = 0.0, uy = 0.0, wx = 0.0, wy = 0.0;
      if (y) {
        ux = (1.0 - phi2[j].x) * P.logeps; uy = (1.0 - phi2[j].y) * P.logeps;
        wx = (1.0 - phi1[j].x) * P.logeps; wy = (1.0 - phi1[j].y) * P.logeps;
      }
      t1[j].x = ep[j].x + ef[j].x * phi2[j].x + ux;
      t1[j].y = ep[j].y + ef[j].y * phi2[j].y + uy;
      t2[j].x = eq[j].x + ef[j].x * phi1[j].x + wx;
      t2[j].y = eq[j].y + ef[j].y * phi1[j].y + wy;
    }
    group_softmax<G, V>(t1, mask);
    group_softmax<G, V>(t2, mask);
    double d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      d1 += fabs(t1[j].x - old1[j].x) + fabs(t1[j].y - old1[j].y);
      d2 += fabs(t2[j].x - old2[j].x) + fabs(t2[j].y - old2[j].y);
      phi1[j] = t1[j];
      phi2[j] = t2[j];
    }
    rounds++;
    if ((i & 1u) == 0u) continue;
    d1 = group_sum<G>(d1, mask);
    d2 = group_sum<G>(d2, mask);
    if (d1 * inv_k < P.thresh && d2 * inv_k < P.thresh) break;
  }
  return rounds;
}

template <int G, int V>
__device__ __forceinline__ void load_kvec(const double *v, uint32_t lane, uint32_t ld, double2 (&o)[V], double pad) {
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c = 2u * (lane + G * j);
    o[j] = c < ld ? *reinterpret_cast<const double2 *>(v + c) : make_double2(pad, pad);
  }
}

// ---- prep: expectations shared by the whole minibatch ---------------------------------------------
template <int G, int V>
__global__ void __launch_bounds__(128) k_fa2_prep(const Fa2Params P) {
  const Fa2Ctrl c = *P.ctrl;
  if (threadIdx.x == 0) P.ctrl->last_rounds = 0;
  for (uint32_t z = threadIdx.x; z < P.ld; z += blockDim.x) {
    double e0 = -CUDART_INF, e1 = -CUDART_INF;
    if (z < P.k) {
      // FastAMM2::set_dir_exp(lambda, Elogbeta), src/fastamm2.hh:401-422 (non-positive -> alpha)
      const double l0 = P.lambda[2 * z], l1 = P.lambda[2 * z + 1];
      const double ps = digamma_pos(l0 + l1);
      e0 = digamma_pos(l0 <= 0.0 ? P.alpha : l0) - ps;
      e1 = digamma_pos(l1 <= 0.0 ? P.alpha : l1) - ps;
    }
    P.elogbeta[z] = e0;
    P.elogbeta[P.ld + z] = e1;
    // compute_Elogf (src/fastamm2.hh:139-149): y = 1 -> column 0, y = 0 -> column 1
    P.elogf[z] = z < P.k ? (c.type == 0 ? e0 : e1) : 0.0;
  }
  if (threadIdx.x < G) {   // Elogpi of the start node, by the first group
    const unsigned mask = group_mask<G>();
    double2 e[V];
    elogpi_row<G, V>(P, fa2_map(P, c), c.start, threadIdx.x, mask, e);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t col = 2u * (threadIdx.x + G * j);
      if (col < P.ld) *reinterpret_cast<double2 *>(P.epi_start + col) = e[j];
    }
  }
}

// ---- the pairs -------------------------------------------------------------------------------------
#ifndef SVI_FA2_MINB
#define SVI_FA2_MINB 1
#endif
template <int G, int V, int T>
__global__ void __launch_bounds__(T, SVI_FA2_MINB) k_fa2_pairs(const Fa2Params P) {
  extern __shared__ double smem[];
  constexpr int CAP = 2 * G * V;
  constexpr int GPB = T / G;
  double *accS = smem;                 // [GPB][CAP] start-node phi, summed over this group's pairs
  double *accL = smem + GPB * CAP;     // [GPB][CAP] phi1*phi2
  const unsigned mask = group_mask<G>();
  const uint32_t lane = threadIdx.x & (G - 1), grp = threadIdx.x / G;
  const Fa2Ctrl c = *P.ctrl;
  const int y = c.type == 0 ? 1 : 0;
  for (uint32_t i = threadIdx.x; i < (uint32_t)(2 * GPB * CAP); i += T) smem[i] = 0.0;
  __syncthreads();
  double *myS = accS + grp * CAP, *myL = accL + grp * CAP;

  const Fa2Map gm = fa2_map(P, c);
  // lazy: u_a += rho*scale*phi_a / c', c' = (1 - rho) * c the scale AFTER this iteration
  const double lazy_coef = c.rho_node * c.scale / ((1.0 - c.rho_node) * c.cscale);
  double2 ef[V], es[V];
  load_kvec<G, V>(P.elogf, lane, P.ld, ef, 0.0);
  load_kvec<G, V>(P.epi_start, lane, P.ld, es, -CUDART_INF);
  uint32_t my_rounds = 0;
  const uint32_t ngroups = gridDim.x * GPB;
  for (uint32_t i = blockIdx.x * GPB + grp; i < c.npairs; i += ngroups) {
    const uint32_t p = P.pairs[2 * i], q = P.pairs[2 * i + 1];
    const bool start_is_p = p == c.start;
    const uint32_t a = start_is_p ? q : p;
    double2 ep[V], eq[V], phi1[V], phi2[V];
    elogpi_row<G, V>(P, gm, a, lane, mask, ep);
    // one call site (the fixed point is ~40 KB of SASS with its inlined exp's): order the two Elogpi rows here
#pragma unroll
    for (int j = 0; j < V; ++j) {
      eq[j] = start_is_p ? ep[j] : es[j];
      ep[j] = start_is_p ? es[j] : ep[j];
    }
    const uint32_t r = fa2_pair_core<G, V>(P, mask, y, ep, eq, ef, lane, phi1, phi2);
    if (lane == 0) my_rounds += r;
    // far endpoint: gammat[a] = its phi, touched exactly once -> blend here (src/fastamm2.cc:609-613)
    double *grow = P.gamma + (size_t)a * P.ld;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint32_t col = 2u * (lane + G * j);
      const double2 pa = start_is_p ? phi2[j] : phi1[j];
      const double2 ps = start_is_p ? phi1[j] : phi2[j];
      if (col < P.ld) {
        double2 g = *reinterpret_cast<const double2 *>(grow + col);
        if (P.lazy) {
          g.x = col < P.k ? fma(lazy_coef, pa.x, g.x) : 0.0;
          g.y = col + 1u < P.k ? fma(lazy_coef, pa.y, g.y) : 0.0;
        } else {
          g.x = col < P.k ? (1.0 - c.rho_node) * g.x + c.rho_node * (P.alpha + c.scale * pa.x) : 0.0;
          g.y = col + 1u < P.k ? (1.0 - c.rho_node) * g.y + c.rho_node * (P.alpha + c.scale * pa.y) : 0.0;
        }
        *reinterpret_cast<double2 *>(grow + col) = g;
      }
      myS[col] += ps.x;
      myS[col + 1] += ps.y;
      myL[col] += phi1[j].x * phi2[j].x;       // lambdat[k][t] += phi1*phi2 (src/fastamm2.cc:994-996)
      myL[col + 1] += phi1[j].y * phi2[j].y;
    }
    if (lane == 0 && !P.lazy) P.touched[a] = 1;
  }
  __syncthreads();
  for (uint32_t col = threadIdx.x; col < (uint32_t)CAP; col += T) {
    double s = 0.0, l = 0.0;
    for (int g = 0; g < GPB; ++g) { s += accS[g * CAP + col]; l += accL[g * CAP + col]; }
    P.partS[(size_t)blockIdx.x * CAP + col] = s;
    P.partL[(size_t)blockIdx.x * CAP + col] = l;
  }
  if (lane == 0 && my_rounds) atomicAdd((unsigned long long *)&P.ctrl->last_rounds, (unsigned long long)my_rounds);
}

// ---- blend of all the rows the pair kernel did not touch (src/fastamm2.cc:605-624) -----------------
template <int G, int V, int T>
__global__ void __launch_bounds__(T) k_fa2_blend(const Fa2Params P) {
  constexpr int CAP = 2 * G * V;
  constexpr int U = 4;                       // rows in flight per group: a row is only V 16-byte loads per lane
  const uint32_t lane = threadIdx.x & (G - 1);
  const uint32_t grp = (blockIdx.x * T + threadIdx.x) / G, ngroups = gridDim.x * T / G;
  const Fa2Ctrl c = *P.ctrl;
  for (uint32_t base = grp; base < P.n; base += ngroups * U) {
    uint32_t row[U];
    bool live[U];
    double2 g[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      row[u] = base + u * ngroups;
      live[u] = row[u] < P.n && !P.touched[row[u]];
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const uint32_t col = 2u * (lane + G * j);
        g[u][j] = live[u] && col < P.ld ? __ldcs(reinterpret_cast<const double2 *>(P.gamma + (size_t)row[u] * P.ld + col))
                                        : make_double2(0.0, 0.0);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (row[u] >= P.n) continue;
      if (!live[u]) {                        // blended by the pair kernel: just clear the flag
        if (lane == 0) P.touched[row[u]] = 0;
        continue;
      }
      const bool is_start = row[u] == c.start;
      double *grow = P.gamma + (size_t)row[u] * P.ld;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const uint32_t col = 2u * (lane + G * j);
        if (col >= P.ld) continue;
        double2 t = make_double2(0.0, 0.0);
        if (is_start) {
          // (only the pair-kernel blocks that had pairs: the others' partials are zero)
          const uint32_t active = min(P.pair_blocks, (c.npairs + (uint32_t)(T / G) - 1u) / (uint32_t)(T / G));
          for (uint32_t b = 0; b < active; ++b) {
            const double2 v = *reinterpret_cast<const double2 *>(P.partS + (size_t)b * CAP + col);
            t.x += v.x;
            t.y += v.y;
          }
        }
        const double tx = is_start ? P.alpha + c.scale * t.x : P.alpha;
        const double ty = is_start ? P.alpha + c.scale * t.y : P.alpha;
        double2 o;
        o.x = col < P.k ? (1.0 - c.rho_node) * g[u][j].x + c.rho_node * tx : 0.0;
        o.y = col + 1u < P.k ? (1.0 - c.rho_node) * g[u][j].y + c.rho_node * ty : 0.0;
        __stcs(reinterpret_cast<double2 *>(grow + col), o);
      }
    }
  }
}

// ---- lambda blend (src/fastamm2.cc:626-638) + counters ----------------------------------------------
// One block.  The column sums over the pair kernel's block partials are taken by one WARP per column (lane l adds
// the blocks l, l+32, ... in order, then a fixed butterfly) over the blocks that had pairs only: the first version
// walked all 296 partials in one thread per column, 250 us per launch at K = 4.
static __global__ void __launch_bounds__(256) k_fa2_lambda(const Fa2Params P, uint32_t cap, uint32_t groups_per_block) {
  const Fa2Ctrl c = *P.ctrl;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const uint32_t active = min(P.pair_blocks, (c.npairs + groups_per_block - 1u) / groups_per_block);
  const double coef = c.rho_node * c.scale / ((1.0 - c.rho_node) * c.cscale);
  double *urow = P.gamma + (size_t)c.start * P.ld;
  for (uint32_t z = warp; z < P.k; z += nwarps) {
    double s = 0.0, l = 0.0;
    for (uint32_t b = lane; b < active; b += 32u) {
      s += P.partS[(size_t)b * cap + z];
      l += P.partL[(size_t)b * cap + z];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      l += __shfl_xor_sync(0xffffffffu, l, o);
    }
    if (lane == 0) {
      // lazy mode, the start row: u += rho*scale*(sum of its phis) / c'   (eager mode: k_fa2_blend does it)
      if (P.lazy) urow[z] = fma(coef, s, urow[z]);
      if (!P.nolambda)
        for (uint32_t t = 0; t < 2; ++t) {
          const double raw = t == c.type ? l : 0.0;            // phi1*phi2*(t==0 ? y : 1-y)
          const double ldt = (t == 0 ? P.eta0 : P.eta1) + c.scale * raw;
          P.lambda[2 * z + t] = (1.0 - c.rho_t) * P.lambda[2 * z + t] + c.rho_t * ldt;
        }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Fa2Ctrl *w = P.ctrl;
    w->nodec = c.nodec + 1.0;
    if (P.lazy) w->cscale = (1.0 - c.rho_node) * c.cscale;
    w->total_sampled = c.total_sampled + c.sampled_inc;
    w->total_rounds = c.total_rounds + c.last_rounds;
  }
}

// ---- device-side minibatch draw -----------------------------------------------------------------------
__device__ __forceinline__ bool fa2_is_link(const Fa2Params &P, uint32_t a, uint32_t b) {
  uint64_t lo = P.adj_off[a], hi = P.adj_off[a + 1];
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    const uint32_t v = P.adj[mid];
    if (v == b) return true;
    if (v < b) lo = mid + 1; else hi = mid;
  }
  return false;
}
__device__ __forceinline__ bool fa2_is_heldout(const Fa2Params &P, uint32_t a, uint32_t b) {
  const uint64_t key = ((uint64_t)min(a, b) << 32) | max(a, b);
  uint64_t lo = 0, hi = P.nheldout;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    const uint64_t v = P.heldout[mid];
    if (v == key) return true;
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return false;
}

// The draw runs as three launches so that the candidate tests (binary searches in the start node's
// adjacency and in the held-out set: dependent global loads) spread over many blocks:
//   k_fa2_draw_count : decide (type, start, first candidate) from the Philox stream; every block tests its
//                      1024 candidates of the WINDOW, keeps the validity ballots and its count
//   k_fa2_draw_emit  : block prefix over the counts, order-preserving write of the first `want` valid pairs,
//                      control block
//   k_fa2_draw_tail  : one block, only when the window held fewer than `want` valid candidates (a start
//                      node with more excluded partners than the window's slack): continues serially
// Sampling rules:
//   type 0: the start node's neighbours minus held-out pairs                  (src/fastamm2.cc:943-960)
//   type 1: walk the shuffled node order from a random block boundary, keep non-links that are not held
//           out, until n/m nodes are collected                                (src/fastamm2.cc:1095-1125)
struct Fa2Draw {
  uint32_t type, start, q0, want;
  uint64_t ncand;          // candidates available (type 0: degree; type 1: one lap = n)
};

__device__ __forceinline__ Fa2Draw fa2_draw_header(const Fa2Params &P, uint32_t iter, uint32_t seed_lo, uint32_t seed_hi) {
  uint32_t r[4] = {iter, 0u, 0u, 0u};
  philox4x32_10(r, seed_lo, seed_hi);
  Fa2Draw d;
  d.type = (double)r[0] * (1.0 / 4294967296.0) < P.inf_epsilon ? 1u : 0u;   // gsl_ran_bernoulli
  d.start = __umulhi(r[1], P.n);
  const uint32_t setsize = (uint32_t)((double)P.n / (double)P.m_sets);
  if (d.type == 0) {
    d.q0 = 0;
    d.ncand = P.adj_off[d.start + 1] - P.adj_off[d.start];
    d.want = 0xffffffffu;
  } else {
    d.q0 = setsize ? (__umulhi(r[2], P.n) / setsize) * setsize : 0u;
    d.ncand = P.n;
    d.want = setsize;
  }
  return d;
}

__device__ __forceinline__ bool fa2_candidate(const Fa2Params &P, const Fa2Draw &d, uint64_t ci, uint32_t *node) {
  if (ci >= d.ncand) return false;
  if (d.type == 0) {
    *node = P.adj[P.adj_off[d.start] + ci];
    return !fa2_is_heldout(P, d.start, *node);
  }
  *node = P.shuffled[(d.q0 + ci) % P.n];
  return *node != d.start && !fa2_is_link(P, d.start, *node) && !fa2_is_heldout(P, d.start, *node);
}

// window of block b: candidates [b*1024, (b+1)*1024)
static __global__ void __launch_bounds__(1024) k_fa2_draw_count(const Fa2Params P, uint32_t iter, uint32_t seed_lo,
                                                                uint32_t seed_hi, uint32_t *ballots, uint32_t *counts) {
  __shared__ uint32_t warp_tot[32];
  const Fa2Draw d = fa2_draw_header(P, iter, seed_lo, seed_hi);
  const uint32_t lanei = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint64_t ci = (uint64_t)blockIdx.x * 1024u + threadIdx.x;
  uint32_t node = 0;
  const bool ok = fa2_candidate(P, d, ci, &node);
  const uint32_t bal = __ballot_sync(0xffffffffu, ok);
  if (lanei == 0) {
    ballots[blockIdx.x * 32u + warp] = bal;
    warp_tot[warp] = __popc(bal);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 32; ++w) t += warp_tot[w];
    counts[blockIdx.x] = t;
  }
}

static __global__ void __launch_bounds__(1024) k_fa2_draw_emit(const Fa2Params P, uint32_t iter, uint32_t seed_lo,
                                                               uint32_t seed_hi, const uint32_t *ballots,
                                                               const uint32_t *counts, uint32_t nblocks) {
  __shared__ uint32_t s_prefix;
  const Fa2Draw d = fa2_draw_header(P, iter, seed_lo, seed_hi);
  const uint32_t lanei = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (warp == 0) {   // valid candidates in the blocks before this one
    uint32_t t = 0;
    for (uint32_t b = lanei; b < blockIdx.x; b += 32) t += counts[b];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lanei == 0) s_prefix = t;
  }
  __syncthreads();
  uint32_t pos = s_prefix;
  for (uint32_t w = 0; w < warp; ++w) pos += __popc(ballots[blockIdx.x * 32u + w]);
  const uint32_t bal = ballots[blockIdx.x * 32u + warp];
  pos += __popc(bal & ((1u << lanei) - 1u));
  if (((bal >> lanei) & 1u) && pos < d.want && pos < P.cap_pairs) {
    const uint64_t ci = (uint64_t)blockIdx.x * 1024u + threadIdx.x;
    const uint32_t node = d.type == 0 ? P.adj[P.adj_off[d.start] + ci] : P.shuffled[(d.q0 + ci) % P.n];
    P.pairs[2 * pos] = min(d.start, node);
    P.pairs[2 * pos + 1] = max(d.start, node);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    uint32_t tot = 0;
    for (uint32_t b = 0; b < nblocks; ++b) tot += counts[b];
    Fa2Ctrl *w = P.ctrl;
    const uint32_t np = min(min(tot, d.want), P.cap_pairs);
    w->type = d.type;
    w->start = d.start;
    w->npairs = np;          // k_fa2_draw_tail may still raise it
    w->iter = iter;
    w->sampled_inc = d.type == 0 ? d.ncand : np;
    w->last_rounds = 0;
    w->rho_node = pow(P.nodetau0 + w->nodec, -1.0 * P.nodekappa);              // :606
    w->rho_t = pow(P.tau0 + ((double)iter + 1.0), -1.0 * P.kappa);             // :627, _lambda_start_iter = 0
    w->scale = d.type == 0 ? (double)P.n / (2.0 * (1.0 - P.inf_epsilon))       // :591-592
                           : ((double)P.n * (double)P.m_sets) / (2.0 * P.inf_epsilon);
  }
}

// serial continuation past the window (rare): same chunked block scan as a single-block draw
static __global__ void __launch_bounds__(1024) k_fa2_draw_tail(const Fa2Params P, uint32_t iter, uint32_t seed_lo,
                                                               uint32_t seed_hi, uint32_t nblocks) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t s_base, s_done;
  const Fa2Draw d = fa2_draw_header(P, iter, seed_lo, seed_hi);
  const uint64_t window = (uint64_t)nblocks * 1024u;
  if (P.ctrl->npairs >= d.want || window >= d.ncand) return;   // the window sufficed (uniform across the block)
  const uint32_t lanei = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_base = P.ctrl->npairs; s_done = 0; }
  __syncthreads();
  for (uint64_t c0 = window; c0 < d.ncand; c0 += 1024u) {
    uint32_t node = 0;
    const bool ok = fa2_candidate(P, d, c0 + threadIdx.x, &node);
    const uint32_t bal = __ballot_sync(0xffffffffu, ok);
    if (lanei == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    uint32_t pos = s_base + __popc(bal & ((1u << lanei) - 1u));
    for (uint32_t w = 0; w < warp; ++w) pos += warp_tot[w];
    if (ok && pos < d.want && pos < P.cap_pairs) {
      P.pairs[2 * pos] = min(d.start, node);
      P.pairs[2 * pos + 1] = max(d.start, node);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_tot[w];
      s_base += tot;
      if (s_base >= d.want) s_done = 1;
    }
    __syncthreads();
    if (s_done) break;
  }
  if (threadIdx.x == 0) {
    const uint32_t np = min(min(s_base, d.want), P.cap_pairs);
    P.ctrl->npairs = np;
    if (d.type == 1) P.ctrl->sampled_inc = np;
  }
}

// ---- FastAMM2::edge_likelihood (src/fastamm2.hh:477-520) ---------------------------------------------
template <int G, int V>
__global__ void __launch_bounds__(128) k_fa2_heldout(const Fa2Params P, uint64_t npairs, const uint32_t *pp,
                                                     const uint32_t *qq, const uint8_t *yy, double *out) {
  const unsigned mask = group_mask<G>();
  const uint32_t lane = threadIdx.x & (G - 1);
  const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  if (i >= npairs) return;
  const uint32_t p = pp[i], q = qq[i];
  const int y = yy[i];
  const Fa2Map gm = fa2_map(P, *P.ctrl);
  double2 gp[V], gq[V];
  double sp = 0.0, sq = 0.0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c0 = 2u * (lane + G * j);
    gp[j] = ld_row2(P.gamma + (size_t)p * P.ld, c0, P.ld);
    gq[j] = ld_row2(P.gamma + (size_t)q * P.ld, c0, P.ld);
    gp[j].x = c0 < P.k ? gm(gp[j].x) : 0.0; gp[j].y = c0 + 1u < P.k ? gm(gp[j].y) : 0.0;
    gq[j].x = c0 < P.k ? gm(gq[j].x) : 0.0; gq[j].y = c0 + 1u < P.k ? gm(gq[j].y) : 0.0;
    sp += gp[j].x + gp[j].y;
    sq += gq[j].x + gq[j].y;
  }
  sp = group_sum<G>(sp, mask);
  sq = group_sum<G>(sq, mask);
  double s = 0.0, sum = 0.0;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c0 = 2u * (lane + G * j);
    double2 rate = make_double2(0.0, 0.0);
    if (c0 < P.k) rate.x = P.lambda[2 * c0] / (P.lambda[2 * c0] + P.lambda[2 * c0 + 1]);
    if (c0 + 1u < P.k) rate.y = P.lambda[2 * c0 + 2] / (P.lambda[2 * c0 + 2] + P.lambda[2 * c0 + 3]);
    const double ax = (gp[j].x / sp) * (gq[j].x / sq), ay = (gp[j].y / sp) * (gq[j].y / sq);
    if (y) {
      s += ax * rate.x + ay * rate.y;
    } else {
      s += ax * (1.0 - rate.x) + ay * (1.0 - rate.y);
      sum += ax + ay;
    }
  }
  s = group_sum<G>(s, mask);
  if (!y) {
    sum = group_sum<G>(sum, mask);
    s += (1.0 - sum) * (1.0 - P.epsilon);
  }
  if (s < 1e-30) s = 1e-30;
  if (lane == 0) out[i] = log(s);
}

// ---- one pair, no side effects (svi_fa2_phi_pair) ------------------------------------------------------
template <int G, int V>
__global__ void __launch_bounds__(32) k_fa2_one_pair(const Fa2Params P, uint32_t p, uint32_t q, int y, double *phi_out,
                                                     uint32_t *rounds_out) {
  if (threadIdx.x >= G) return;
  const unsigned mask = group_mask<G>();
  const uint32_t lane = threadIdx.x;
  double2 ep[V], eq[V], ef[V], phi1[V], phi2[V];
  const Fa2Map gm = fa2_map(P, *P.ctrl);
  elogpi_row<G, V>(P, gm, p, lane, mask, ep);
  elogpi_row<G, V>(P, gm, q, lane, mask, eq);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c = 2u * (lane + G * j);
    ef[j].x = c < P.k ? P.elogbeta[(y ? 0 : P.ld) + c] : 0.0;
    ef[j].y = c + 1u < P.k ? P.elogbeta[(y ? 0 : P.ld) + c + 1] : 0.0;
  }
  const uint32_t r = fa2_pair_core<G, V>(P, mask, y, ep, eq, ef, lane, phi1, phi2);
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const uint32_t c = 2u * (lane + G * j);
    if (c < P.k) { phi_out[c] = phi1[j].x; phi_out[P.k + c] = phi2[j].x; }
    if (c + 1u < P.k) { phi_out[c + 1] = phi1[j].y; phi_out[P.k + c + 1] = phi2[j].y; }
  }
  if (lane == 0) *rounds_out = r;
}

// ---- lazy-mode plumbing: dense caller layout <-> stored rows, and the re-basing pass ------------------------
// import: dst[n*ld] = (src[n*k] - a0) (lazy: u = gamma - alpha with c = 1; eager a0 = 0), pad = 0
static __global__ void k_fa2_import(const double *src, double *dst, uint32_t n, uint32_t k, uint32_t ld, double a0) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)n * ld; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t c = i % ld;
    dst[i] = c < k ? src[(i / ld) * k + c] - a0 : 0.0;
  }
}
// export: dst[n*k] = gamma = alpha + c*u (lazy) or the stored value (eager)
static __global__ void k_fa2_export(const Fa2Params P, double *dst) {
  const Fa2Map gm = fa2_map(P, *P.ctrl);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)P.n * P.k; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = gm(P.gamma[(i / P.k) * P.ld + i % P.k]);
}
// fold: u <- c*u for every row (then the host resets c to 1): bounds the dynamic range of c*u
static __global__ void k_fa2_fold(const Fa2Params P) {
  const double cs = P.ctrl->cscale;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)P.n * P.ld; i += (size_t)gridDim.x * blockDim.x)
    P.gamma[i] *= cs;
}
static __global__ void k_fa2_reset_scale(const Fa2Params P) { P.ctrl->cscale = 1.0; }

}  // namespace svi
