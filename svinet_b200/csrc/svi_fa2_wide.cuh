// svi_fa2_wide.cuh -- the `-rnode -stratified` kernels (class FastAMM2) for K > 512: one BLOCK per work item.
//
// The register-tiled kernels of svi_fa2_kernels.cuh keep the whole coordinate ascent of a pair (Elogpi rows, two phis,
// the phis of two rounds ago) in the registers of one lane group and stop at 32 lanes x 8 double2.  Here a block of
// kWideT threads owns a pair; thread t holds the columns t, t + kWideT, ... and the K-wide state of the fixed point
// lives in five scratch rows per block in global memory (every column of them is read and written by one thread
// only).  Same arithmetic as the register tiles: max / sum-of-exp softmax, both sides of a round read the OLD phis,
// convergence tested on odd rounds against the phis of two rounds earlier (src/fastamm2.hh:151-209); the far
// endpoint's Robbins-Monro blend is done by the pair's block, start-node phi and phi1*phi2 accumulate in the block's
// own partial rows (no atomics, fixed order).  Correctness path of a range no benchmarked configuration reaches; not
// tuned.  Like svi_ls_wide.cuh the source uses only threadIdx/blockIdx, __syncthreads and block-shared arrays and is
// also run as host code by tests/cc/fa2_wide_emul.cc (against the FastAMM2 oracle, and under ThreadSanitizer).
#pragma once
#include "svi_fa2_kernels.cuh"
#include "svi_wide_reduce.cuh"

namespace svi {

constexpr uint32_t kFa2WideRows = 5;       // scratch rows per pair block: Elogpi of the far endpoint, phi1, phi2, old1, old2
constexpr uint32_t kFa2WideOneRows = 6;    // svi_fa2_phi_pair: Elogpi of both endpoints + the four above

// Elogpi row of node a from its stored row (FastAMM2::set_dir_exp(a,..), src/fastamm2.hh:424-435) -> e[0..k)
__device__ __forceinline__ void wide_elogpi_row(const Fa2Params &P, const Fa2Map &gm, uint32_t a, double *e, double *red) {
  const uint32_t t = threadIdx.x;
  const double *row = P.gamma + (size_t)a * P.ld;
  double s = 0.0;
  for (uint32_t c = t; c < P.k; c += kWideT) s += gm(row[c]);
  s = wide_sum(s, red);
  const double psi_sum = digamma_pos(s);
  for (uint32_t c = t; c < P.k; c += kWideT) e[c] = digamma_pos(gm(row[c])) - psi_sum;
}

// The coordinate ascent of one pair (fa2_pair_core of the register tiles).  e1/e2: Elogpi rows of p and q; ef: Elogf;
// on return phi1/phi2 hold the phis.  Returns the number of rounds (the same value in every thread).
__device__ __forceinline__ uint32_t fa2_pair_core_wide(const Fa2Params &P, int y, const double *e1, const double *e2,
                                                       const double *ef, double *phi1, double *phi2, double *old1,
                                                       double *old2, double *red, double *red2) {
  const uint32_t t = threadIdx.x;
  const double u0 = 1.0 / (double)P.k, inv_k = 1.0 / (double)P.k;
  for (uint32_t c = t; c < P.k; c += kWideT) {
    phi1[c] = phi2[c] = u0;
    old1[c] = old2[c] = 0.0;
  }
  uint32_t rounds = 0;
  for (uint32_t i = 0; i < P.online_iters; ++i) {
    // anext[k] = Elogpi[c][k] + Elogf[k]*b[k] + [y=1](1-b[k])*log(eps)      (src/fastamm2.hh:112-118)
    double m1 = -CUDART_INF, m2 = -CUDART_INF;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      const double f1 = phi1[c], f2 = phi2[c];
      if ((i & 1u) == 0u) {
        old1[c] = f1;
        old2[c] = f2;
      }
      const double ux = y ? (1.0 - f2) * P.logeps : 0.0, wx = y ? (1.0 - f1) * P.logeps : 0.0;
      const double t1 = e1[c] + ef[c] * f2 + ux, t2 = e2[c] + ef[c] * f1 + wx;
      phi1[c] = t1;
      phi2[c] = t2;
      m1 = fmax(m1, t1);
      m2 = fmax(m2, t2);
    }
    wide_max2(m1, m2, red, red2);
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      const double x1 = exp(phi1[c] - m1), x2 = exp(phi2[c] - m2);
      phi1[c] = x1;
      phi2[c] = x2;
      s1 += x1;
      s2 += x2;
    }
    wide_sum2(s1, s2, red, red2);
    const double inv1 = 1.0 / s1, inv2 = 1.0 / s2;
    double d1 = 0.0, d2 = 0.0;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      const double v1 = phi1[c] * inv1, v2 = phi2[c] * inv2;
      d1 += fabs(v1 - old1[c]);
      d2 += fabs(v2 - old2[c]);
      phi1[c] = v1;
      phi2[c] = v2;
    }
    rounds++;
    if ((i & 1u) == 0u) continue;
    wide_sum2(d1, d2, red, red2);
    if (d1 * inv_k < P.thresh && d2 * inv_k < P.thresh) break;   // (the same d1, d2 in every thread)
  }
  return rounds;
}

// ---- prep: expectations shared by the whole minibatch (k_fa2_prep); one block -------------------------------------
static __global__ void __launch_bounds__(SVI_WIDE_T) k_fa2_prep_wide(const Fa2Params P) {
  SVI_BLOCK_SHARED double red[kWideT];
  const uint32_t t = threadIdx.x, type = P.ctrl->type, start = P.ctrl->start;
  const Fa2Map gm = fa2_map(P, *P.ctrl);
  if (t == 0) P.ctrl->last_rounds = 0;
  for (uint32_t z = t; z < P.ld; z += kWideT) {
    double e0 = -CUDART_INF, e1 = -CUDART_INF;
    if (z < P.k) {   // FastAMM2::set_dir_exp(lambda, Elogbeta), src/fastamm2.hh:401-422 (non-positive -> alpha)
      const double l0 = P.lambda[2 * z], l1 = P.lambda[2 * z + 1];
      const double ps = digamma_pos(l0 + l1);
      e0 = digamma_pos(l0 <= 0.0 ? P.alpha : l0) - ps;
      e1 = digamma_pos(l1 <= 0.0 ? P.alpha : l1) - ps;
    }
    P.elogbeta[z] = e0;
    P.elogbeta[P.ld + z] = e1;
    P.elogf[z] = z < P.k ? (type == 0 ? e0 : e1) : 0.0;   // compute_Elogf (src/fastamm2.hh:139-149)
    if (z >= P.k) P.epi_start[z] = -CUDART_INF;
  }
  wide_elogpi_row(P, gm, start, P.epi_start, red);
}

// ---- the pairs (k_fa2_pairs): persistent blocks, block b takes the pairs b, b + gridDim.x, ... ---------------------
static __global__ void __launch_bounds__(SVI_WIDE_T) k_fa2_pairs_wide(const Fa2Params P, const uint32_t cap) {
  SVI_BLOCK_SHARED double red[kWideT];
  SVI_BLOCK_SHARED double red2[kWideT];
  const uint32_t t = threadIdx.x;
  const uint32_t type = P.ctrl->type, start = P.ctrl->start, npairs = P.ctrl->npairs;
  const double rho = P.ctrl->rho_node, scale = P.ctrl->scale, cscale = P.ctrl->cscale;
  const int y = type == 0 ? 1 : 0;
  const Fa2Map gm = fa2_map(P, *P.ctrl);
  // lazy: u_a += rho*scale*phi_a / c', c' = (1 - rho) * c the scale AFTER this iteration
  const double lazy_coef = rho * scale / ((1.0 - rho) * cscale);
  double *outS = P.partS + (size_t)blockIdx.x * cap, *outL = P.partL + (size_t)blockIdx.x * cap;
  for (uint32_t c = t; c < P.ld; c += kWideT) outS[c] = outL[c] = 0.0;
  double *ea = P.wide + (size_t)blockIdx.x * kFa2WideRows * P.ld;
  double *phi1 = ea + P.ld, *phi2 = phi1 + P.ld, *old1 = phi2 + P.ld, *old2 = old1 + P.ld;
  uint32_t my_rounds = 0;
  for (uint32_t i = blockIdx.x; i < npairs; i += gridDim.x) {
    const uint32_t p = P.pairs[2 * i], q = P.pairs[2 * i + 1];
    const bool start_is_p = p == start;
    const uint32_t a = start_is_p ? q : p;
    wide_elogpi_row(P, gm, a, ea, red);
    const double *e1 = start_is_p ? P.epi_start : ea, *e2 = start_is_p ? ea : P.epi_start;
    my_rounds += fa2_pair_core_wide(P, y, e1, e2, P.elogf, phi1, phi2, old1, old2, red, red2);
    // far endpoint: gammat[a] = its phi, touched exactly once -> blend here (src/fastamm2.cc:609-613)
    double *grow = P.gamma + (size_t)a * P.ld;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      const double pa = start_is_p ? phi2[c] : phi1[c], ps = start_is_p ? phi1[c] : phi2[c];
      const double g = grow[c];
      grow[c] = P.lazy ? fma(lazy_coef, pa, g) : (1.0 - rho) * g + rho * (P.alpha + scale * pa);
      outS[c] += ps;
      outL[c] += phi1[c] * phi2[c];   // lambdat[k][t] += phi1*phi2 (src/fastamm2.cc:994-996)
    }
    if (t == 0 && !P.lazy) P.touched[a] = 1;
  }
  if (t == 0 && my_rounds) atomicAdd((unsigned long long *)&P.ctrl->last_rounds, (unsigned long long)my_rounds);
}

// ---- blend of all the rows the pair kernel did not touch (k_fa2_blend; EAGER mode only) ----------------------------
static __global__ void __launch_bounds__(SVI_WIDE_T) k_fa2_blend_wide(const Fa2Params P, const uint32_t cap) {
  const uint32_t t = threadIdx.x;
  const uint32_t start = P.ctrl->start, npairs = P.ctrl->npairs;
  const double rho = P.ctrl->rho_node, scale = P.ctrl->scale;
  const uint32_t active = min(P.pair_blocks, npairs);   // pair blocks that had pairs: the others' partials are zero
  for (uint32_t row = blockIdx.x; row < P.n; row += gridDim.x) {
    const bool touched = P.touched[row] != 0;
    __syncthreads();   // everybody has read the flag before it is cleared
    if (touched) {     // blended by the pair kernel: just clear the flag
      if (t == 0) P.touched[row] = 0;
      continue;
    }
    const bool is_start = row == start;
    double *grow = P.gamma + (size_t)row * P.ld;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      double s = 0.0;
      if (is_start)
        for (uint32_t b = 0; b < active; ++b) s += P.partS[(size_t)b * cap + c];
      const double target = is_start ? P.alpha + scale * s : P.alpha;
      grow[c] = (1.0 - rho) * grow[c] + rho * target;
    }
  }
}

// ---- FastAMM2::edge_likelihood (src/fastamm2.hh:477-520), k_fa2_heldout; one block per pair, grid-stride -----------
static __global__ void __launch_bounds__(SVI_WIDE_T) k_fa2_heldout_wide(const Fa2Params P, uint64_t npairs, const uint32_t *pp,
                                                                       const uint32_t *qq, const uint8_t *yy, double *out) {
  SVI_BLOCK_SHARED double red[kWideT];
  SVI_BLOCK_SHARED double red2[kWideT];
  const uint32_t t = threadIdx.x;
  const Fa2Map gm = fa2_map(P, *P.ctrl);
  for (uint64_t i = blockIdx.x; i < npairs; i += gridDim.x) {
    const double *rp = P.gamma + (size_t)pp[i] * P.ld, *rq = P.gamma + (size_t)qq[i] * P.ld;
    const int y = yy[i];
    double sp = 0.0, sq = 0.0;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      sp += gm(rp[c]);
      sq += gm(rq[c]);
    }
    wide_sum2(sp, sq, red, red2);
    double s = 0.0, sum = 0.0;
    for (uint32_t c = t; c < P.k; c += kWideT) {
      const double rate = P.lambda[2 * c] / (P.lambda[2 * c] + P.lambda[2 * c + 1]);
      const double ax = (gm(rp[c]) / sp) * (gm(rq[c]) / sq);
      if (y) {
        s += ax * rate;
      } else {
        s += ax * (1.0 - rate);
        sum += ax;
      }
    }
    wide_sum2(s, sum, red, red2);
    if (!y) s += (1.0 - sum) * (1.0 - P.epsilon);
    if (s < 1e-30) s = 1e-30;
    if (t == 0) out[i] = log(s);
  }
}

// ---- one pair, no side effects (svi_fa2_phi_pair), k_fa2_one_pair; one block; rows = kFa2WideOneRows scratch rows ---
static __global__ void __launch_bounds__(SVI_WIDE_T) k_fa2_one_pair_wide(const Fa2Params P, double *rows, uint32_t p, uint32_t q,
                                                                        int y, double *phi_out, uint32_t *rounds_out) {
  SVI_BLOCK_SHARED double red[kWideT];
  SVI_BLOCK_SHARED double red2[kWideT];
  const uint32_t t = threadIdx.x;
  const Fa2Map gm = fa2_map(P, *P.ctrl);
  double *ep = rows, *eq = ep + P.ld, *phi1 = eq + P.ld, *phi2 = phi1 + P.ld, *old1 = phi2 + P.ld, *old2 = old1 + P.ld;
  wide_elogpi_row(P, gm, p, ep, red);
  wide_elogpi_row(P, gm, q, eq, red);
  const uint32_t r = fa2_pair_core_wide(P, y, ep, eq, P.elogbeta + (y ? 0 : P.ld), phi1, phi2, old1, old2, red, red2);
  for (uint32_t c = t; c < P.k; c += kWideT) {
    phi_out[c] = phi1[c];
    phi_out[P.k + c] = phi2[c];
  }
  if (t == 0) *rounds_out = r;
}

}  // namespace svi
