/* oracle_ls_omp.c -- the oracle's sweep (oracle_ls.c: orc_step) spread over the host cores with OpenMP.
 *
 * TEST INFRASTRUCTURE ONLY: the "all host cores" leg of the CPU baseline and nothing
 * else; the reference path itself is serial (src/linksampling.cc:557-790).  Same arithmetic per link as orc_step --
 * running log-sum-exp (:685-694), exp(phi - r) (src/matrix.hh:320-325), one-hot shortcut (:619-631), mean indicators
 * (:526-545), s3 with the Q4 off-by-one (:731-746), set_dir_exp (src/linksampling.hh:171-187), prune (:456-491) -- but
 * the scatter into the two gammanext rows uses atomic adds and the K-vectors are reduced from per-thread copies, so
 * the summation ORDER differs from run to run: results agree with orc_step to rounding (tests/test_oracle_omp.py,
 * 1e-9), not bit for bit.  The iter > 1000 active-set branch is not restated here (orc_step covers it).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_ls.h"

int orc_omp_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_step_omp(orc_state *s, int annealing, int write_comm, int threads) {
  const uint32_t n = s->n, k = s->k;
  const size_t nk = (size_t)n * k;
  double *gnext = s->gammanext, *lnext = s->lambdanext;
  if (threads < 1) threads = orc_omp_max_threads();
  if (write_comm) memset(s->member, 0, nk);
  memset(s->s1, 0, k * sizeof(double));
  memset(s->s2, 0, k * sizeof(double));
  memset(s->s3, 0, k * sizeof(double));
  memset(s->sum, 0, k * sizeof(double));
  uint64_t dense = 0, shortcut = 0;

#pragma omp parallel num_threads(threads) reduction(+ : dense, shortcut)
  {
    double *phi = (double *)calloc(k ? k : 1, sizeof(double));
    double *tsum = (double *)calloc(k ? k : 1, sizeof(double));
    /* ---- phi sweep, :605-725 ---- */
#pragma omp for schedule(static)
    for (uint64_t e = 0; e < s->nlinks; ++e) {
      const uint32_t p = s->links[2 * e], q = s->links[2 * e + 1];
      const uint32_t pc = s->converged[p], qc = s->converged[q];
      double *gp = gnext + (size_t)p * k, *gq = gnext + (size_t)q * k;
      const double *ep = s->Elogpi + (size_t)p * k, *eq = s->Elogpi + (size_t)q * k;
      if ((pc != 0) != (qc != 0)) {                       /* :622-631 */
        const uint32_t c = (pc ? pc : qc) - 1;
#pragma omp atomic
        gp[c] += 1;
#pragma omp atomic
        gq[c] += 1;
        tsum[c] += 2;
        shortcut++;
        continue;
      }
      double r = .0;
      for (uint32_t c = 0; c < k; ++c) {                  /* :685-694 */
        phi[c] = ep[c] + eq[c] + s->Elogbeta[2 * c];
        if (c == 0) r = phi[c];
        else if (phi[c] < r) r = r + log(1 + exp(phi[c] - r));
        else r = phi[c] + log(1 + exp(r - phi[c]));
      }
      uint32_t max_k = 65535;
      double maxv = .0;
      for (uint32_t c = 0; c < k; ++c) {
        phi[c] = exp(phi[c] - r);
#pragma omp atomic
        gp[c] += phi[c];
#pragma omp atomic
        gq[c] += phi[c];
        tsum[c] += 2 * phi[c];
        if (phi[c] > maxv) { maxv = phi[c]; max_k = c; }
      }
      dense++;
      if (write_comm && maxv > 0.0) {                     /* :704-717; byte stores of the same value: benign */
        s->member[(size_t)p * k + max_k] = 1;
        s->member[(size_t)q * k + max_k] = 1;
      }
    }
#pragma omp critical
    for (uint32_t c = 0; c < k; ++c) s->sum[c] += tsum[c];
#pragma omp barrier
#pragma omp single
    for (uint32_t c = 0; c < k; ++c) lnext[2 * c] += s->sum[c];

    /* ---- compute_mean_indicators, :526-545 ---- */
    memset(tsum, 0, k * sizeof(double));
    double *ts2 = phi;
    memset(ts2, 0, k * sizeof(double));
#pragma omp for schedule(static)
    for (uint32_t p = 0; p < n; ++p) {
      if (s->tl[p] == 0) continue;
      for (uint32_t c = 0; c < k; ++c) {
        double *m = s->mphi + (size_t)p * k + c, *g = gnext + (size_t)p * k + c;
        *m = (*g - s->alpha) / s->tl[p];
        tsum[c] += *m;
        ts2[c] += *m * *m;
        *g += (n - s->tl[p] - 1) * *m;
        if (annealing) *g *= s->ones / s->sum[c];
      }
    }
#pragma omp critical
    for (uint32_t c = 0; c < k; ++c) { s->s1[c] += tsum[c]; s->s2[c] += ts2[c]; }
#pragma omp barrier

    /* ---- s3 sweep, :731-746 ---- */
    memset(tsum, 0, k * sizeof(double));
#pragma omp for schedule(static)
    for (uint64_t e = 0; e < s->nlinks; ++e) {
      const uint32_t p = s->links[2 * e], q = s->links[2 * e + 1];
      const uint32_t pc = s->converged[p], qc = s->converged[q];
      const double *mp = s->mphi + (size_t)p * k, *mq = s->mphi + (size_t)q * k;
      if (pc && !qc) tsum[pc - 1] += pc < k ? mq[pc] : 0.0;
      else if (!pc && qc) tsum[qc - 1] += qc < k ? mp[qc] : 0.0;
      else for (uint32_t c = 0; c < k; ++c) tsum[c] += mp[c] * mq[c];
    }
#pragma omp critical
    for (uint32_t c = 0; c < k; ++c) s->s3[c] += tsum[c];
    free(phi); free(tsum);
  }
  s->cnt_dense = dense; s->cnt_shortcut = shortcut; s->cnt_sparse = 0;

  for (uint32_t c = 0; c < k; ++c) lnext[2 * c + 1] += s->s1[c] * s->s1[c] - s->s2[c] - s->s3[c];   /* :748-749 */
  { double *t = s->gamma; s->gamma = s->gammanext; s->gammanext = t; }                            /* :751-755 */
  { double *t = s->lambda; s->lambda = s->lambdanext; s->lambdanext = t; }
  for (uint32_t c = 0; c < k; ++c) { s->lambdanext[2 * c] = s->eta0; s->lambdanext[2 * c + 1] = s->eta1; }
  const uint32_t lim = k / 10, stride = lim ? lim : 1;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (uint32_t p = 0; p < n; ++p) {
    double *g = s->gamma + (size_t)p * k, *e = s->Elogpi + (size_t)p * k, *gn = s->gammanext + (size_t)p * k;
    double sum = .0;
    for (uint32_t c = 0; c < k; ++c) { sum += g[c]; gn[c] = s->alpha; }
    const double psi_sum = orc_digamma(sum);                                                      /* :757 */
    uint32_t active = 0, pk = 0, len = 0;
    for (uint32_t c = 0; c < k; ++c) {
      e[c] = orc_digamma(g[c]) - psi_sum;
      if (g[c] - s->alpha >= 1) {                                                                /* prune, :456-491 */
        active++;
        if (active <= lim) s->active_k[(size_t)p * stride + len++] = (uint16_t)c;
        pk = c;
      }
    }
    if (active > lim) len = 0;
    if (active == 1) s->converged[p] = pk + 1;
    s->active_comms[p] = active;
    s->active_len[p] = len;
  }
  orc_set_dir_exp(s->lambda, s->Elogbeta, k, 2);                                                  /* :758-759 */
}
