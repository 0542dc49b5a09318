/* oracle_fa2.c -- see oracle_fa2.h.  TEST INFRASTRUCTURE ONLY. */
#include "oracle_fa2.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_fa2 {
  const orc_graph *g;
  orc_fa2_options o;
  orc_rng rng;
  uint32_t n, k;
  uint32_t m_sets;                 /* _m = 10 (fastamm2.cc:11) */
  double alpha, logepsilon, inf_epsilon, link_thresh;
  double tau0, nodetau0;           /* env value + 1 (fastamm2.cc:19-20) */
  uint32_t iter, lambda_start_iter;
  double zeros_prob, ones_prob;
  uint64_t total_pairs_sampled;
  double *gamma, *gammat, *Elogpi; /* [n*k] */
  double *lambda, *lambdat, *Elogbeta, *eta;   /* [k*2] */
  double *nodec;                   /* [n] */
  uint32_t *shuffled;              /* [n] */
  uint8_t *touched;                /* NodeMap of the current iteration */
  /* held-out set: draw order + std::map<Edge,bool> order */
  uint32_t *ho, *ho_sorted;
  uint64_t nho;
  /* planned minibatch */
  uint32_t plan_type, plan_start;
  uint32_t *plan_pairs;
  uint64_t plan_npairs, plan_cap, plan_sampled_inc;
  double prev_h, max_h;
  uint32_t nh;
  int stopped;
  char *hlog;
  size_t hlog_len, hlog_cap;
};

/* ---- GSL boundary, restated (see header) ---------------------------------------------------- */
static double rng_uniform_pos(orc_rng *r) {
  double x;
  do { x = orc_rng_uniform(r); } while (x == 0);
  return x;
}
static double ran_gaussian_polar(orc_rng *r) {
  double x, y, r2;
  do {
    x = -1 + 2 * rng_uniform_pos(r);
    y = -1 + 2 * rng_uniform_pos(r);
    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0);
  return y * sqrt(-2.0 * log(r2) / r2);
}
static double ran_gamma(orc_rng *r, double a, double b) {     /* Marsaglia & Tsang 2000 */
  if (a < 1) {
    double u = rng_uniform_pos(r);
    return ran_gamma(r, 1.0 + a, b) * pow(u, 1.0 / a);
  }
  double d = a - 1.0 / 3.0, c = (1.0 / 3.0) / sqrt(d), x, v, u;
  for (;;) {
    do { x = ran_gaussian_polar(r); v = 1.0 + c * x; } while (v <= 0);
    v = v * v * v;
    u = rng_uniform_pos(r);
    if (u < 1 - 0.0331 * x * x * x * x) break;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
  }
  return b * d * v;
}

void orc_fa2_options_default(orc_fa2_options *o, uint32_t k) {
  memset(o, 0, sizeof *o);
  o->k = k; o->heldout_ratio = 0.01; o->reportfreq = 100; o->eta0 = 1; o->eta1 = 1; o->epsilon = 1e-30;
  o->tau0 = 1024; o->kappa = 0.9; o->nodetau0 = 1024; o->nodekappa = 0.5;
  o->online_iterations = 50; o->meanchangethresh = 0.00001;
}

/* ---- held-out set ------------------------------------------------------------------------- */
static int ho_contains(const orc_fa2 *m, uint32_t a, uint32_t b) {
  uint64_t lo = 0, hi = m->nho;
  while (lo < hi) {
    uint64_t mid = (lo + hi) / 2;
    uint32_t x = m->ho_sorted[2 * mid], y = m->ho_sorted[2 * mid + 1];
    if (x == a && y == b) return 1;
    if (x < a || (x == a && y < b)) lo = mid + 1; else hi = mid;
  }
  return 0;
}
static void ho_insert(orc_fa2 *m, uint32_t a, uint32_t b) {
  m->ho = (uint32_t *)realloc(m->ho, (m->nho + 1) * 2 * sizeof(uint32_t));
  m->ho_sorted = (uint32_t *)realloc(m->ho_sorted, (m->nho + 1) * 2 * sizeof(uint32_t));
  m->ho[2 * m->nho] = a; m->ho[2 * m->nho + 1] = b;
  uint64_t pos = m->nho;
  while (pos > 0 && (m->ho_sorted[2 * (pos - 1)] > a ||
                     (m->ho_sorted[2 * (pos - 1)] == a && m->ho_sorted[2 * (pos - 1) + 1] > b))) {
    m->ho_sorted[2 * pos] = m->ho_sorted[2 * (pos - 1)];
    m->ho_sorted[2 * pos + 1] = m->ho_sorted[2 * (pos - 1) + 1];
    pos--;
  }
  m->ho_sorted[2 * pos] = a; m->ho_sorted[2 * pos + 1] = b;
  m->nho++;
}

/* FastAMM2::edge_ok, fastamm2.hh:524-541 (single_heldout_set = true: only the held-out map) */
static int edge_ok(const orc_fa2 *m, uint32_t a, uint32_t b) {
  if (a == b) return 0;
  return !ho_contains(m, a, b);
}

/* FastAMM2::get_random_edge, fastamm2.hh:543-565 */
static void get_random_edge(orc_fa2 *m, int link, uint32_t *a, uint32_t *b) {
  if (!link) {
    do {
      uint32_t f = (uint32_t)orc_rng_uniform_int(&m->rng, m->n);
      uint32_t s = (uint32_t)orc_rng_uniform_int(&m->rng, m->n);
      *a = f < s ? f : s; *b = f < s ? s : f;
    } while (!edge_ok(m, *a, *b));
  } else {
    do {
      uint32_t i = (uint32_t)orc_rng_uniform_int(&m->rng, m->g->ones);
      *a = m->g->edges[2 * i]; *b = m->g->edges[2 * i + 1];
    } while (!edge_ok(m, *a, *b));
  }
}

/* FastAMM2::set_heldout_sample, fastamm2.cc:393-421 */
static void set_heldout_sample(orc_fa2 *m, int s) {
  int c0 = 0, c1 = 0, p = s / 2;
  while (c0 < p || c1 < p) {
    uint32_t a, b;
    if (c0 == p) get_random_edge(m, 1, &a, &b); else get_random_edge(m, 0, &a, &b);
    int y = orc_graph_y(m->g, a, b);
    if (y == 0 && c0 < p) { c0++; ho_insert(m, a, b); }
    if (y == 1 && c1 < p) { c1++; ho_insert(m, a, b); }
  }
}

/* FastAMM2::set_dir_exp(a, u, exp), fastamm2.hh:424-435 */
static void set_dir_exp_row(const double *u, double *e, uint32_t cols) {
  double s = .0;
  for (uint32_t j = 0; j < cols; ++j) s += u[j];
  double psi_sum = orc_digamma(s);
  for (uint32_t j = 0; j < cols; ++j) e[j] = orc_digamma(u[j]) - psi_sum;
}
/* FastAMM2::set_dir_exp(u, exp), fastamm2.hh:401-422 (non-positive entries read as alpha) */
static void set_dir_exp_all(const orc_fa2 *m, const double *u, double *e, uint32_t rows, uint32_t cols) {
  for (uint32_t i = 0; i < rows; ++i) {
    double s = .0;
    for (uint32_t j = 0; j < cols; ++j) s += u[(size_t)i * cols + j];
    double psi_sum = orc_digamma(s);
    for (uint32_t j = 0; j < cols; ++j) {
      double tt = u[(size_t)i * cols + j];
      if (tt <= .0) tt = m->alpha;
      e[(size_t)i * cols + j] = orc_digamma(tt) - psi_sum;
    }
  }
}

/* D1Array::logsum, matrix.hh:296-310 */
static double logsum(const double *d, uint32_t n) {
  if (n == 1) return d[0];
  double r = d[0];
  for (uint32_t i = 1; i < n; ++i)
    if (d[i] < r) r = r + log(1 + exp(d[i] - r));
    else r = d[i] + log(1 + exp(r - d[i]));
  return r;
}

/* PhiCompute::update_phis (fastamm2.hh:105-137) for one side: anext from the OTHER side's phi `b` */
static void update_phis(uint32_t k, const double *elogpi_c, const double *elogf, const double *b, int y,
                        double logepsilon, double *anext) {
  for (uint32_t i = 0; i < k; ++i) {
    double u = .0;
    if (y == 1) u = (1 - b[i]) * logepsilon;
    anext[i] = elogpi_c[i] + (elogf[i] * b[i]) + u;
  }
  double s = logsum(anext, k);
  for (uint32_t i = 0; i < k; ++i) anext[i] = exp(anext[i] - s);     /* lognormalize, matrix.hh:312-318 */
}

/* PhiCompute::update_phis_until_conv, fastamm2.hh:151-209 (phifix = false) */
uint32_t orc_fa2_phi_pair(uint32_t k, const double *elogpi_p, const double *elogpi_q, const double *elogf,
                          int y, double logepsilon, uint32_t online_iterations, double thresh,
                          double *phi1, double *phi2) {
  double *buf = (double *)calloc((size_t)6 * k, sizeof(double));
  double *next1 = buf, *next2 = buf + k, *old1 = buf + 2 * k, *old2 = buf + 3 * k, *v1 = buf + 4 * k, *v2 = buf + 5 * k;
  double u = 1. / k;
  for (uint32_t i = 0; i < k; ++i) phi1[i] = phi2[i] = u;
  uint32_t rounds = 0;
  for (uint32_t i = 0; i < online_iterations; ++i) {
    if (i % 2 == 0) {
      memcpy(old1, phi1, k * sizeof(double));
      memcpy(old2, phi2, k * sizeof(double));
    }
    update_phis(k, elogpi_p, elogf, phi2, y, logepsilon, next1);     /* both sides read the OLD phis */
    update_phis(k, elogpi_q, elogf, phi1, y, logepsilon, next2);
    for (uint32_t c = 0; c < k; ++c) { v1[c] = next1[c] - old1[c]; v2[c] = next2[c] - old2[c]; }
    memcpy(phi1, next1, k * sizeof(double));
    memcpy(phi2, next2, k * sizeof(double));
    rounds++;
    if (i % 2 == 0) continue;
    double s1 = .0, s2 = .0;
    for (uint32_t c = 0; c < k; ++c) s1 += fabs(v1[c]);
    for (uint32_t c = 0; c < k; ++c) s2 += fabs(v2[c]);
    if (s1 / k < thresh && s2 / k < thresh) break;
  }
  free(buf);
  return rounds;
}

/* ---- likelihood ---------------------------------------------------------------------------- */
double orc_fa2_edge_likelihood(const orc_fa2 *m, uint32_t p, uint32_t q, int y) {   /* fastamm2.hh:477-520 */
  const uint32_t k = m->k;
  const double *gp = m->gamma + (size_t)p * k, *gq = m->gamma + (size_t)q * k;
  double sp = .0, sq = .0;
  for (uint32_t i = 0; i < k; ++i) sp += gp[i];
  for (uint32_t i = 0; i < k; ++i) sq += gq[i];
  double v = 1 - m->o.epsilon;
  double s = .0;
  if (y == 1) {
    for (uint32_t z = 0; z < k; ++z) {
      double rate = m->lambda[2 * z] / (m->lambda[2 * z] + m->lambda[2 * z + 1]);
      s += (gp[z] / sp) * (gq[z] / sq) * rate;
    }
  } else {
    double sum = .0;
    for (uint32_t z = 0; z < k; ++z) {
      double rate = m->lambda[2 * z] / (m->lambda[2 * z] + m->lambda[2 * z + 1]);
      s += (gp[z] / sp) * (gq[z] / sq) * (1 - rate);
      sum += (gp[z] / sp) * (gq[z] / sq);
    }
    s += (1 - sum) * v;
  }
  if (s < 1e-30) s = 1e-30;
  return log(s);
}

static void hlog_append(orc_fa2 *m, const char *line) {
  size_t l = strlen(line);
  if (m->hlog_len + l + 1 > m->hlog_cap) {
    m->hlog_cap = (m->hlog_cap + l + 1) * 2;
    m->hlog = (char *)realloc(m->hlog, m->hlog_cap);
  }
  memcpy(m->hlog + m->hlog_len, line, l + 1);
  m->hlog_len += l;
}

/* FastAMM2::heldout_likelihood, fastamm2.cc:1297-1392 (the stop branch cannot fire while
 * _zeros_prob = _ones_prob = 0: nshol is -0 every time) */
static void heldout_likelihood(orc_fa2 *m) {
  uint32_t k = 0, kzeros = 0, kones = 0;
  double s = .0, szeros = 0, sones = 0;
  for (uint64_t i = 0; i < m->nho; ++i) {
    uint32_t p = m->ho_sorted[2 * i], q = m->ho_sorted[2 * i + 1];
    int y = orc_graph_y(m->g, p, q);
    double u = orc_fa2_edge_likelihood(m, p, q, y);
    s += u; k += 1;
    if (y) { sones += u; kones++; } else { szeros += u; kzeros++; }
  }
  double nshol = (m->zeros_prob * (szeros / kzeros)) + (m->ones_prob * (sones / kones));
  char line[640];
  snprintf(line, sizeof line, "%d\t%d\t%.9f\t%d\t%.9f\t%d\t%.9f\t%d\t%.9f\t%.9f\t%.9f\t%llu\n",
           m->iter, 0, s / k, k, szeros / kzeros, kzeros, sones / kones, kones,
           m->zeros_prob * (szeros / kzeros), m->ones_prob * (sones / kones), nshol,
           (unsigned long long)m->total_pairs_sampled);
  hlog_append(m, line);
  double a = nshol;
  if (m->iter > m->n || m->iter > 5000) {
    if (a > m->prev_h && m->prev_h != 0 && fabs((a - m->prev_h) / m->prev_h) < 0.00001) { /* stop: unreachable */ }
    else if (a < m->prev_h) m->nh++;
    else if (a > m->prev_h) m->nh = 0;
    if (a > m->max_h) m->max_h = a;
  }
  m->prev_h = a;
}

/* ---- constructor, fastamm2.cc:8-248 --------------------------------------------------------- */
orc_fa2 *orc_fa2_create(const orc_graph *g, const orc_fa2_options *o) {
  orc_fa2 *m = (orc_fa2 *)calloc(1, sizeof *m);
  m->g = g; m->o = *o;
  const uint32_t n = g->n, k = o->k;
  m->n = n; m->k = k; m->m_sets = 10;
  m->alpha = (double)1 / k;                       /* env.hh:344 */
  m->logepsilon = log(o->epsilon);                /* env.hh:396 */
  m->inf_epsilon = 0.5; m->link_thresh = 0.9;     /* fastamm2.cc:15-16 */
  m->tau0 = o->tau0 + 1; m->nodetau0 = o->nodetau0 + 1;
  m->prev_h = -2147483647; m->max_h = -2147483647;
  const size_t nk = (size_t)(n ? n : 1) * k;
  m->gamma = (double *)calloc(nk, sizeof(double));
  m->gammat = (double *)calloc(nk, sizeof(double));
  m->Elogpi = (double *)calloc(nk, sizeof(double));
  m->lambda = (double *)calloc(2 * (size_t)k, sizeof(double));
  m->lambdat = (double *)calloc(2 * (size_t)k, sizeof(double));
  m->Elogbeta = (double *)calloc(2 * (size_t)k, sizeof(double));
  m->eta = (double *)calloc(2 * (size_t)k, sizeof(double));
  m->nodec = (double *)calloc(n ? n : 1, sizeof(double));
  m->shuffled = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  m->touched = (uint8_t *)calloc(n ? n : 1, 1);
  for (uint32_t i = 0; i < k; ++i) { m->eta[2 * i] = o->eta0; m->eta[2 * i + 1] = o->eta1; }

  orc_rng_seed(&m->rng, 0);
  if (o->seed) orc_rng_seed(&m->rng, (unsigned long)o->seed);          /* :89-90 */
  /* shuffle_nodes, :489-495; gsl_ran_shuffle = Fisher-Yates from the top */
  for (uint32_t i = 0; i < n; ++i) m->shuffled[i] = i;
  for (uint32_t i = n - 1; n && i > 0; i--) {
    uint32_t j = (uint32_t)orc_rng_uniform_int(&m->rng, (unsigned long)i + 1);
    uint32_t t = m->shuffled[i]; m->shuffled[i] = m->shuffled[j]; m->shuffled[j] = t;
  }
  int s = (int)(o->heldout_ratio * g->ones);                            /* init_heldout, :304 */
  set_heldout_sample(m, s);
  /* init_gamma, :497-515 */
  for (uint32_t i = 0; i < n; ++i)
    for (uint32_t j = 0; j < k; ++j) {
      double *d = m->gamma + (size_t)i * k + j;
      if (o->deterministic) {
        *d = 0.09 + (0.01 * ((i + 1) / (i + j + 1)));                   /* integer division, sic */
        if (*d > 1.) *d = 0.9;
      } else {
        double v = (k < 100) ? 1.0 : (double)100.0 / k;
        *d = ran_gamma(&m->rng, 100 * v, 0.01);
      }
    }
  /* init_lambda, :518-531 */
  for (uint32_t c = 0; c < k; ++c)
    for (uint32_t t = 0; t < 2; ++t) {
      double v = (k <= 100) ? 1.0 : (double)100.0 / k;
      m->lambda[2 * c + t] = m->eta[2 * c + t] + ran_gamma(&m->rng, 100 * v, 0.01);
    }
  set_dir_exp_all(m, m->gamma, m->Elogpi, n, k);                        /* :146-147 */
  set_dir_exp_all(m, m->lambda, m->Elogbeta, k, 2);
  m->iter = 0;
  heldout_likelihood(m);                                                /* :230 */
  return m;
}

void orc_fa2_free(orc_fa2 *m) {
  if (!m) return;
  free(m->gamma); free(m->gammat); free(m->Elogpi); free(m->lambda); free(m->lambdat); free(m->Elogbeta);
  free(m->eta); free(m->nodec); free(m->shuffled); free(m->touched); free(m->ho); free(m->ho_sorted);
  free(m->plan_pairs); free(m->hlog); free(m);
}

static void plan_push(orc_fa2 *m, uint32_t a, uint32_t b) {
  if (m->plan_npairs == m->plan_cap) {
    m->plan_cap = m->plan_cap ? m->plan_cap * 2 : 64;
    m->plan_pairs = (uint32_t *)realloc(m->plan_pairs, m->plan_cap * 2 * sizeof(uint32_t));
  }
  m->plan_pairs[2 * m->plan_npairs] = a < b ? a : b;      /* Network::order_edge */
  m->plan_pairs[2 * m->plan_npairs + 1] = a < b ? b : a;
  m->plan_npairs++;
}

void orc_fa2_plan(orc_fa2 *m) {
  const orc_graph *g = m->g;
  m->plan_npairs = 0;
  m->plan_type = orc_rng_uniform(&m->rng) < m->inf_epsilon ? 1u : 0u;     /* gsl_ran_bernoulli, :574 */
  m->plan_start = (uint32_t)orc_rng_uniform_int(&m->rng, m->n);           /* :936 / :1078 */
  const uint32_t start = m->plan_start;
  if (m->plan_type == 0) {
    /* opt_process, :943-960: the start node's links in adjacency order, held-out ones skipped */
    m->plan_sampled_inc = g->adj_off[start + 1] - g->adj_off[start];
    for (uint64_t r = g->adj_off[start]; r < g->adj_off[start + 1]; ++r) {
      uint32_t a = g->adj[r];
      uint32_t x = start < a ? start : a, y = start < a ? a : start;
      if (!edge_ok(m, x, y)) continue;
      plan_push(m, start, a);
    }
  } else {
    /* opt_process_noninf, :1095-1125 */
    uint32_t setsize = (uint32_t)((double)m->n / (double)m->m_sets);
    double v = (double)(orc_rng_uniform_int(&m->rng, m->n)) / setsize;
    uint32_t q = ((int)v) * setsize;
    while (m->plan_npairs < setsize) {
      uint32_t node = m->shuffled[q];
      if (node == start) { q = (q + 1) % m->n; continue; }
      int y = orc_graph_y(g, start, node);
      uint32_t x = start < node ? start : node, z = start < node ? node : start;
      if (y == 0 && edge_ok(m, x, z)) plan_push(m, start, node);
      q = (q + 1) % m->n;
    }
    m->plan_sampled_inc = m->plan_npairs;
  }
}

void orc_fa2_process(orc_fa2 *m) {
  const uint32_t n = m->n, k = m->k;
  const int y = m->plan_type == 0 ? 1 : 0;
  memset(m->lambdat, 0, 2 * (size_t)k * sizeof(double));                  /* :566 */
  set_dir_exp_all(m, m->lambda, m->Elogbeta, k, 2);                       /* :567 */
  memset(m->touched, 0, n);
  const uint32_t start = m->plan_start;
  set_dir_exp_row(m->gamma + (size_t)start * k, m->Elogpi + (size_t)start * k, k);
  memset(m->gammat + (size_t)start * k, 0, k * sizeof(double));
  m->touched[start] = 1;
  m->total_pairs_sampled += m->plan_sampled_inc;
  double *phi1 = (double *)calloc((size_t)3 * k, sizeof(double)), *phi2 = phi1 + k, *elogf = phi1 + 2 * k;
  /* compute_Elogf, fastamm2.hh:139-149 */
  for (uint32_t c = 0; c < k; ++c) {
    elogf[c] = .0;
    for (uint32_t t = 0; t < 2; ++t) elogf[c] += m->Elogbeta[2 * c + t] * (t == 0 ? y : (1 - y));
  }
  for (uint64_t i = 0; i < m->plan_npairs; ++i) {
    uint32_t p = m->plan_pairs[2 * i], q = m->plan_pairs[2 * i + 1];
    uint32_t a = p != start ? p : q;
    m->touched[a] = 1;
    set_dir_exp_row(m->gamma + (size_t)a * k, m->Elogpi + (size_t)a * k, k);
    memset(m->gammat + (size_t)a * k, 0, k * sizeof(double));
    orc_fa2_phi_pair(k, m->Elogpi + (size_t)p * k, m->Elogpi + (size_t)q * k, elogf, y, m->logepsilon,
                     m->o.online_iterations, m->o.meanchangethresh, phi1, phi2);
    for (uint32_t c = 0; c < k; ++c) m->gammat[(size_t)p * k + c] += phi1[c];
    for (uint32_t c = 0; c < k; ++c) m->gammat[(size_t)q * k + c] += phi2[c];
    for (uint32_t c = 0; c < k; ++c)
      for (uint32_t t = 0; t < 2; ++t) m->lambdat[2 * c + t] += phi1[c] * phi2[c] * (t == 0 ? y : (1 - y));
  }
  free(phi1);
  /* Robbins-Monro blends, :586-638 */
  double scale = (m->plan_type == 0) ? (double)n / (2 * (1 - m->inf_epsilon))
                                     : ((double)n * (double)m->m_sets) / (2 * m->inf_epsilon);
  for (uint32_t i = 0; i < n; ++i) {
    double rho = pow(m->nodetau0 + m->nodec[i], -1 * m->o.nodekappa);
    double *gd = m->gamma + (size_t)i * k, *gdt = m->gammat + (size_t)i * k;
    if (m->touched[i]) {
      for (uint32_t c = 0; c < k; ++c) gd[c] = (1 - rho) * gd[c] + rho * (m->alpha + scale * gdt[c]);
    } else {
      for (uint32_t c = 0; c < k; ++c) gd[c] = (1 - rho) * gd[c] + rho * m->alpha;
    }
    m->nodec[i]++;
  }
  if (!m->o.nolambda) {
    double rhot = pow(m->tau0 + (m->iter - m->lambda_start_iter + 1), -1 * m->o.kappa);
    for (uint32_t c = 0; c < k; ++c)
      for (uint32_t t = 0; t < 2; ++t) {
        m->lambdat[2 * c + t] = m->eta[2 * c + t] + scale * m->lambdat[2 * c + t];
        m->lambda[2 * c + t] = (1 - rhot) * m->lambda[2 * c + t] + rhot * m->lambdat[2 * c + t];
      }
  }
  m->iter++;
}

uint32_t orc_fa2_run(orc_fa2 *m, uint32_t max_steps) {
  uint32_t steps = 0;
  while (!m->stopped) {
    if (m->o.max_iterations && m->iter > m->o.max_iterations) { m->stopped = 1; break; }   /* :546 */
    if (max_steps && steps >= max_steps) break;
    orc_fa2_plan(m);
    orc_fa2_process(m);
    steps++;
    if (m->iter % m->o.reportfreq == 0) heldout_likelihood(m);                             /* :651-669 */
  }
  return steps;
}

uint32_t orc_fa2_n(const orc_fa2 *m) { return m->n; }
uint32_t orc_fa2_k(const orc_fa2 *m) { return m->k; }
uint32_t orc_fa2_iter(const orc_fa2 *m) { return m->iter; }
int orc_fa2_stopped(const orc_fa2 *m) { return m->stopped; }
double *orc_fa2_gamma(orc_fa2 *m) { return m->gamma; }
double *orc_fa2_lambda(orc_fa2 *m) { return m->lambda; }
double orc_fa2_alpha(const orc_fa2 *m) { return m->alpha; }
const uint32_t *orc_fa2_shuffled(const orc_fa2 *m) { return m->shuffled; }
uint32_t orc_fa2_plan_type(const orc_fa2 *m) { return m->plan_type; }
uint32_t orc_fa2_plan_start(const orc_fa2 *m) { return m->plan_start; }
uint64_t orc_fa2_plan_npairs(const orc_fa2 *m) { return m->plan_npairs; }
const uint32_t *orc_fa2_plan_pairs(const orc_fa2 *m) { return m->plan_pairs; }
uint64_t orc_fa2_total_pairs_sampled(const orc_fa2 *m) { return m->total_pairs_sampled; }
uint64_t orc_fa2_nheldout(const orc_fa2 *m) { return m->nho; }
const uint32_t *orc_fa2_heldout_pairs(const orc_fa2 *m) { return m->ho; }
const uint32_t *orc_fa2_heldout_sorted(const orc_fa2 *m) { return m->ho_sorted; }
const char *orc_fa2_heldout_log(const orc_fa2 *m) { return m->hlog ? m->hlog : ""; }

/* ---- writers -------------------------------------------------------------------------------- */
int orc_fa2_write_outputs(orc_fa2 *m, const char *dir) {
  const uint32_t n = m->n, k = m->k;
  const orc_graph *g = m->g;
  char path[4096];
  FILE *f;
  /* save_model, fastamm2.cc:705-739 */
  snprintf(path, sizeof path, "%s/gamma.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  for (uint32_t i = 0; i < n; ++i) {
    fprintf(f, "%d\t%d\t", i, g->seq2id[i]);
    for (uint32_t c = 0; c < k; ++c) fprintf(f, c == k - 1 ? "%.5f\n" : "%.5f\t", m->gamma[(size_t)i * k + c]);
  }
  fclose(f);
  snprintf(path, sizeof path, "%s/lambda.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  for (uint32_t c = 0; c < k; ++c) fprintf(f, "%d\t%.5f\t%.5f\n", c, m->lambda[2 * c], m->lambda[2 * c + 1]);
  fclose(f);
  snprintf(path, sizeof path, "%s/heldout.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  fputs(orc_fa2_heldout_log(m), f);
  fclose(f);
  snprintf(path, sizeof path, "%s/heldout-pairs.txt", dir);    /* init_heldout, :326-327 */
  if (!(f = fopen(path, "w"))) return -1;
  for (uint64_t i = 0; i < m->nho; ++i) fprintf(f, "%d\t%d\n", g->seq2id[m->ho[2 * i]], g->seq2id[m->ho[2 * i + 1]]);
  fprintf(f, "\n");
  fclose(f);

  /* estimate_all_pi (fastamm2.hh:451-463) + compute_and_log_groups (fastamm2.cc:743-876) */
  double *epi = (double *)calloc((size_t)(n ? n : 1) * k, sizeof(double));
  for (uint32_t i = 0; i < n; ++i) {
    double s = .0;
    for (uint32_t c = 0; c < k; ++c) s += m->gamma[(size_t)i * k + c];
    for (uint32_t c = 0; c < k; ++c) epi[(size_t)i * k + c] = m->gamma[(size_t)i * k + c] / s;
  }
  double *beta = (double *)calloc(k, sizeof(double));
  for (uint32_t c = 0; c < k; ++c) beta[c] = m->lambda[2 * c] / (m->lambda[2 * c] + m->lambda[2 * c + 1]);
  uint32_t *groups = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  /* _communities[max_k] as growing vectors */
  uint32_t **comm = (uint32_t **)calloc(k, sizeof(uint32_t *));
  uint64_t *clen = (uint64_t *)calloc(k, sizeof(uint64_t)), *ccap = (uint64_t *)calloc(k, sizeof(uint64_t));
  uint32_t unlikely = 0;
  snprintf(path, sizeof path, "%s/groups.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  for (uint32_t i = 0; i < n; ++i) {
    fprintf(f, "%d\t%d\t", i, g->seq2id[i]);
    const double *pi_i = epi + (size_t)i * k;
    double max = .0;
    for (uint32_t j = 0; j < k; ++j) {
      fprintf(f, "%.3f\t", pi_i[j]);
      if (pi_i[j] > max) { max = pi_i[j]; groups[i] = j; }
    }
    for (uint64_t r = g->adj_off[i]; r < g->adj_off[i + 1]; ++r) {
      uint32_t mm = g->adj[r];
      if (i < mm) {
        const double *pi_m = epi + (size_t)mm * k;
        /* inner_prod_max, matrix.hh:459-476 */
        double u = .0, s = .0;
        uint32_t idx = 0;
        for (uint32_t c = 0; c < k; ++c) {
          double v = pi_i[c] * pi_m[c] * beta[c];
          s += v;
          if (v > u) { u = v; idx = c; }
        }
        if (u / s < m->link_thresh) { unlikely++; continue; }
        for (int e = 0; e < 2; ++e) {
          if (clen[idx] == ccap[idx]) {
            ccap[idx] = ccap[idx] ? ccap[idx] * 2 : 16;
            comm[idx] = (uint32_t *)realloc(comm[idx], ccap[idx] * sizeof(uint32_t));
          }
          comm[idx][clen[idx]++] = e == 0 ? i : mm;
        }
      }
    }
    fprintf(f, "%d\n", groups[i]);
  }
  fclose(f);
  snprintf(path, sizeof path, "%s/summary.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  uint32_t *sz = (uint32_t *)calloc(k, sizeof(uint32_t));
  for (uint32_t i = 0; i < n; ++i) sz[groups[i]]++;
  for (uint32_t c = 0; c < k; ++c) fprintf(f, "%d\t", sz[c]);
  fprintf(f, ":%d\n\n", unlikely);
  fclose(f);
  free(sz);
  snprintf(path, sizeof path, "%s/communities.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  snprintf(path, sizeof path, "%s/communities_size.txt", dir);
  FILE *fs = fopen(path, "w");
  if (!fs) { fclose(f); return -1; }
  uint8_t *seen = (uint8_t *)calloc(n ? n : 1, 1);
  for (uint32_t c = 0; c < k; ++c) {
    if (!clen[c]) continue;                       /* std::map holds only communities that got a link */
    uint64_t uniq = 0;
    for (uint64_t i = 0; i < clen[c]; ++i) {
      uint32_t u = comm[c][i];
      if (seen[u]) continue;
      seen[u] = 1; uniq++;
      fprintf(f, "%d ", g->seq2id[u]);
    }
    fprintf(f, "\n");
    fprintf(fs, "%d\t%ld\n", c, (long)uniq);
    for (uint64_t i = 0; i < clen[c]; ++i) seen[comm[c][i]] = 0;
  }
  fclose(f); fclose(fs);
  for (uint32_t c = 0; c < k; ++c) free(comm[c]);
  free(comm); free(clen); free(ccap); free(seen); free(groups); free(beta); free(epi);
  return 0;
}
