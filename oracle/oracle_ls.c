/* oracle_ls.c -- CPU restatement of svinet's `-link-sampling` path (see oracle_ls.h).
 *
 * TEST INFRASTRUCTURE ONLY: the checker, never the product.  Serial, FP64, reference
 * order of operations (push form over the _links array, running log-sum-exp), i.e. a
 * deliberately DIFFERENT formulation from the CUDA path (pull form over CSR segments,
 * factorised exp), so that agreement between the two is evidence and not tautology.
 *
 * All file:line citations are relative to /root/reference/src.
 */
#define _POSIX_C_SOURCE 200809L
#include "oracle_ls.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ===================================================================================
 * GSL boundary.  GSL is an un-vendored dependency of the reference (configure.ac:14-16,
 * version not pinned).  Restated from the published algorithms:
 *   - gsl_rng_default = mt19937 (Matsumoto & Nishimura, 2002 seeding), seed 0 -> 4357
 *   - gsl_rng_uniform  = genrand_int32 / 2^32
 *   - gsl_rng_uniform_int(n): scale = 0xffffffff / n; repeat k = next/scale until k < n
 * Call sites on the path: linksampling.cc:71-75,392; linksampling.hh:336,337,344.
 * =================================================================================== */
void orc_rng_seed(orc_rng *r, unsigned long s) {
  if (s == 0) s = 4357;
  r->mt[0] = (uint32_t)(s & 0xffffffffUL);
  for (int i = 1; i < 624; ++i)
    r->mt[i] = 1812433253U * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i;
  r->mti = 624;
}

uint32_t orc_rng_next(orc_rng *r) {
  static const uint32_t mag01[2] = {0U, 0x9908b0dfU};
  uint32_t y;
  if (r->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
      y = (r->mt[kk] & 0x80000000U) | (r->mt[kk + 1] & 0x7fffffffU);
      r->mt[kk] = r->mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1U];
    }
    for (; kk < 623; kk++) {
      y = (r->mt[kk] & 0x80000000U) | (r->mt[kk + 1] & 0x7fffffffU);
      r->mt[kk] = r->mt[kk - 227] ^ (y >> 1) ^ mag01[y & 1U];
    }
    y = (r->mt[623] & 0x80000000U) | (r->mt[0] & 0x7fffffffU);
    r->mt[623] = r->mt[396] ^ (y >> 1) ^ mag01[y & 1U];
    r->mti = 0;
  }
  y = r->mt[r->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680U;
  y ^= (y << 15) & 0xefc60000U;
  y ^= (y >> 18);
  return y;
}

double orc_rng_uniform(orc_rng *r) { return orc_rng_next(r) / 4294967296.0; }

unsigned long orc_rng_uniform_int(orc_rng *r, unsigned long n) {
  unsigned long scale = 0xffffffffUL / n, k;
  do { k = orc_rng_next(r) / scale; } while (k >= n);
  return k;
}

/* gsl_sf_psi for x > 0 (call sites linksampling.hh:181,184,198,200): upward recurrence
 * psi(x) = psi(x+1) - 1/x until x >= 10, then ln x - 1/(2x) - sum B_2n/(2n x^2n). */
double orc_digamma(double x) {
  double acc = 0.0;
  if (!(x > 0.0)) return NAN;
  while (x < 10.0) { acc -= 1.0 / x; x += 1.0; }
  double inv = 1.0 / x, inv2 = inv * inv;
  double series = inv2 * (1.0 / 12.0 - inv2 * (1.0 / 120.0 - inv2 * (1.0 / 252.0 - inv2 * (1.0 / 240.0
                  - inv2 * (1.0 / 132.0 - inv2 * (691.0 / 32760.0 - inv2 * (1.0 / 12.0)))))));
  return acc + log(x) - 0.5 * inv - series;
}

/* ===================================================================================
 * Graph ingest: Network::read (network.cc:11-159) with Network::add (network.hh:134-148)
 * and Network::y (network.hh:158-176).
 * =================================================================================== */
typedef struct { uint32_t *v; uint32_t len, cap; } u32vec;

static void u32vec_push(u32vec *a, uint32_t x) {
  if (a->len == a->cap) {
    a->cap = a->cap ? a->cap * 2 : 4;
    a->v = (uint32_t *)realloc(a->v, (size_t)a->cap * sizeof(uint32_t));
  }
  a->v[a->len++] = x;
}

typedef struct { uint32_t *keys, *vals; uint8_t *used; uint64_t cap; } idmap;

static void idmap_init(idmap *m, uint64_t expect) {
  m->cap = 16;
  while (m->cap < expect * 2 + 2) m->cap <<= 1;
  m->keys = (uint32_t *)calloc(m->cap, sizeof(uint32_t));
  m->vals = (uint32_t *)calloc(m->cap, sizeof(uint32_t));
  m->used = (uint8_t *)calloc(m->cap, 1);
}
static void idmap_free(idmap *m) { free(m->keys); free(m->vals); free(m->used); }
static uint64_t idmap_slot(const idmap *m, uint32_t key) {
  uint64_t h = ((uint64_t)key * 0x9E3779B97F4A7C15ULL) >> 20;
  h &= m->cap - 1;
  while (m->used[h] && m->keys[h] != key) h = (h + 1) & (m->cap - 1);
  return h;
}

typedef struct {
  orc_graph *g;
  idmap id2seq;
  u32vec *nbr;      /* per-sequence-id adjacency, insertion order */
  u32vec edges;
  uint32_t curr_seq;
} gbuild;

/* Network::add (network.hh:134-148): refuse new ids once n_arg sequence ids are used */
static int gb_add(gbuild *b, uint32_t id) {
  if (b->curr_seq >= b->g->n_arg) return 0;
  uint64_t s = idmap_slot(&b->id2seq, id);
  b->id2seq.used[s] = 1; b->id2seq.keys[s] = id; b->id2seq.vals[s] = b->curr_seq;
  b->g->seq2id[b->curr_seq] = id;
  b->curr_seq++;
  return 1;
}

/* Network::y (network.hh:158-176): linear scan of the lower endpoint's list */
static int gb_y(const gbuild *b, uint32_t p, uint32_t q) {
  uint32_t lo = p < q ? p : q, hi = p < q ? q : p;
  const u32vec *v = &b->nbr[lo];
  for (uint32_t j = 0; j < v->len; ++j) if (v->v[j] == hi) return 1;
  return 0;
}

/* one input line, network.cc:56-98 */
static void gb_line(gbuild *b, uint32_t id1, uint32_t id2) {
  uint64_t s1 = idmap_slot(&b->id2seq, id1);
  if (!b->id2seq.used[s1] && !gb_add(b, id1)) return;
  uint64_t s2 = idmap_slot(&b->id2seq, id2);
  if (!b->id2seq.used[s2] && !gb_add(b, id2)) return;
  uint32_t p = b->id2seq.vals[idmap_slot(&b->id2seq, id1)];
  uint32_t q = b->id2seq.vals[idmap_slot(&b->id2seq, id2)];
  if (p != q && gb_y(b, p, q) == 0) {
    uint32_t lo = p < q ? p : q, hi = p < q ? q : p;
    u32vec_push(&b->edges, lo);
    u32vec_push(&b->edges, hi);
    u32vec_push(&b->nbr[lo], hi);
    u32vec_push(&b->nbr[hi], lo);
    b->g->ones++;
  }
}

static gbuild *gb_begin(uint32_t n_arg) {
  gbuild *b = (gbuild *)calloc(1, sizeof(gbuild));
  b->g = (orc_graph *)calloc(1, sizeof(orc_graph));
  b->g->n_arg = n_arg;
  b->g->seq2id = (uint32_t *)calloc(n_arg ? n_arg : 1, sizeof(uint32_t));
  b->nbr = (u32vec *)calloc(n_arg ? n_arg : 1, sizeof(u32vec));
  idmap_init(&b->id2seq, n_arg);
  return b;
}

static orc_graph *gb_finish(gbuild *b) {
  orc_graph *g = b->g;
  /* pad with synthetic single nodes, network.cc:107-113 (SINGLE_NODE_START_ID = 100000) */
  if (b->curr_seq != g->n_arg) {
    g->singles = g->n_arg - b->curr_seq;
    for (uint32_t c = b->curr_seq, k = 0; c < g->n_arg; ++c, ++k) gb_add(b, 100000u + k);
  }
  g->n = g->n_arg - g->singles;                    /* main.cc:291 */
  g->adj_off = (uint64_t *)calloc((size_t)g->n_arg + 1, sizeof(uint64_t));
  for (uint32_t i = 0; i < g->n_arg; ++i) g->adj_off[i + 1] = g->adj_off[i] + b->nbr[i].len;
  g->adj = (uint32_t *)malloc((size_t)(g->adj_off[g->n_arg] ? g->adj_off[g->n_arg] : 1) * sizeof(uint32_t));
  for (uint32_t i = 0; i < g->n_arg; ++i) {
    if (b->nbr[i].len) memcpy(g->adj + g->adj_off[i], b->nbr[i].v, (size_t)b->nbr[i].len * sizeof(uint32_t));
    free(b->nbr[i].v);
  }
  g->edges = b->edges.v ? b->edges.v : (uint32_t *)malloc(sizeof(uint32_t));
  free(b->nbr);
  idmap_free(&b->id2seq);
  free(b);
  return g;
}

orc_graph *orc_graph_read(const char *path, uint32_t n_arg) {
  FILE *f = fopen(path, "r");
  if (!f) return NULL;
  gbuild *b = gb_begin(n_arg);
  unsigned a, c;
  while (fscanf(f, "%u %u", &a, &c) == 2) gb_line(b, a, c);   /* "%d\t%d\n", network.cc:49 */
  fclose(f);
  return gb_finish(b);
}

orc_graph *orc_graph_from_pairs(const uint32_t *pairs, uint64_t npairs, uint32_t n_arg) {
  gbuild *b = gb_begin(n_arg);
  for (uint64_t i = 0; i < npairs; ++i) gb_line(b, pairs[2 * i], pairs[2 * i + 1]);
  return gb_finish(b);
}

void orc_graph_free(orc_graph *g) {
  if (!g) return;
  free(g->seq2id); free(g->adj_off); free(g->adj); free(g->edges); free(g);
}

int orc_graph_y(const orc_graph *g, uint32_t a, uint32_t b) {
  uint32_t lo = a < b ? a : b, hi = a < b ? b : a;
  for (uint64_t j = g->adj_off[lo]; j < g->adj_off[lo + 1]; ++j) if (g->adj[j] == hi) return 1;
  return 0;
}

/* ===================================================================================
 * State
 * =================================================================================== */
orc_state *orc_state_alloc(uint32_t n, uint32_t k, uint64_t nlinks) {
  orc_state *s = (orc_state *)calloc(1, sizeof(orc_state));
  size_t nk = (size_t)n * k;
  uint32_t kk = k / 10 ? k / 10 : 1;
  s->n = n; s->k = k; s->nlinks = nlinks;
  s->links = (uint32_t *)calloc(nlinks ? 2 * nlinks : 1, sizeof(uint32_t));
  s->tl = (double *)calloc(n ? n : 1, sizeof(double));
  s->gamma = (double *)calloc(nk ? nk : 1, sizeof(double));
  s->gammanext = (double *)calloc(nk ? nk : 1, sizeof(double));
  s->Elogpi = (double *)calloc(nk ? nk : 1, sizeof(double));
  s->mphi = (double *)calloc(nk ? nk : 1, sizeof(double));
  s->lambda = (double *)calloc(2 * (size_t)k, sizeof(double));
  s->lambdanext = (double *)calloc(2 * (size_t)k, sizeof(double));
  s->Elogbeta = (double *)calloc(2 * (size_t)k, sizeof(double));
  s->s1 = (double *)calloc(k, sizeof(double));
  s->s2 = (double *)calloc(k, sizeof(double));
  s->s3 = (double *)calloc(k, sizeof(double));
  s->sum = (double *)calloc(k, sizeof(double));
  s->converged = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  s->active_comms = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  s->active_k = (uint16_t *)calloc((size_t)(n ? n : 1) * kk, sizeof(uint16_t));
  s->active_len = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  s->member = (uint8_t *)calloc(nk ? nk : 1, 1);
  return s;
}

void orc_state_free(orc_state *s) {
  if (!s) return;
  free(s->links); free(s->tl); free(s->gamma); free(s->gammanext); free(s->Elogpi); free(s->mphi);
  free(s->lambda); free(s->lambdanext); free(s->Elogbeta); free(s->s1); free(s->s2); free(s->s3);
  free(s->sum); free(s->converged); free(s->active_comms); free(s->active_k); free(s->active_len);
  free(s->member); free(s);
}

/* LinkSampling::set_dir_exp, linksampling.hh:171-187 */
void orc_set_dir_exp(const double *u, double *e, uint32_t rows, uint32_t cols) {
  for (uint32_t i = 0; i < rows; ++i) {
    double s = .0;
    for (uint32_t j = 0; j < cols; ++j) s += u[(size_t)i * cols + j];
    double psi_sum = orc_digamma(s);
    for (uint32_t j = 0; j < cols; ++j) e[(size_t)i * cols + j] = orc_digamma(u[(size_t)i * cols + j]) - psi_sum;
  }
}

/* LinkSampling::prune + check_and_set_converged, linksampling.cc:456-491 */
void orc_prune(orc_state *s) {
  const uint32_t k = s->k, lim = k / 10, stride = lim ? lim : 1;
  for (uint32_t p = 0; p < s->n; ++p) {
    uint32_t active = 0, pk = 0, len = 0;
    for (uint32_t c = 0; c < k; ++c)
      if (s->gamma[(size_t)p * k + c] - s->alpha >= 1) {
        active++;
        if (active <= lim) s->active_k[(size_t)p * stride + len++] = (uint16_t)c;
        pk = c;
      }
    if (active > lim) len = 0;
    if (active == 1) s->converged[p] = pk + 1;
    s->active_comms[p] = active;
    s->active_len[p] = len;
  }
}

/* link-community tally, linksampling.cc:668-681 / 704-717 with link_thresh = lt_min_deg = 0
 * (both members are uninitialised in the reference and read as ~0) */
static void tally(orc_state *s, uint32_t p, uint32_t q, const double *phi) {
  uint32_t max_k = 65535;
  double maxv = .0;                                    /* D1Array::max, matrix.hh:521-532 */
  for (uint32_t i = 0; i < s->k; ++i) if (phi[i] > maxv) { maxv = phi[i]; max_k = i; }
  if (maxv > 0.0) {
    s->member[(size_t)p * s->k + max_k] = 1;
    s->member[(size_t)q * s->k + max_k] = 1;
  }
}

/* Loop body of LinkSampling::infer, linksampling.cc:584-761 */
void orc_step(orc_state *s, uint32_t iter, int annealing, int write_comm) {
  const uint32_t n = s->n, k = s->k, lim = k / 10, stride = lim ? lim : 1;
  const size_t nk = (size_t)n * k;
  double *gnext = s->gammanext, *lnext = s->lambdanext;
  double *phi = (double *)calloc(k ? k : 1, sizeof(double));
  uint16_t *uni = (uint16_t *)calloc(2 * stride + 2, sizeof(uint16_t));

  if (write_comm) memset(s->member, 0, nk);            /* :584-587 */
  memset(s->s1, 0, k * sizeof(double));                /* clear(), :548-554 */
  memset(s->s2, 0, k * sizeof(double));
  memset(s->s3, 0, k * sizeof(double));
  memset(s->sum, 0, k * sizeof(double));
  s->cnt_dense = s->cnt_sparse = s->cnt_shortcut = 0;

  /* ---- phi sweep, :605-725 ---- */
  for (uint64_t e = 0; e < s->nlinks; ++e) {
    const uint32_t p = s->links[2 * e], q = s->links[2 * e + 1];
    const uint32_t pc = s->converged[p], qc = s->converged[q];
    double *gp = gnext + (size_t)p * k, *gq = gnext + (size_t)q * k;
    const double *ep = s->Elogpi + (size_t)p * k, *eq = s->Elogpi + (size_t)q * k;
    if (pc && !qc) {                                   /* :622-626 */
      gp[pc - 1] += 1; gq[pc - 1] += 1; s->sum[pc - 1] += 2; lnext[2 * (pc - 1)] += 2;
      s->cnt_shortcut++;
    } else if (!pc && qc) {                            /* :627-631 */
      gq[qc - 1] += 1; gp[qc - 1] += 1; s->sum[qc - 1] += 2; lnext[2 * (qc - 1)] += 2;
      s->cnt_shortcut++;
    } else if (iter > 1000 && s->active_comms[p] < lim && s->active_comms[q] < lim) {
      /* sparse active-set phi, :634-681: sorted unique union of the two active lists */
      uint32_t a = 0, b = 0, m = 0;
      const uint16_t *la = s->active_k + (size_t)p * stride, *lb = s->active_k + (size_t)q * stride;
      const uint32_t na = s->active_len[p], nb = s->active_len[q];
      while (a < na || b < nb) {
        uint16_t v;
        if (b >= nb || (a < na && la[a] <= lb[b])) v = la[a++]; else v = lb[b++];
        if (m == 0 || uni[m - 1] != v) uni[m++] = v;
      }
      memset(phi, 0, k * sizeof(double));
      double r = .0;
      for (uint32_t i = 0; i < m; ++i) {
        const uint32_t c = uni[i];
        phi[c] = ep[c] + eq[c] + s->Elogbeta[2 * c];
        if (i == 0) r = phi[c];
        else if (phi[c] < r) r = r + log(1 + exp(phi[c] - r));
        else r = phi[c] + log(1 + exp(r - phi[c]));
      }
      for (uint32_t i = 0; i < m; ++i) phi[uni[i]] = exp(phi[uni[i]] - r);
      for (uint32_t i = 0; i < m; ++i) {
        const uint32_t c = uni[i];
        gp[c] += phi[c]; gq[c] += phi[c]; lnext[2 * c] += 2 * phi[c]; s->sum[c] += 2 * phi[c];
      }
      s->cnt_sparse++;
      if (write_comm) tally(s, p, q, phi);
    } else {                                           /* dense phi, :685-717 */
      double r = .0;
      for (uint32_t c = 0; c < k; ++c) {
        phi[c] = ep[c] + eq[c] + s->Elogbeta[2 * c];
        if (c == 0) r = phi[c];
        else if (phi[c] < r) r = r + log(1 + exp(phi[c] - r));
        else r = phi[c] + log(1 + exp(r - phi[c]));
      }
      for (uint32_t c = 0; c < k; ++c) phi[c] = exp(phi[c] - r);       /* matrix.hh:320-325 */
      for (uint32_t c = 0; c < k; ++c) {
        gp[c] += phi[c]; gq[c] += phi[c]; lnext[2 * c] += 2 * phi[c]; s->sum[c] += 2 * phi[c];
      }
      s->cnt_dense++;
      if (write_comm) tally(s, p, q, phi);
    }
  }

  /* ---- compute_mean_indicators, :526-545 ---- */
  for (uint32_t p = 0; p < n; ++p) {
    if (s->tl[p] == 0) continue;
    for (uint32_t c = 0; c < k; ++c) {
      double *m = s->mphi + (size_t)p * k + c, *g = gnext + (size_t)p * k + c;
      *m = (*g - s->alpha) / s->tl[p];
      s->s1[c] += *m;
      s->s2[c] += *m * *m;
      *g += (n - s->tl[p] - 1) * *m;
      if (annealing) *g *= s->ones / s->sum[c];
    }
  }

  /* ---- s3 sweep, :731-746 (Q4: the shortcut reads column pc, not pc-1; column k is the
   *      never-written heap slack after the row and reads as 0) ---- */
  for (uint64_t e = 0; e < s->nlinks; ++e) {
    const uint32_t p = s->links[2 * e], q = s->links[2 * e + 1];
    const uint32_t pc = s->converged[p], qc = s->converged[q];
    const double *mp = s->mphi + (size_t)p * k, *mq = s->mphi + (size_t)q * k;
    if (pc && !qc) s->s3[pc - 1] += pc < k ? mq[pc] : 0.0;
    else if (!pc && qc) s->s3[qc - 1] += qc < k ? mp[qc] : 0.0;
    else for (uint32_t c = 0; c < k; ++c) s->s3[c] += mp[c] * mq[c];
  }

  for (uint32_t c = 0; c < k; ++c)                    /* :748-749 */
    lnext[2 * c + 1] += s->s1[c] * s->s1[c] - s->s2[c] - s->s3[c];

  /* swap + reset, :751-755 */
  { double *t = s->gamma; s->gamma = s->gammanext; s->gammanext = t; }
  { double *t = s->lambda; s->lambda = s->lambdanext; s->lambdanext = t; }
  for (size_t i = 0; i < nk; ++i) s->gammanext[i] = s->alpha;
  for (uint32_t c = 0; c < k; ++c) { s->lambdanext[2 * c] = s->eta0; s->lambdanext[2 * c + 1] = s->eta1; }

  orc_set_dir_exp(s->gamma, s->Elogpi, n, k);          /* :757-759 */
  orc_set_dir_exp(s->lambda, s->Elogbeta, k, 2);
  orc_prune(s);                                        /* :761 */
  free(phi); free(uni);
}

/* LinkSampling::edge_likelihood, linksampling.hh:259-292 */
double orc_edge_likelihood(const orc_state *st, uint32_t p, uint32_t q, int y, double epsilon) {
  const uint32_t k = st->k;
  const double *gp = st->gamma + (size_t)p * k, *gq = st->gamma + (size_t)q * k;
  double sp = .0, sq = .0, s = .0;
  for (uint32_t c = 0; c < k; ++c) sp += gp[c];
  for (uint32_t c = 0; c < k; ++c) sq += gq[c];
  if (y == 1) {
    for (uint32_t z = 0; z < k; ++z) {
      double tot = .0;
      for (uint32_t t = 0; t < 2; ++t) tot += st->lambda[2 * z + t];
      double brate = st->lambda[2 * z] / tot;
      s += (gp[z] / sp) * (gq[z] / sq) * brate;
    }
  } else {
    for (uint32_t zp = 0; zp < k; ++zp)
      for (uint32_t zq = 0; zq < k; ++zq) {
        double brate;
        if (zp == zq) {
          double tot = .0;
          for (uint32_t t = 0; t < 2; ++t) tot += st->lambda[2 * zp + t];
          brate = st->lambda[2 * zp] / tot;
        } else brate = epsilon;
        s += (gp[zp] / sp) * (gq[zq] / sq) * (1 - brate);
      }
  }
  if (s < 1e-30) s = 1e-30;
  return log(s);
}

/* ===================================================================================
 * Whole run
 * =================================================================================== */
struct orc_model {
  const orc_graph *g;
  orc_options o;
  orc_state *s;
  orc_rng rng;
  double total_pairs, ones_prob, zeros_prob;
  uint32_t *val;            /* [2*nval] validation pairs in draw order (ordered)           */
  uint64_t nval;
  uint32_t *val_sorted;     /* same pairs in std::map<Edge,bool> (lexicographic) order      */
  uint32_t iter;
  int annealing, write_comm, stopped;
  double prev_h, max_h, max_t;
  uint32_t nh;
  /* logs */
  char *vlog; size_t vlog_len, vlog_cap;
  char maxline[256];
};

void orc_options_default(orc_options *o, uint32_t k) {
  memset(o, 0, sizeof(*o));
  o->k = k; o->seed = 0; o->heldout_ratio = 0.01; o->accuracy = 0; o->max_iterations = 0;
  o->use_validation_stop = 1; o->reportfreq = 1; o->eta0 = 1; o->eta1 = 1; o->epsilon = 1e-30;
}

static int val_contains(const orc_model *m, uint32_t a, uint32_t b) {
  /* std::map<Edge,bool>::find; linear scan is fine at oracle sizes... but keep it O(log) */
  uint64_t lo = 0, hi = m->nval;
  while (lo < hi) {
    uint64_t mid = (lo + hi) / 2;
    uint32_t x = m->val_sorted[2 * mid], y = m->val_sorted[2 * mid + 1];
    if (x == a && y == b) return 1;
    if (x < a || (x == a && y < b)) lo = mid + 1; else hi = mid;
  }
  return 0;
}

static void val_insert(orc_model *m, uint32_t a, uint32_t b) {
  m->val = (uint32_t *)realloc(m->val, (m->nval + 1) * 2 * sizeof(uint32_t));
  m->val_sorted = (uint32_t *)realloc(m->val_sorted, (m->nval + 1) * 2 * sizeof(uint32_t));
  m->val[2 * m->nval] = a; m->val[2 * m->nval + 1] = b;
  uint64_t pos = m->nval;
  while (pos > 0 && (m->val_sorted[2 * (pos - 1)] > a ||
                     (m->val_sorted[2 * (pos - 1)] == a && m->val_sorted[2 * (pos - 1) + 1] > b))) {
    m->val_sorted[2 * pos] = m->val_sorted[2 * (pos - 1)];
    m->val_sorted[2 * pos + 1] = m->val_sorted[2 * (pos - 1) + 1];
    pos--;
  }
  m->val_sorted[2 * pos] = a; m->val_sorted[2 * pos + 1] = b;
  m->nval++;
}

/* LinkSampling::edge_ok (linksampling.hh:296-326) with empty test/precision maps */
static int edge_ok(const orc_model *m, uint32_t a, uint32_t b) {
  if (a == b) return 0;
  return !val_contains(m, a, b);
}

/* LinkSampling::get_random_edge, linksampling.hh:328-349 */
static void get_random_edge(orc_model *m, int link, uint32_t *a, uint32_t *b) {
  if (!link) {
    do {
      uint32_t f = (uint32_t)orc_rng_uniform_int(&m->rng, m->s->n);
      uint32_t s = (uint32_t)orc_rng_uniform_int(&m->rng, m->s->n);
      *a = f < s ? f : s; *b = f < s ? s : f;
      if (f == s) { *a = f; *b = s; }
    } while (!edge_ok(m, *a, *b));
  } else {
    do {
      uint32_t i = (uint32_t)orc_rng_uniform_int(&m->rng, m->g->ones);
      *a = m->g->edges[2 * i]; *b = m->g->edges[2 * i + 1];
    } while (!edge_ok(m, *a, *b));
  }
}

/* LinkSampling::set_validation_sample, linksampling.cc:281-309 */
static void set_validation_sample(orc_model *m, int sz) {
  int c0 = 0, c1 = 0, p = sz / 2;
  while (c0 < p || c1 < p) {
    uint32_t a, b;
    if (c0 == p) get_random_edge(m, 1, &a, &b); else get_random_edge(m, 0, &a, &b);
    int y = orc_graph_y(m->g, a, b);
    if (y == 0 && c0 < p) { c0++; val_insert(m, a, b); }
    if (y == 1 && c1 < p) { c1++; val_insert(m, a, b); }
  }
}

/* LinkSampling::init_gamma2, linksampling.cc:374-401 */
static void init_gamma2(orc_model *m) {
  orc_state *s = m->s;
  const uint32_t k = s->k;
  double *phi = (double *)calloc(k ? k : 1, sizeof(double));
  for (uint32_t p = 0; p < s->n; ++p)
    for (uint64_t r = m->g->adj_off[p]; r < m->g->adj_off[p + 1]; ++r) {
      uint32_t q = m->g->adj[r];
      if (p >= q) continue;
      for (uint32_t c = 0; c < k; ++c) phi[c] = orc_rng_uniform(&m->rng);
      double sum = .0;
      for (uint32_t c = 0; c < k; ++c) sum += phi[c];
      for (uint32_t c = 0; c < k; ++c) phi[c] = phi[c] / sum;
      for (uint32_t c = 0; c < k; ++c) s->gamma[(size_t)p * k + c] += phi[c];
      for (uint32_t c = 0; c < k; ++c) s->gamma[(size_t)q * k + c] += phi[c];
    }
  free(phi);
}

/* LinkSampling::assign_training_links, linksampling.cc:493-523 */
static void assign_training_links(orc_model *m) {
  orc_state *s = m->s;
  uint64_t nl = 0;
  for (uint32_t p = 0; p < s->n; ++p)
    for (uint64_t r = m->g->adj_off[p]; r < m->g->adj_off[p + 1]; ++r) {
      uint32_t q = m->g->adj[r];
      if (!m->o.accuracy) {
        uint32_t a = p < q ? p : q, b = p < q ? q : p;
        if (!edge_ok(m, a, b)) continue;
      }
      s->tl[p]++; s->tl[q]++;
      if (p >= q) continue;
      s->links[2 * nl] = p; s->links[2 * nl + 1] = q;
      nl++;
    }
  s->nlinks = nl;
}

static void vlog_append(orc_model *m, const char *line) {
  size_t l = strlen(line);
  if (m->vlog_len + l + 1 > m->vlog_cap) {
    m->vlog_cap = (m->vlog_cap + l + 1) * 2;
    m->vlog = (char *)realloc(m->vlog, m->vlog_cap);
  }
  memcpy(m->vlog + m->vlog_len, line, l + 1);
  m->vlog_len += l;
}

void orc_model_heldout(const orc_model *m, double *nshol, double *mean0, double *mean1,
                       uint32_t *k0, uint32_t *k1) {
  uint32_t kz = 0, ko = 0;
  double sz = 0, so = 0;
  for (uint64_t i = 0; i < m->nval; ++i) {
    uint32_t p = m->val_sorted[2 * i], q = m->val_sorted[2 * i + 1];
    int y = orc_graph_y(m->g, p, q);
    double u = orc_edge_likelihood(m->s, p, q, y, m->o.epsilon);
    if (y) { so += u; ko++; } else { sz += u; kz++; }
  }
  if (mean0) *mean0 = sz / kz;
  if (mean1) *mean1 = so / ko;
  if (k0) *k0 = kz;
  if (k1) *k1 = ko;
  if (nshol) *nshol = (m->zeros_prob * (sz / kz)) + (m->ones_prob * (so / ko));
}

/* LinkSampling::validation_likelihood, linksampling.cc:966-1050; returns 1 if the run must end */
static int validation_likelihood(orc_model *m) {
  if (m->o.accuracy) return 0;
  uint32_t k = 0, kzeros = 0, kones = 0;
  double s = .0, szeros = 0, sones = 0;
  for (uint64_t i = 0; i < m->nval; ++i) {
    uint32_t p = m->val_sorted[2 * i], q = m->val_sorted[2 * i + 1];
    int y = orc_graph_y(m->g, p, q);
    double u = orc_edge_likelihood(m->s, p, q, y, m->o.epsilon);
    s += u; k += 1;
    if (y) { sones += u; kones++; } else { szeros += u; kzeros++; }
  }
  double nshol = (m->zeros_prob * (szeros / kzeros)) + (m->ones_prob * (sones / kones));
  char line[512];
  snprintf(line, sizeof line, "%d\t%d\t%.9f\t%d\t%.9f\t%d\t%.9f\t%d\t%.9f\t%.9f\t%.9f\n",
           m->iter, 0, s / k, k, szeros / kzeros, kzeros, sones / kones, kones,
           m->zeros_prob * (szeros / kzeros), m->ones_prob * (sones / kones), nshol);
  vlog_append(m, line);

  double a = nshol;
  int stop = 0, why = -1;
  if (m->iter > 10) {
    if (a > m->prev_h && m->prev_h != 0 && fabs((a - m->prev_h) / m->prev_h) < 0.00001) {
      stop = 1; why = 100;
    } else if (a < m->prev_h) m->nh++;
    else if (a > m->prev_h) m->nh = 0;
    if (a > m->max_h) { m->max_h = a; m->max_t = 0; }
    if (m->nh > 2) { why = 1; stop = 1; }
  }
  m->prev_h = nshol;
  snprintf(m->maxline, sizeof m->maxline, "%d\t%d\t%.5f\t%.5f\t%.5f\t%d\n", m->iter, 0, a, m->max_t, m->max_h, why);
  if (m->annealing && stop) {
    m->annealing = 0; m->nh = 0; m->prev_h = 0;
  } else if (!m->annealing && stop) {
    if (m->o.use_validation_stop) return 1;
  }
  return 0;
}

/* LinkSampling::LinkSampling, linksampling.cc:5-155 */
/* LinkSampling::init_gamma_external, linksampling.cc:404-452, fed by Network::load_init_communities,
 * network.cc:374-437 (one community per line: external node ids).  For every node p and every adjacency entry of p
 * the SAME vector phi[k] = alpha + [k in communities(p)] * n / |communities(p)|, normalised by its plain sum
 * (matrix.hh:341-346), is added to row p of gamma, which starts at alpha.  No RNG use. */
static int init_gamma_external(orc_model *m, const char *path) {
  const orc_graph *g = m->g;
  orc_state *s = m->s;
  const uint32_t n = s->n, k = s->k;
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  uint32_t *cnt = (uint32_t *)calloc(n ? n : 1, sizeof(uint32_t));
  uint32_t *mem = NULL;                 /* (node, community) pairs in file order */
  size_t nmem = 0, cap = 0;
  char *line = NULL;
  size_t lcap = 0;
  uint32_t cid = 0;
  while (getline(&line, &lcap, f) > 0) {
    char *p = line, *e = NULL;
    int any = 0;
    for (;; p = e) {
      long u = strtol(p, &e, 10);
      if (p == e) break;
      uint32_t seq = n;
      for (uint32_t i = 0; i < g->n_arg; ++i) if (g->seq2id[i] == (uint32_t)u) { seq = i; break; }
      if (seq < n) {
        if (nmem == cap) { cap = cap ? 2 * cap : 256; mem = (uint32_t *)realloc(mem, 2 * cap * sizeof(uint32_t)); }
        mem[2 * nmem] = seq; mem[2 * nmem + 1] = cid; nmem++;
        cnt[seq]++;
      }
      any = 1;
    }
    if (any) cid++;
  }
  free(line);
  fclose(f);
  double *phi = (double *)calloc(k ? k : 1, sizeof(double));
  for (uint32_t p = 0; p < n; ++p) {
    double *row = s->gamma + (size_t)p * k;
    for (uint32_t c = 0; c < k; ++c) row[c] = s->alpha;
    for (uint32_t c = 0; c < k; ++c) phi[c] = s->alpha;
    for (size_t j = 0; j < nmem; ++j)
      if (mem[2 * j] == p && mem[2 * j + 1] < k) phi[mem[2 * j + 1]] += (double)n / cnt[p];
    double sum = .0;
    for (uint32_t c = 0; c < k; ++c) sum += phi[c];
    for (uint32_t c = 0; c < k; ++c) phi[c] = phi[c] / sum;
    const uint64_t deg = g->adj_off[p + 1] - g->adj_off[p];
    for (uint64_t r = 0; r < deg; ++r)
      for (uint32_t c = 0; c < k; ++c) row[c] += phi[c];
  }
  free(phi); free(mem); free(cnt);
  return 0;
}

orc_model *orc_model_create(const orc_graph *g, const orc_options *o) {
  orc_model *m = (orc_model *)calloc(1, sizeof(orc_model));
  m->g = g; m->o = *o;
  const uint32_t n = g->n, k = o->k;
  m->s = orc_state_alloc(n, k, g->ones);
  orc_state *s = m->s;
  s->alpha = (double)1 / k;                            /* env.hh:344 */
  s->eta0 = o->eta0; s->eta1 = o->eta1; s->ones = g->ones;
  uint32_t tp = n * (n - 1) / 2;                       /* uint32 arithmetic, :37 (Q6) */
  m->total_pairs = tp;
  m->ones_prob = (double)g->ones / m->total_pairs;     /* :49-50 */
  m->zeros_prob = 1 - m->ones_prob;
  m->max_t = m->max_h = m->prev_h = -2147483647;
  m->annealing = 1;
  orc_rng_seed(&m->rng, 0);                            /* gsl_rng_alloc -> default seed */
  if (o->seed) orc_rng_seed(&m->rng, (unsigned long)o->seed);   /* :74-75 */
  int s1 = (int)(o->heldout_ratio * g->ones);          /* init_validation, :167 */
  set_validation_sample(m, s1);
  if (o->init_communities) {                           /* :113-116 */
    if (init_gamma_external(m, o->init_communities)) { orc_model_free(m); return NULL; }
  } else {
    init_gamma2(m);                                    /* :117 */
  }
  for (size_t i = 0; i < (size_t)n * k; ++i) s->gammanext[i] = s->alpha;
  for (uint32_t c = 0; c < k; ++c) {                   /* init_lambda, :364-372 */
    s->lambda[2 * c] = s->lambdanext[2 * c] = s->eta0;
    s->lambda[2 * c + 1] = s->lambdanext[2 * c + 1] = s->eta1;
  }
  orc_set_dir_exp(s->gamma, s->Elogpi, n, k);          /* :123-124 */
  orc_set_dir_exp(s->lambda, s->Elogbeta, k, 2);
  m->iter = 0;                                         /* Q1 */
  validation_likelihood(m);                            /* :150 */
  /* top of infer(), :559-566 */
  memset(s->converged, 0, (size_t)(n ? n : 1) * sizeof(uint32_t));
  assign_training_links(m);
  m->write_comm = 0;
  return m;
}

void orc_model_free(orc_model *m) {
  if (!m) return;
  orc_state_free(m->s); free(m->val); free(m->val_sorted); free(m->vlog); free(m);
}

/* the while(1) of LinkSampling::infer, linksampling.cc:571-789 */
uint32_t orc_model_run(orc_model *m, uint32_t max_sweeps) {
  uint32_t sweeps = 0;
  while (!m->stopped) {
    if (m->o.max_iterations && m->iter > m->o.max_iterations) { m->stopped = 1; break; }
    if (max_sweeps && sweeps >= max_sweeps) break;
    if (m->o.max_iterations == 1) m->write_comm = 1;
    orc_step(m->s, m->iter, m->annealing, m->write_comm);
    sweeps++;
    m->write_comm = (m->iter % m->o.reportfreq == m->o.reportfreq - 1);
    if (m->iter % m->o.reportfreq == 0) {
      if (validation_likelihood(m)) { m->stopped = 1; break; }
    }
    m->iter++;
  }
  return sweeps;
}

orc_state *orc_model_state(orc_model *m) { return m->s; }
uint32_t orc_model_iter(const orc_model *m) { return m->iter; }
int orc_model_annealing(const orc_model *m) { return m->annealing; }
int orc_model_write_comm(const orc_model *m) { return m->write_comm; }
int orc_model_stopped(const orc_model *m) { return m->stopped; }
uint64_t orc_model_nvalidation(const orc_model *m) { return m->nval; }
const uint32_t *orc_model_validation_pairs(const orc_model *m) { return m->val; }

static int cmp_u32(const void *a, const void *b) {
  uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
  return x < y ? -1 : x > y;
}

int orc_model_write_outputs(const orc_model *m, const char *dir) {
  const orc_state *s = m->s;
  const uint32_t n = s->n, k = s->k;
  char path[4096];
  FILE *f;
  /* save_model, :805-837 */
  snprintf(path, sizeof path, "%s/gamma.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  for (uint32_t i = 0; i < n; ++i) {
    fprintf(f, "%d\t", i);
    fprintf(f, "%d\t", m->g->seq2id[i]);
    for (uint32_t c = 0; c < k; ++c)
      fprintf(f, c == k - 1 ? "%.5f\n" : "%.5f\t", s->gamma[(size_t)i * k + c]);
  }
  fclose(f);
  snprintf(path, sizeof path, "%s/lambda.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  for (uint32_t c = 0; c < k; ++c) {
    fprintf(f, "%d\t", c);
    fprintf(f, "%.5f\t", s->lambda[2 * c]);
    fprintf(f, "%.5f\n", s->lambda[2 * c + 1]);
  }
  fclose(f);
  /* write_communities, :883-917 */
  snprintf(path, sizeof path, "%s/communities.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  uint32_t *ids = (uint32_t *)malloc((size_t)(n ? n : 1) * sizeof(uint32_t));
  for (uint32_t c = 0; c < k; ++c) {
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < n; ++i) if (s->member[(size_t)i * k + c]) ids[cnt++] = m->g->seq2id[i];
    if (!cnt) continue;
    qsort(ids, cnt, sizeof(uint32_t), cmp_u32);
    for (uint32_t j = 0; j < cnt; ++j) fprintf(f, "%d ", ids[j]);
    fprintf(f, "\n");
  }
  free(ids);
  fclose(f);
  /* write_groups, :1453-1476 */
  snprintf(path, sizeof path, "%s/groups.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  for (uint32_t i = 0; i < n; ++i) {
    double sum = .0;
    for (uint32_t c = 0; c < k; ++c) sum += s->gamma[(size_t)i * k + c];
    fprintf(f, "%d\t%d\t", i, m->g->seq2id[i]);
    for (uint32_t c = 0; c < k; ++c)
      fprintf(f, c == k - 1 ? "%.3f\n" : "%.3f\t", s->gamma[(size_t)i * k + c] / sum);
  }
  fclose(f);
  snprintf(path, sizeof path, "%s/validation.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  if (m->vlog) fputs(m->vlog, f);
  fclose(f);
  snprintf(path, sizeof path, "%s/max.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  fputs(m->maxline, f);
  fclose(f);
  /* edgelist_s, :190-206 */
  snprintf(path, sizeof path, "%s/validation-edges.txt", dir);
  if (!(f = fopen(path, "w"))) return -1;
  for (uint64_t i = 0; i < m->nval; ++i) {
    uint32_t a = m->val[2 * i], b = m->val[2 * i + 1];
    fprintf(f, "%d\t%d\t%d\n", m->g->seq2id[a], m->g->seq2id[b], orc_graph_y(m->g, a, b));
  }
  fprintf(f, "\n");
  fclose(f);
  return 0;
}
