/* oracle_fa2.h -- CPU restatement of svinet's `-rnode -stratified` path (class FastAMM2).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as oracle_ls.h): nothing under svinet_b200/ may include,
 * link or call this.  Plain serial C, reference order of operations, FP64.
 *
 * Parity status: PINNED against the unmodified reference compiled here (oracle/_ref/svinet_ref)
 * through the fixtures tests/golden/fa2_* (tests/test_oracle_fa2_golden.py).  Unpinned boundary:
 * GSL (mt19937 stream, gsl_ran_shuffle, gsl_ran_gamma, gsl_ran_bernoulli restated; the gamma
 * variates follow the stand-in in oracle/gsl_shim, i.e. Marsaglia-Tsang with a polar normal, and
 * only shape the INITIAL gamma/lambda).
 *
 * Uninitialised members of the reference class (fastamm2.hh: _iter, _lambda_start_iter,
 * _zeros_prob, _ones_prob are never assigned before use) are taken as 0, which is what the
 * compiled reference exhibits (heldout.txt starts at iteration 0 and prints -0.000000000 for the
 * weighted columns).
 *
 * Citations are file:line relative to /root/reference/src.
 */
#ifndef ORACLE_FA2_H
#define ORACLE_FA2_H
#include "oracle_ls.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_fa2_options {
  uint32_t k;
  double   seed;               /* -seed, 0 = GSL default                          */
  double   heldout_ratio;      /* 0.01                                            */
  uint32_t max_iterations;     /* -max-iterations; the loop runs max_iterations+1 */
  uint32_t reportfreq;         /* 100 for -rnode -stratified (main.cc:138-151)    */
  double   eta0, eta1;         /* 1, 1                                            */
  double   epsilon;            /* 1e-30 (env.hh:395)                              */
  double   tau0, kappa;        /* 1024, 0.9  (env.hh:405-408)                     */
  double   nodetau0, nodekappa;/* 1024, 0.5                                       */
  uint32_t online_iterations;  /* 50 (env.hh:415)                                 */
  double   meanchangethresh;   /* 1e-5 (env.hh:337)                               */
  int      deterministic;      /* env.deterministic (init_gamma, fastamm2.cc:504) */
  int      nolambda;
} orc_fa2_options;

typedef struct orc_fa2 orc_fa2;

void     orc_fa2_options_default(orc_fa2_options *o, uint32_t k);
orc_fa2 *orc_fa2_create(const orc_graph *g, const orc_fa2_options *o);    /* fastamm2.cc:8-248 */
void     orc_fa2_free(orc_fa2 *m);

/* One pair's coordinate ascent, PhiCompute::update_phis_until_conv (fastamm2.hh:151-209).
 * elogpi_p/q: K-rows; elogf: K (compute_Elogf); returns the number of rounds executed. */
uint32_t orc_fa2_phi_pair(uint32_t k, const double *elogpi_p, const double *elogpi_q, const double *elogf,
                          int y, double logepsilon, uint32_t online_iterations, double thresh,
                          double *phi1, double *phi2);

/* The iteration in two halves so a test can drive a device engine with the same minibatch:
 *   plan    = the RNG draws + pair selection of opt_process / opt_process_noninf
 *             (fastamm2.cc:574-579, 936, 943-954, 1078, 1100-1125)
 *   process = set_dir_exp(lambda), the per-pair updates and both Robbins-Monro blends
 *             (fastamm2.cc:566-567, 956-1000, 1130-1163, 586-638), then _iter++            */
void     orc_fa2_plan(orc_fa2 *m);
void     orc_fa2_process(orc_fa2 *m);
/* infer(): plan+process+report until max_iterations (or `max_steps` of this call, 0 = no bound);
 * returns the number of iterations executed by this call */
uint32_t orc_fa2_run(orc_fa2 *m, uint32_t max_steps);

/* accessors */
uint32_t orc_fa2_n(const orc_fa2 *m);
uint32_t orc_fa2_k(const orc_fa2 *m);
uint32_t orc_fa2_iter(const orc_fa2 *m);
int      orc_fa2_stopped(const orc_fa2 *m);
double  *orc_fa2_gamma(orc_fa2 *m);                  /* [n*k] */
double  *orc_fa2_lambda(orc_fa2 *m);                 /* [k*2] */
double   orc_fa2_alpha(const orc_fa2 *m);
const uint32_t *orc_fa2_shuffled(const orc_fa2 *m);  /* [n] */
/* the planned minibatch: type (0 = links of start node, 1 = non-informative set), start node,
 * pairs as (p<q) couples; all of one type, so y = 1 - type */
uint32_t orc_fa2_plan_type(const orc_fa2 *m);
uint32_t orc_fa2_plan_start(const orc_fa2 *m);
uint64_t orc_fa2_plan_npairs(const orc_fa2 *m);
const uint32_t *orc_fa2_plan_pairs(const orc_fa2 *m);
uint64_t orc_fa2_total_pairs_sampled(const orc_fa2 *m);
uint64_t orc_fa2_nheldout(const orc_fa2 *m);
const uint32_t *orc_fa2_heldout_pairs(const orc_fa2 *m);   /* [2*nheldout], draw order */
const uint32_t *orc_fa2_heldout_sorted(const orc_fa2 *m);  /* std::map iteration order */
/* FastAMM2::edge_likelihood (fastamm2.hh:477-520) */
double   orc_fa2_edge_likelihood(const orc_fa2 *m, uint32_t p, uint32_t q, int y);
/* heldout.txt lines accumulated so far (duration column printed as 0) */
const char *orc_fa2_heldout_log(const orc_fa2 *m);
/* gamma.txt, lambda.txt (save_model, fastamm2.cc:705-739), heldout.txt, heldout-pairs.txt,
 * groups.txt, communities.txt, communities_size.txt, summary.txt (compute_and_log_groups, :743-876) */
int      orc_fa2_write_outputs(orc_fa2 *m, const char *dir);

#ifdef __cplusplus
}
#endif
#endif
