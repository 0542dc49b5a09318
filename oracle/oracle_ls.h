/* oracle_ls.h -- CPU restatement of svinet's `-link-sampling` path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under svinet_b200/ (the product) may include, link or
 * call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it,
 * and only as the checker.  Plain serial C, reference order of operations, FP64.
 *
 * Parity status: PINNED against the unmodified reference compiled here (oracle/_ref/svinet_ref,
 * see oracle/Makefile) through the committed fixtures in tests/golden/ (tests/test_oracle_golden.py).
 * The one unpinned boundary is GSL itself (absent from /root/reference, version not stated by
 * configure.ac:14-16): its MT19937 stream and digamma are restated from the published
 * algorithms; see DESIGN.md "Oracle".
 *
 * Every function cites the reference file:line (relative to /root/reference/src) it follows.
 */
#ifndef ORACLE_LS_H
#define ORACLE_LS_H
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- GSL boundary (restated) ------------------------------------------------------ */
typedef struct { uint32_t mt[624]; int mti; } orc_rng;
void     orc_rng_seed(orc_rng *r, unsigned long seed);           /* gsl_rng_set, 0 -> 4357 */
uint32_t orc_rng_next(orc_rng *r);
double   orc_rng_uniform(orc_rng *r);                            /* next/2^32 */
unsigned long orc_rng_uniform_int(orc_rng *r, unsigned long n);  /* rejection, scale=0xffffffff/n */
double   orc_digamma(double x);                                  /* gsl_sf_psi, x>0 */

/* ---- graph (network.cc:11-159, network.hh:134-176) -------------------------------- */
typedef struct orc_graph {
  uint32_t n_arg;      /* the -n argument (size of the adjacency table)                    */
  uint32_t n;          /* inference n = n_arg - singles (main.cc:291)                      */
  uint32_t singles;
  uint32_t ones;       /* undirected links kept after self-loop / duplicate removal        */
  uint32_t *seq2id;    /* [n_arg] external id of each sequence id                          */
  uint64_t *adj_off;   /* [n_arg+1] CSR offsets                                            */
  uint32_t *adj;       /* [2*ones] neighbours in INSERTION order (defines _links order)    */
  uint32_t *edges;     /* [2*ones] (first<second) pairs in read order (Network::_edges)    */
} orc_graph;

orc_graph *orc_graph_read(const char *path, uint32_t n_arg);
/* build from an in-memory list of (u,v) external-id pairs, same semantics as reading a file */
orc_graph *orc_graph_from_pairs(const uint32_t *pairs, uint64_t npairs, uint32_t n_arg);
void       orc_graph_free(orc_graph *g);
int        orc_graph_y(const orc_graph *g, uint32_t a, uint32_t b);   /* network.hh:158-176 */

/* ---- variational state (linksampling.hh:86-160) ------------------------------------ */
typedef struct orc_state {
  uint32_t n, k;
  double alpha, eta0, eta1;
  uint32_t ones;            /* numerator of the annealing rescale (linksampling.cc:542)    */
  uint64_t nlinks;
  uint32_t *links;          /* [2*nlinks] training links p<q in reference order (:494-523) */
  double *tl;               /* [n] _training_links = 2 x training degree (Q3)              */
  double *gamma, *gammanext, *Elogpi, *mphi;     /* [n*k] row-major                        */
  double *lambda, *lambdanext, *Elogbeta;        /* [k*2]                                  */
  double *s1, *s2, *s3, *sum;                    /* [k]                                    */
  uint32_t *converged;      /* [n] 0 or community+1, sticky (:456-475)                     */
  uint32_t *active_comms;   /* [n]                                                         */
  uint16_t *active_k;       /* [n * max(1,k/10)] first <=k/10 active communities           */
  uint32_t *active_len;     /* [n]                                                         */
  uint8_t  *member;         /* [n*k] node-in-link-community flags (fmap>0, :704-717)       */
  uint64_t cnt_dense, cnt_sparse, cnt_shortcut;  /* branch counters of the last sweep      */
} orc_state;

orc_state *orc_state_alloc(uint32_t n, uint32_t k, uint64_t nlinks);
void       orc_state_free(orc_state *s);
void       orc_set_dir_exp(const double *u, double *e, uint32_t rows, uint32_t cols);  /* linksampling.hh:171-187 */
void       orc_prune(orc_state *s);                                                    /* linksampling.cc:456-491 */

/* One full sweep of LinkSampling::infer's loop body, linksampling.cc:584-761
 * (clear .. prune).  `iter` is the reference's _iter (only used for the >1000 test, :634). */
void orc_step(orc_state *s, uint32_t iter, int annealing, int write_comm);

/* the same sweep over `threads` host cores (OpenMP; oracle_ls_omp.c): timing baseline, equal to orc_step up to the
 * summation order; dense + shortcut branches only (iter <= 1000).  threads < 1 = all. */
void orc_step_omp(orc_state *s, int annealing, int write_comm, int threads);
int  orc_omp_max_threads(void);

/* held-out log-likelihood of one pair, linksampling.hh:259-292 (literal O(K^2) non-link form) */
double orc_edge_likelihood(const orc_state *s, uint32_t p, uint32_t q, int y, double epsilon);

/* ---- whole run: ctor + infer() + do_on_stop() (linksampling.cc:5-155, 557-802) ------ */
typedef struct orc_model orc_model;

typedef struct orc_options {
  uint32_t k;
  double   seed;               /* -seed (0 = GSL default 4357)                 */
  double   heldout_ratio;      /* -heldout-ratio, default 0.01                 */
  int      accuracy;           /* -accuracy                                    */
  uint32_t max_iterations;     /* -max-iterations, 0 = unlimited               */
  int      use_validation_stop;/* cleared by -no-stop                          */
  uint32_t reportfreq;         /* 1 for -link-sampling (main.cc:149-153)       */
  double   eta0, eta1;         /* 1,1 for -eta-type uniform (network.cc:238)   */
  double   epsilon;            /* 1e-30 (env.hh:395)                           */
  const char *init_communities;/* -init-communities <file> (NULL = init_gamma2)  */
} orc_options;

void       orc_options_default(orc_options *o, uint32_t k);
orc_model *orc_model_create(const orc_graph *g, const orc_options *o);  /* the constructor */
void       orc_model_free(orc_model *m);
/* run infer(): returns number of sweeps executed; stops on max-iterations / validation stop.
 * max_sweeps > 0 additionally bounds the number of sweeps of THIS call (state stays resumable). */
uint32_t   orc_model_run(orc_model *m, uint32_t max_sweeps);
orc_state *orc_model_state(orc_model *m);
uint32_t   orc_model_iter(const orc_model *m);
int        orc_model_annealing(const orc_model *m);
int        orc_model_write_comm(const orc_model *m);
int        orc_model_stopped(const orc_model *m);
uint64_t   orc_model_nvalidation(const orc_model *m);
const uint32_t *orc_model_validation_pairs(const orc_model *m);  /* [2*nval] in draw order */
/* validation_likelihood (:967-1050) without the stop machine: fills a=nshol, a0, a1 means */
void       orc_model_heldout(const orc_model *m, double *nshol, double *mean0, double *mean1,
                             uint32_t *k0, uint32_t *k1);
/* text writers with the reference's exact formats (:805-837, :883-917, :1453-1476, :996-1002,
 * :1030-1034, :190-206); `dir` must exist */
int        orc_model_write_outputs(const orc_model *m, const char *dir);

#ifdef __cplusplus
}
#endif
#endif
