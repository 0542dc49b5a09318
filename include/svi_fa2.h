/* svi_fa2.h -- C ABI of the B200 engine for svinet's `-rnode -stratified` path (libsvi_ls.so).
 *
 * Drop-in boundary for the iteration of FastAMM2::infer (reference src/fastamm2.cc:535-702):
 * set_dir_exp(lambda) (:567), the per-pair two-phi coordinate ascent of opt_process /
 * opt_process_noninf (:933-1165, PhiCompute::update_phis_until_conv src/fastamm2.hh:151-209),
 * the Robbins-Monro blend of ALL gamma rows (:605-624) and of lambda (:626-638), plus the held-out
 * likelihood FastAMM2::edge_likelihood (src/fastamm2.hh:477-520).  The reference has no FFI; its
 * seam is the C++ class `FastAMM2` used at src/main.cc:368-372.  A replacement class keeps the
 * reference's host responsibilities (GSL stream, shuffle_nodes, held-out draw, init_gamma /
 * init_lambda, file writers) and drives the device through the calls below.
 *
 * Two ways to feed minibatches:
 *   svi_fa2_step  -- the HOST chose the minibatch (it replays the reference's mt19937 draws, so a
 *                    run is comparable with the reference iteration by iteration);
 *   svi_fa2_run   -- the DEVICE draws the minibatches from a counter-based Philox4x32-10 stream
 *                    keyed by (seed, iteration): no host round trip per iteration.  Same sampling
 *                    distribution as the reference (Bernoulli(0.5) set type, uniform start node,
 *                    uniform block of the shuffled node order), different variates.
 *
 * Conventions as in svi_ls.h: plain C types, caller-owned host buffers, 0 or a negative
 * svi_status, message from svi_ls_last_error().  FP64 throughout.
 */
#ifndef SVI_FA2_H
#define SVI_FA2_H

#include "svi_ls.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svi_fa2 svi_fa2; /* opaque */

/* Constructor state of FastAMM2 that the iteration reads (src/fastamm2.cc:8-48). */
typedef struct svi_fa2_config {
  uint32_t n, k;
  double   alpha;             /* env.alpha = 1/k (src/env.hh:344)                              */
  double   eta0, eta1;
  double   epsilon;           /* env.epsilon = 1e-30; logepsilon = log(epsilon) (env.hh:395)   */
  double   tau0, kappa;       /* _tau0 = env.tau0 + 1 = 1025, _kappa = 0.9 (fastamm2.cc:19)    */
  double   nodetau0, nodekappa; /* _nodetau0 = 1025, _nodekappa = 0.5 (fastamm2.cc:20)         */
  double   inf_epsilon;       /* _inf_epsilon = 0.5 (fastamm2.cc:15)                           */
  uint32_t m_sets;            /* _m = 10 non-informative sets per node (fastamm2.cc:11)        */
  uint32_t online_iterations; /* env.online_iterations = 50 (env.hh:415)                       */
  double   meanchangethresh;  /* env.meanchangethresh = 1e-5 (env.hh:337)                      */
  int32_t  nolambda;          /* env.nolambda                                                  */
  int32_t  device;            /* CUDA ordinal, -1 = current                                    */
  /* How the decay of the rows a minibatch does NOT touch (src/fastamm2.cc:614-620) is applied.  Every
   * node shares one step size (the reference bumps every _nodec[i] every iteration), so
   * (gamma - alpha) of all untouched rows shrinks by the same factor (1 - rho):
   *   0 (default): LAZY -- that factor is carried as one scalar c (rows store u, gamma = alpha + c*u);
   *                no O(N*K) pass per iteration, the result differs from the eager one by rounding only;
   *   1          : EAGER -- every row is read and written every iteration (k_fa2_blend), as the reference does. */
  int32_t  eager_blend;
} svi_fa2_config;

/* Fill `cfg` with the reference's defaults for (n, k). */
void svi_fa2_default_config(svi_fa2_config *cfg, uint32_t n, uint32_t k);

int  svi_fa2_create(const svi_fa2_config *cfg, svi_fa2 **out);
void svi_fa2_destroy(svi_fa2 *h);
int  svi_fa2_set_stream(svi_fa2 *h, void *cuda_stream);
int  svi_fa2_sync(svi_fa2 *h);

/* gamma [n*k row-major], lambda [k*2].  set: after init_gamma/init_lambda/load_model
 * (src/fastamm2.cc:133-143); also resets the per-node update counter _nodec to `nodec`
 * (0 at construction).  get: save_model / compute_and_log_groups (:705-739, :743-876). */
int svi_fa2_set_state(svi_fa2 *h, const double *gamma, const double *lambda, uint64_t nodec);
int svi_fa2_get_state(svi_fa2 *h, double *gamma, double *lambda);

/* One iteration of FastAMM2::infer's loop body (src/fastamm2.cc:566-640) on a host-chosen minibatch.
 *   iter   : the reference's _iter (rho_t = (tau0 + iter + 1)^-kappa, :627)
 *   type   : 0 = the links of `start` (opt_process, y = 1), 1 = a non-informative set
 *            (opt_process_noninf, y = 0); selects `scale` (:591-592) and the lambda column
 *   start  : _start_node
 *   pairs  : [2*npairs] (p<q) couples, each containing `start`; the other endpoint of every pair
 *            must be distinct (true for both samplers of the reference)
 * Asynchronous on the handle's stream (the pair list is copied before returning). */
int svi_fa2_step(svi_fa2 *h, uint32_t iter, uint32_t type, uint32_t start, uint64_t npairs,
                 const uint32_t *pairs);

/* Graph + held-out set + shuffled node order for device-side minibatch draws.
 *   links    : [2*nlinks] (p<q), all links of the network (Network::_edges)
 *   heldout  : [2*nheldout] (p<q) pairs excluded from training (FastAMM2::edge_ok, fastamm2.hh:524)
 *   shuffled : [n] _shuffled_nodes (fastamm2.cc:489-495) */
int svi_fa2_set_graph(svi_fa2 *h, uint64_t nlinks, const uint32_t *links, uint64_t nheldout,
                      const uint32_t *heldout, const uint32_t *shuffled);
/* `iters` iterations starting at _iter = iter0 with device-drawn minibatches (needs set_graph).
 * pairs_sampled (nullable) receives the pairs processed (the reference's _total_pairs_sampled
 * increment: all links of the start node for type 0, the set size for type 1). */
int svi_fa2_run(svi_fa2 *h, uint32_t iter0, uint32_t iters, uint64_t philox_seed, uint64_t *pairs_sampled);
/* The minibatch svi_fa2_run would draw at iteration `iter` (diagnostics / tests): returns type,
 * start node and the pair list (up to `cap` pairs are copied; *npairs is the full count). */
int svi_fa2_draw(svi_fa2 *h, uint32_t iter, uint64_t philox_seed, uint32_t *type, uint32_t *start,
                 uint64_t *npairs, uint32_t *pairs, uint64_t cap);

/* FastAMM2::edge_likelihood (src/fastamm2.hh:477-520) of `npairs` pairs, incl. the 1e-30 floor. */
int svi_fa2_heldout(svi_fa2 *h, uint64_t npairs, const uint32_t *p, const uint32_t *q, const uint8_t *y,
                    double *loglik);

/* The phi pair of ONE (p,q,y) under the current gamma/lambda -- PhiCompute::update_phis_until_conv
 * (src/fastamm2.hh:151-209).  phi1/phi2: [k]; rounds (nullable): coordinate-ascent rounds executed.
 * Diagnostics and parity tests. */
int svi_fa2_phi_pair(svi_fa2 *h, uint32_t p, uint32_t q, int y, double *phi1, double *phi2, uint32_t *rounds);

typedef struct svi_fa2_info {
  uint32_t ld, lanes, vec;      /* padded row length; lanes per pair; 16-byte vectors per lane */
  uint32_t pair_blocks;         /* persistent grid of the pair kernel                          */
  uint64_t device_bytes;
  uint64_t last_npairs;         /* pairs of the last iteration                                 */
  uint64_t last_rounds;         /* coordinate-ascent rounds summed over those pairs            */
  uint32_t kernels_per_step;
} svi_fa2_info;
int svi_fa2_get_info(svi_fa2 *h, svi_fa2_info *info);

#ifdef __cplusplus
}
#endif
#endif /* SVI_FA2_H */
