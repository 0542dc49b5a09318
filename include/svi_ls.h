/* svi_ls.h -- C ABI of the B200 link-sampling engine (libsvi_ls.so).
 *
 * This is the drop-in boundary for ONE path of premgopalan/svinet: the body of
 * LinkSampling::infer's while(1) loop (reference src/linksampling.cc:571-789) and the
 * held-out likelihood it evaluates every report (src/linksampling.cc:966-1050,
 * src/linksampling.hh:259-292).  The reference has no FFI; its seam is the C++ class
 * `LinkSampling` used at src/main.cc:337-341.  A replacement `LinkSampling` keeps the
 * reference's host responsibilities (Env, Network, held-out draw, init_gamma2, file
 * writers, stop state machine) and drives the device through the calls below; see
 * INTEGRATION.md for the binding a maintainer adds on the reference side.
 *
 * Conventions: plain C types; caller-owned HOST buffers unless a name says `_dev`;
 * every call returns 0 on success or a negative svi_status, with a message available
 * from svi_ls_last_error().  A handle is bound to one CUDA device and one stream and
 * must be driven by one host thread at a time.  All arithmetic is FP64 (the reference
 * is FP64 throughout).
 */
#ifndef SVI_LS_H
#define SVI_LS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVI_LS_ABI_VERSION 2

typedef enum svi_status {
  SVI_OK = 0,
  SVI_ERR_INVALID = -1,   /* bad argument                                   */
  SVI_ERR_CUDA = -2,      /* a CUDA runtime call or kernel failed           */
  SVI_ERR_NOMEM = -3,     /* host or device allocation failed               */
  SVI_ERR_UNSUPPORTED = -4/* e.g. K > 65535 (the reference's uint16_t ids)   */
} svi_status;

typedef struct svi_ls svi_ls; /* opaque */

/* Static description of one inference problem.
 * Replaces the constructor state of LinkSampling (src/linksampling.cc:5-33) that the
 * sweep reads: _n, _k, env.alpha (src/env.hh:344), env.eta0/eta1 (src/network.cc:233-250),
 * _network.ones() (numerator of the annealing rescale, src/linksampling.cc:542). */
typedef struct svi_ls_config {
  uint32_t n;          /* inference nodes (env.n after src/main.cc:291)                    */
  uint32_t k;          /* communities                                                      */
  uint64_t nlinks;     /* training links, one (p<q) pair each                              */
  double   alpha;      /* Dirichlet prior, 1/k in the reference                            */
  double   eta0, eta1; /* Beta prior                                                       */
  uint32_t ones;       /* all links of the network incl. held-out ones                     */
  int32_t  device;     /* CUDA ordinal; -1 = the calling thread's current device           */
  uint32_t seg_len;    /* neighbours per work segment; 0 = choose automatically            */
  /* Node-block shard owned by this handle (multi-GPU).  A single
   * GPU run uses node_begin = 0, node_end = n.  The handle sweeps only half-edges whose
   * source node is in [node_begin, node_end) and refreshes only those rows; the caller
   * exchanges the row blocks and the K-vectors between phases (see svi_ls_phase_*). */
  uint32_t node_begin, node_end;
} svi_ls_config;

/* Build the device-side problem: CSR adjacency over the training links, work segments,
 * state matrices.  Replaces LinkSampling::assign_training_links' product `_links` /
 * `_training_links` (src/linksampling.cc:493-523).
 *   links : [2*nlinks] uint32, (p,q) with p<q, any order (the reference order is not
 *           needed: the device path is order-independent by construction)
 *   tl    : [n] the reference's _training_links (= 2 x training degree),
 *           or NULL to derive it from `links`. */
int svi_ls_create(const svi_ls_config *cfg, const uint32_t *links, const double *tl, svi_ls **out);
void svi_ls_destroy(svi_ls *h);

/* Launch on the caller's stream (a cudaStream_t passed as void*; NULL = legacy default). */
int svi_ls_set_stream(svi_ls *h, void *cuda_stream);
int svi_ls_sync(svi_ls *h);

/* Upload gamma [n*k row-major] and lambda [k*2] and derive the expectations the sweep
 * reads.  Replaces init_gamma2/init_lambda/load_model hand-off + set_dir_exp at
 * src/linksampling.cc:110-124 and :561-563.  Does not touch `converged`. */
int svi_ls_set_state(svi_ls *h, const double *gamma, const double *lambda);
/* Download gamma [n*k] / lambda [k*2] (either may be NULL).  For save_model /
 * write_groups / SIGTERM dumps (src/linksampling.cc:793-837, :1453-1476). */
int svi_ls_get_state(svi_ls *h, double *gamma, double *lambda);

/* _converged (src/linksampling.cc:456-475): 0 or community+1 per node.  create() zeroes it
 * (infer() does, :559); set is for resuming from a known state. */
int svi_ls_set_converged(svi_ls *h, const uint32_t *converged);
int svi_ls_get_converged(svi_ls *h, uint32_t *converged, uint32_t *active_comms);

/* One full iteration on the device == the loop body src/linksampling.cc:584-761:
 * phi sweep (dense :685-701, converged shortcut :619-631, active-set form :634-681 when
 * iter > 1000), compute_mean_indicators (:526-545), s3 sweep (:731-746), lambda finish
 * (:748-755), set_dir_exp x2 (:757-759), prune (:761).
 *   iter       : the reference's _iter (only its relation to 1000 matters)
 *   annealing  : _annealing_phase (:541)
 *   write_comm : when non-zero the link-community tally (:704-717) is rebuilt; read it
 *                back with svi_ls_get_membership.
 * Asynchronous on the handle's stream. */
int svi_ls_step(svi_ls *h, uint32_t iter, int annealing, int write_comm);

/* Link-community membership of the last write_comm sweep: bit c of word
 * bits[p*words + c/32] is set iff node p is an endpoint of a full-phi link whose arg-max
 * community is c (replaces _fmap/_communities, src/linksampling.cc:668-681,704-717, with
 * link_thresh = lt_min_deg = 0).  words = (k+31)/32. */
int svi_ls_get_membership(svi_ls *h, uint32_t *bits);

/* Held-out log-likelihood of `npairs` node pairs under the current gamma/lambda:
 * LinkSampling::edge_likelihood (src/linksampling.hh:259-292) incl. the 1e-30 floor;
 * y[i] = 1 for a link.  The caller sums in its own order (validation_likelihood,
 * src/linksampling.cc:966-1002). */
int svi_ls_heldout(svi_ls *h, uint64_t npairs, const uint32_t *p, const uint32_t *q,
                   const uint8_t *y, double epsilon, double *loglik);

/* The four K-vectors of the last sweep (_sum,_s1,_s2,_s3; src/linksampling.hh:157); any
 * pointer may be NULL.  Diagnostics and parity tests. */
int svi_ls_get_kvectors(svi_ls *h, double *sum, double *s1, double *s2, double *s3);

/* ---- phase-level entry points (multi-GPU drivers) -------------
 * svi_ls_step == phase_phi; phase_node; phase_s3; phase_finish run back to back.  A
 * sharded driver interleaves its collectives on the buffers returned by
 * svi_ls_device_buffer between the phases:
 *   phase_phi   : phi sweep over the shard's half-edges            -> per-segment partial rows
 *   phase_node  : mean indicators for the shard's rows             -> SVI_BUF_MPHI rows, local
 *                 column sums in SVI_BUF_KVEC (sum,s1,s2)            [all-reduce sum,s1,s2;
 *                                                                     all-gather mphi rows]
 *   phase_s3    : s3 over the links the shard's nodes own          -> SVI_BUF_KVEC (s3) [all-reduce]
 *   phase_finish: lambda, Elogbeta, gamma rescale, Elogpi, prune   -> SVI_BUF_EXPPI rows,
 *                 SVI_BUF_CONVERGED                                  [all-gather both] */
int svi_ls_phase_phi(svi_ls *h, uint32_t iter, int write_comm);
int svi_ls_phase_node(svi_ls *h);
int svi_ls_phase_s3(svi_ls *h);
int svi_ls_phase_finish(svi_ls *h, int annealing);
/* phase_finish in two halves, for drivers that hide an exchange behind the s3 sweep:
 *   phase_refresh : gamma rescale, Elogpi, prune for the shard's rows.  Needs only `sum` of SVI_BUF_KVEC, so it
 *                   may run right after phase_node's all-reduce, BEFORE phase_s3: `converged` is double-buffered,
 *                   prune writes the copy the NEXT iteration reads and the s3 sweep still sees this iteration's
 *                   flags (the reference's s3 loop, :731-746, precedes prune, :761).  SVI_BUF_CONVERGED names the
 *                   freshly pruned copy from here on (ask for the pointer again after every refresh).
 *                   -> SVI_BUF_EXPPI rows, SVI_BUF_CONVERGED  [all-gather both, may overlap phase_s3]
 *   phase_lambda  : lambda, Elogbeta (needs s3); ends the iteration.
 * Order: phase_phi; phase_node; phase_refresh; phase_s3; phase_lambda  ==  phase_phi; phase_node; phase_s3;
 * phase_finish. */
int svi_ls_phase_refresh(svi_ls *h, int annealing);
int svi_ls_phase_lambda(svi_ls *h, int annealing);

/* ---- multi-GPU over peer memory (the seam is still src/main.cc:337-341: one LinkSampling
 * object, now spread over the GPUs of one box) ------------------------------------------------------------------
 * Every shard (a handle created with its node block [node_begin, node_end) and the WHOLE link list, or at least all
 * links incident to its block) keeps its exchange buffers in one arena.  After svi_ls_peer_attach[_local] the shards
 * see each other's arenas (CUDA IPC between processes, peer access inside one) and svi_ls_mg_step runs the whole
 * iteration including its exchanges: a shard pushes the rows it produced (mphi after the node pass, exp(Elogpi) /
 * converged / active masks after the refresh) into every peer's arena with the copy engines on a side stream, beside
 * the sweeps, and announces them with epoch-stamped flags that the consumers wait on; the K-vector all-reduces (sum,
 * s1, s2; s3) are a push into per-source slots plus a fixed-order sum, bit-identical on every shard.  No collective
 * library on the data path.  All shards must call svi_ls_mg_step with the same arguments, once per iteration; the
 * call is asynchronous (svi_ls_sync reports an exchange that timed out).  Destroy the handles only after all shards
 * have finished (the caller's barrier).
 *   one process per GPU : svi_ls_peer_export -> exchange the blobs (e.g. torch.distributed all_gather) ->
 *                         svi_ls_peer_attach
 *   one process, N GPUs : svi_ls_peer_attach_local with the N handles (the C++ CLI's -gpus N)
 *   bounds : [world+1] node blocks of all shards (bounds[rank] .. bounds[rank+1] is this handle's)
 *   chunks : pipeline chunks of the shard's block (0 = default 4); the mphi rows of a finished chunk travel beside
 *            the next chunk's sweep
 * A shard works on four streams of its own (sweeps, owned-segment sweeps, node passes, pushes) and parks flag-wait
 * kernels at their heads: a process that hosts shards should start CUDA with CUDA_DEVICE_MAX_CONNECTIONS >= 16 (32
 * when several shards share one device) so that these streams never share a hardware queue. */
size_t svi_ls_peer_blob_bytes(void);
int svi_ls_peer_export(svi_ls *h, void *blob, size_t blob_bytes);
int svi_ls_peer_attach(svi_ls *h, uint32_t world, uint32_t rank, const uint32_t *bounds, const void *blobs,
                       uint32_t chunks);
int svi_ls_peer_attach_local(svi_ls *h, uint32_t world, uint32_t rank, const uint32_t *bounds,
                             svi_ls *const *handles, uint32_t chunks);
int svi_ls_mg_step(svi_ls *h, uint32_t iter, int annealing, int write_comm);
/* also push the refreshed gamma rows, so that every shard holds the whole gamma (svi_ls_heldout on any pair,
 * svi_ls_get_state of the whole matrix) */
int svi_ls_mg_share_gamma(svi_ls *h, int on);
/* Without replication: svi_ls_heldout (any pairs, on any shard, between two svi_ls_mg_step calls of ALL shards) reads
 * the rows of other shards straight from their arenas (peer loads), and
 * the whole gamma is assembled on demand -- every shard calls svi_ls_mg_publish_gamma (pushes its rows to all
 * peers, asynchronous), after which svi_ls_get_state on any shard returns the whole matrix. */
int svi_ls_mg_publish_gamma(svi_ls *h);
int svi_ls_mg_error(svi_ls *h);
/* Per-phase device times of svi_ls_mg_step, measured with events on the handle's stream: enable, run steps, then
 * read the mean over the (at most 32) last steps into phase_ms[10]:
 *   [0] wait for the peers' exp(Elogpi)/converged rows   (exposed exchange)
 *   [1] partition + phi sweep + mean indicators, chunked  [2] all-reduce sum,s1,s2
 *   [3] refresh                                           [4] wait for the peers' mphi rows (exposed exchange)
 *   [5] s3 sweep                                          [6] all-reduce s3 + lambda
 *   [7] wait for the own pushes to drain                  (exposed exchange)
 *   [8] side stream: first mphi push .. mphi flag raised  [9] side stream: first exp(Elogpi) push .. flag raised */
int svi_ls_mg_timing(svi_ls *h, int enable, double *phase_ms, uint32_t *steps);
/* membership words of the rows [first, first+count) only.  After svi_ls_mg_step a shard holds the complete words of
 * its OWN block (the bits other shards set for its nodes are merged in after the sweep); svi_ls_get_membership on a
 * shard returns its local replica, which is complete for those rows only. */
int svi_ls_get_membership_rows(svi_ls *h, uint32_t first, uint32_t count, uint32_t *bits);

typedef enum svi_buffer {
  SVI_BUF_EXPPI = 0,     /* double [n * ld]  exp(Elogpi - rowmax), rows padded to ld  */
  SVI_BUF_MPHI = 1,      /* double [n * ld]                                           */
  SVI_BUF_GAMMA = 2,     /* double [n * ld]                                           */
  SVI_BUF_KVEC = 3,      /* double [4 * ld]: sum, s1, s2, s3 (local to the shard)     */
  SVI_BUF_CONVERGED = 4, /* uint32 [n]                                                */
  SVI_BUF_LAMBDA = 5,    /* double [k * 2]                                            */
  SVI_BUF_ACTIVE = 6,    /* uint32 [n]  active_comms (read by the iter > 1000 branch) */
  SVI_BUF_ACTIVE_BITS = 7,/* uint32 [n * words] active-community mask, same branch    */
  SVI_BUF_MEMBER_BITS = 8/* uint32 [n * words] link-community membership              */
} svi_buffer;
/* Device pointer + leading dimension (in elements) of an exchange buffer. */
int svi_ls_device_buffer(svi_ls *h, svi_buffer which, void **dev_ptr, uint64_t *ld);

/* Work counters of the problem: half-edges swept by phase_phi / phase_s3, segments. */
typedef struct svi_ls_info {
  uint64_t half_edges_phi, half_edges_s3, segments_phi, segments_s3;
  uint32_t ld;            /* padded row length (elements)                              */
  uint32_t seg_len;
  uint32_t lanes, vec;    /* sweep-kernel tiling selected for this K (lanes per segment,  */
                          /* 16-byte vectors per lane)                                    */
  uint32_t ring_depth;    /* rows in flight per group in the TMA ring sweeps (0 = unused) */
  uint64_t device_bytes;  /* HBM allocated by the handle                               */
  uint32_t kernels_per_step;
} svi_ls_info;
int svi_ls_get_info(svi_ls *h, svi_ls_info *info);

const char *svi_ls_last_error(void);
int svi_ls_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SVI_LS_H */
