// svi_fa2.cu -- C ABI (include/svi_fa2.h) over the kernels in svi_fa2_kernels.cuh: the
// `-rnode -stratified` iteration (reference class FastAMM2).  Buffer management, dispatch by K,
// stream plumbing; no CPU compute fallback.
#include "../../include/svi_fa2.h"
#include "svi_common.h"
#include "svi_fa2_kernels.cuh"
#include "svi_fa2_wide.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

using svi::fail;
using svi::Fa2Ctrl;
using svi::Fa2Params;

namespace {

struct Fa2Ops {
  void (*prep)(const Fa2Params &, cudaStream_t);
  void (*pairs)(const Fa2Params &, cudaStream_t);
  void (*blend)(const Fa2Params &, cudaStream_t);
  void (*heldout)(const Fa2Params &, cudaStream_t, uint64_t, const uint32_t *, const uint32_t *, const uint8_t *,
                  double *);
  void (*one_pair)(const Fa2Params &, cudaStream_t, uint32_t, uint32_t, int, double *, uint32_t *);
  int (*pair_blocks)(int sms);
  int lanes, vec, cap;
  int groups_per_block;      // pairs a block of the pair kernel works on at a time (k_fa2_lambda: which blocks had pairs)
  int wide_rows;             // K > 512: scratch rows (of ld doubles) per pair block, else 0
};

template <int G, int V>
struct Fa2Tile {
  static constexpr int T = 128;
  static constexpr int CAP = 2 * G * V;
  static constexpr size_t kSmemPairs = (size_t)2 * (T / G) * CAP * sizeof(double);
  static void prep(const Fa2Params &P, cudaStream_t st) { svi::k_fa2_prep<G, V><<<1, 128, 0, st>>>(P); }
  static void pairs(const Fa2Params &P, cudaStream_t st) {
    svi::k_fa2_pairs<G, V, T><<<P.pair_blocks, T, kSmemPairs, st>>>(P);
  }
  static void blend(const Fa2Params &P, cudaStream_t st) {
    // 4 rows in flight per group (k_fa2_blend): a grid that covers the rows once, capped at 32 blocks per SM
    const uint64_t need = ((uint64_t)P.n * G + 4 * T - 1) / (4 * T);
    const uint32_t blocks = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(need, 148ull * 32));
    svi::k_fa2_blend<G, V, T><<<blocks, T, 0, st>>>(P);
  }
  static void heldout(const Fa2Params &P, cudaStream_t st, uint64_t np, const uint32_t *p, const uint32_t *q,
                      const uint8_t *y, double *out) {
    if (!np) return;
    const uint32_t blocks = (uint32_t)((np * G + 127) / 128);
    svi::k_fa2_heldout<G, V><<<blocks, 128, 0, st>>>(P, np, p, q, y, out);
  }
  static void one_pair(const Fa2Params &P, cudaStream_t st, uint32_t p, uint32_t q, int y, double *phi,
                       uint32_t *rounds) {
    svi::k_fa2_one_pair<G, V><<<1, 32, 0, st>>>(P, p, q, y, phi, rounds);
  }
  static int pair_blocks(int sms) {
    auto kern = svi::k_fa2_pairs<G, V, T>;
    if (kSmemPairs > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPairs);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, kSmemPairs) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    return per_sm * sms;
  }
  static Fa2Ops ops() { return Fa2Ops{prep, pairs, blend, heldout, one_pair, pair_blocks, G, V, CAP, T / G, 0}; }
};

// K > 512: one block per pair / row, columns strided over its threads (svi_fa2_wide.cuh).  The slot width of the block
// partials depends on K, so these launchers read it from the handle's parameters.
struct Fa2WideTile {
  static constexpr int T = (int)svi::kWideT;
  static uint32_t cap_of(const Fa2Params &P) { return svi::wide_cap(P.ld); }
  static void prep(const Fa2Params &P, cudaStream_t st) { svi::k_fa2_prep_wide<<<1, T, 0, st>>>(P); }
  static void pairs(const Fa2Params &P, cudaStream_t st) { svi::k_fa2_pairs_wide<<<P.pair_blocks, T, 0, st>>>(P, cap_of(P)); }
  static void blend(const Fa2Params &P, cudaStream_t st) {
    svi::k_fa2_blend_wide<<<std::max<uint32_t>(1, std::min<uint32_t>(P.n, 148u * 8)), T, 0, st>>>(P, cap_of(P));
  }
  static void heldout(const Fa2Params &P, cudaStream_t st, uint64_t np, const uint32_t *p, const uint32_t *q,
                      const uint8_t *y, double *out) {
    if (!np) return;
    svi::k_fa2_heldout_wide<<<(uint32_t)std::min<uint64_t>(np, 1u << 20), T, 0, st>>>(P, np, p, q, y, out);
  }
  static void one_pair(const Fa2Params &P, cudaStream_t st, uint32_t p, uint32_t q, int y, double *phi,
                       uint32_t *rounds) {
    // its six scratch rows follow the pair blocks' (svi_fa2_create)
    double *rows = P.wide + (size_t)P.pair_blocks * svi::kFa2WideRows * P.ld;
    svi::k_fa2_one_pair_wide<<<1, T, 0, st>>>(P, rows, p, q, y, phi, rounds);
  }
  static int pair_blocks(int sms) { return 2 * sms; }   // (each block keeps 5 K-rows of scratch + 2 partial rows)
  static Fa2Ops ops(uint32_t k) {
    const uint32_t ld = (k + 3u) & ~3u, cap = svi::wide_cap(ld);
    return Fa2Ops{prep, pairs, blend, heldout, one_pair, pair_blocks, T, (int)(cap / (2u * T)), (int)cap, 1,
                  (int)svi::kFa2WideRows};
  }
};

bool pick_fa2(uint32_t k, Fa2Ops *o) {
  if (k == 0) return false;
  if (k <= 4) *o = Fa2Tile<2, 1>::ops();
  else if (k <= 8) *o = Fa2Tile<4, 1>::ops();
  else if (k <= 16) *o = Fa2Tile<8, 1>::ops();
  else if (k <= 32) *o = Fa2Tile<16, 1>::ops();
  else if (k <= 64) *o = Fa2Tile<32, 1>::ops();
  else if (k <= 128) *o = Fa2Tile<32, 2>::ops();
  else if (k <= 192) *o = Fa2Tile<32, 3>::ops();
  else if (k <= 256) *o = Fa2Tile<32, 4>::ops();
  else if (k <= 384) *o = Fa2Tile<32, 6>::ops();
  else if (k <= 512) *o = Fa2Tile<32, 8>::ops();
  else if (k <= 65535) *o = Fa2WideTile::ops(k);   // the reference's limit: communities are uint16_t (src/env.hh:37)
  else return false;
  return true;
}

struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DevGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

}  // namespace

struct svi_fa2 {
  svi_fa2_config cfg{};
  int device = 0, sms = 0;
  cudaStream_t stream = nullptr;
  Fa2Ops ops{};
  Fa2Params P{};
  uint64_t nodec = 0;          // host mirror of _nodec
  double log_c = 0.0;          // host mirror of log(cscale) (lazy mode): decides when to re-base
  uint64_t device_bytes = 0;
  bool have_graph = false;
  // owned device memory
  double *d_gamma = nullptr, *d_lambda = nullptr, *d_elogbeta = nullptr, *d_elogf = nullptr, *d_epi = nullptr;
  double *d_partS = nullptr, *d_partL = nullptr, *d_stage = nullptr, *d_wide = nullptr;
  uint32_t *d_pairs = nullptr, *d_shuffled = nullptr, *d_adj = nullptr;
  uint64_t *d_adj_off = nullptr, *d_heldout = nullptr;
  uint8_t *d_touched = nullptr;
  uint32_t *d_ballots = nullptr, *d_counts = nullptr;   // draw window: 32 ballots + 1 count per block
  uint32_t draw_blocks = 0;
  Fa2Ctrl *d_ctrl = nullptr;
  size_t stage_bytes = 0;
  Fa2Ctrl *h_ctrl = nullptr;   // pinned staging for step()
  uint32_t *h_pairs = nullptr; // pinned staging for the pair list
  size_t h_pairs_cap = 0;
  // step(): TWO sets of pinned staging (pair list + control block) used alternately, each guarded by an event
  // recorded after its upload, so the host stages minibatch i+1 while the device still runs minibatch i
  Fa2Ctrl *st_ctrl[2] = {nullptr, nullptr};
  uint32_t *st_pairs[2] = {nullptr, nullptr};
  size_t st_cap[2] = {0, 0};
  cudaEvent_t st_done[2] = {nullptr, nullptr};
  int st_slot = 0;
};

namespace {

template <class T>
cudaError_t dalloc(svi_fa2 *h, T **p, size_t count) {
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void **)p, bytes);
  if (e == cudaSuccess) {
    h->device_bytes += bytes;
    e = cudaMemset(*p, 0, bytes);
  }
  return e;
}

int ensure_stage(svi_fa2 *h, size_t bytes) {
  if (h->stage_bytes >= bytes) return SVI_OK;
  if (h->d_stage) cudaFree(h->d_stage);
  h->d_stage = nullptr;
  h->stage_bytes = 0;
  SVI_CK(cudaMalloc((void **)&h->d_stage, std::max<size_t>(bytes, 8)));
  h->stage_bytes = bytes;
  return SVI_OK;
}

int ensure_pairs(svi_fa2 *h, uint64_t npairs) {
  if (npairs <= h->P.cap_pairs) return SVI_OK;
  const uint64_t cap = std::max<uint64_t>(npairs, 2 * (uint64_t)h->P.cap_pairs);
  if (cap > 0x7fffffffull) return fail(SVI_ERR_UNSUPPORTED, "svi_fa2: %llu pairs in one minibatch", (unsigned long long)npairs);
  SVI_CK(cudaStreamSynchronize(h->stream));
  if (h->d_pairs) cudaFree(h->d_pairs);
  h->d_pairs = nullptr;
  SVI_CK(cudaMalloc((void **)&h->d_pairs, cap * 2 * sizeof(uint32_t)));
  h->P.pairs = h->d_pairs;
  h->P.cap_pairs = (uint32_t)cap;
  return SVI_OK;
}

// the three launches of one device-side minibatch draw (svi_fa2_kernels.cuh)
void launch_draw(svi_fa2 *h, uint32_t iter, uint64_t seed) {
  const uint32_t lo = (uint32_t)seed, hi = (uint32_t)(seed >> 32);
  svi::k_fa2_draw_count<<<h->draw_blocks, 1024, 0, h->stream>>>(h->P, iter, lo, hi, h->d_ballots, h->d_counts);
  svi::k_fa2_draw_emit<<<h->draw_blocks, 1024, 0, h->stream>>>(h->P, iter, lo, hi, h->d_ballots, h->d_counts, h->draw_blocks);
  svi::k_fa2_draw_tail<<<1, 1024, 0, h->stream>>>(h->P, iter, lo, hi, h->draw_blocks);
}

// prep, pairs, [blend], lambda; `nodec` = the node counter this iteration runs with
void launch_iteration(svi_fa2 *h, uint64_t nodec) {
  h->ops.prep(h->P, h->stream);
  h->ops.pairs(h->P, h->stream);
  if (!h->P.lazy) h->ops.blend(h->P, h->stream);
  svi::k_fa2_lambda<<<1, 256, 0, h->stream>>>(h->P, (uint32_t)h->ops.cap, (uint32_t)h->ops.groups_per_block);
  if (h->P.lazy) {
    // c shrinks by (1 - rho) per iteration: exp(-2 sqrt(T)) in the long run.  Re-base the stored rows long
    // before c*u leaves the FP64 range (every ~2e4 iterations at the default step sizes).
    const double rho = std::pow(h->cfg.nodetau0 + (double)nodec, -1 * h->cfg.nodekappa);
    h->log_c += std::log1p(-rho);
    if (h->log_c < -230.0) {
      svi::k_fa2_fold<<<h->sms * 8, 256, 0, h->stream>>>(h->P);
      svi::k_fa2_reset_scale<<<1, 1, 0, h->stream>>>(h->P);
      h->log_c = 0.0;
    }
  }
}

}  // namespace

extern "C" {

void svi_fa2_default_config(svi_fa2_config *c, uint32_t n, uint32_t k) {
  if (!c) return;
  memset(c, 0, sizeof *c);
  c->n = n; c->k = k;
  c->alpha = k ? 1.0 / k : 0.0;
  c->eta0 = 1.0; c->eta1 = 1.0; c->epsilon = 1e-30;
  c->tau0 = 1025.0; c->kappa = 0.9; c->nodetau0 = 1025.0; c->nodekappa = 0.5;
  c->inf_epsilon = 0.5; c->m_sets = 10; c->online_iterations = 50; c->meanchangethresh = 1e-5;
  c->nolambda = 0; c->device = -1; c->eager_blend = 0;
}

int svi_fa2_create(const svi_fa2_config *cfg, svi_fa2 **out) {
  if (!cfg || !out) return fail(SVI_ERR_INVALID, "svi_fa2_create: null argument");
  *out = nullptr;
  if (cfg->n < 2 || cfg->k == 0) return fail(SVI_ERR_INVALID, "svi_fa2_create: need n >= 2 and k >= 1");
  if (!(cfg->epsilon > 0.0) || cfg->m_sets == 0) return fail(SVI_ERR_INVALID, "svi_fa2_create: bad epsilon / m_sets");
  Fa2Ops ops;
  if (!pick_fa2(cfg->k, &ops)) return fail(SVI_ERR_UNSUPPORTED, "svi_fa2_create: k=%u not supported (max 65535)", cfg->k);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(SVI_ERR_CUDA, "svi_fa2_create: no CUDA device");
  int dev = cfg->device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return fail(SVI_ERR_CUDA, "svi_fa2_create: cudaGetDevice failed");
  if (dev >= ndev) return fail(SVI_ERR_INVALID, "svi_fa2_create: device %d of %d", dev, ndev);
  svi_fa2 *h = new (std::nothrow) svi_fa2();
  if (!h) return fail(SVI_ERR_NOMEM, "svi_fa2_create: host allocation failed");
  h->cfg = *cfg;
  h->device = dev;
  h->ops = ops;
  DevGuard guard(dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    delete h;
    return fail(SVI_ERR_CUDA, "svi_fa2_create: cudaGetDeviceProperties failed");
  }
  h->sms = prop.multiProcessorCount;
  Fa2Params &P = h->P;
  P.n = cfg->n; P.k = cfg->k; P.ld = (cfg->k + 3u) & ~3u;
  P.alpha = cfg->alpha; P.eta0 = cfg->eta0; P.eta1 = cfg->eta1;
  P.epsilon = cfg->epsilon; P.logeps = std::log(cfg->epsilon); P.thresh = cfg->meanchangethresh;
  P.tau0 = cfg->tau0; P.kappa = cfg->kappa; P.nodetau0 = cfg->nodetau0; P.nodekappa = cfg->nodekappa;
  P.inf_epsilon = cfg->inf_epsilon; P.online_iters = cfg->online_iterations; P.m_sets = cfg->m_sets;
  P.nolambda = cfg->nolambda ? 1u : 0u;
  P.lazy = cfg->eager_blend ? 0u : 1u;
  P.pair_blocks = (uint32_t)ops.pair_blocks(h->sms);
  const uint32_t setsize = (uint32_t)((double)cfg->n / (double)cfg->m_sets);
  P.cap_pairs = std::max<uint32_t>(setsize, 1024);
  cudaError_t e = cudaSuccess;
  auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  A(dalloc(h, &h->d_gamma, (size_t)P.n * P.ld));
  A(dalloc(h, &h->d_lambda, 2 * (size_t)P.k));
  A(dalloc(h, &h->d_elogbeta, 2 * (size_t)P.ld));
  A(dalloc(h, &h->d_elogf, P.ld));
  A(dalloc(h, &h->d_epi, P.ld));
  A(dalloc(h, &h->d_partS, (size_t)P.pair_blocks * ops.cap));
  A(dalloc(h, &h->d_partL, (size_t)P.pair_blocks * ops.cap));
  if (ops.wide_rows)
    A(dalloc(h, &h->d_wide, ((size_t)P.pair_blocks * ops.wide_rows + svi::kFa2WideOneRows) * P.ld));
  A(dalloc(h, &h->d_pairs, 2 * (size_t)P.cap_pairs));
  A(dalloc(h, &h->d_touched, P.n));
  A(dalloc(h, &h->d_ctrl, 1));
  A(cudaMallocHost((void **)&h->h_ctrl, sizeof(Fa2Ctrl)));
  if (e != cudaSuccess) {
    const int rc = fail(e == cudaErrorMemoryAllocation ? SVI_ERR_NOMEM : SVI_ERR_CUDA, "svi_fa2_create: %s",
                        cudaGetErrorString(e));
    svi_fa2_destroy(h);
    return rc;
  }
  P.gamma = h->d_gamma; P.lambda = h->d_lambda; P.elogbeta = h->d_elogbeta; P.elogf = h->d_elogf;
  P.epi_start = h->d_epi; P.partS = h->d_partS; P.partL = h->d_partL; P.pairs = h->d_pairs;
  P.touched = h->d_touched; P.ctrl = h->d_ctrl; P.wide = h->d_wide;
  *out = h;
  return SVI_OK;
}

void svi_fa2_destroy(svi_fa2 *h) {
  if (!h) return;
  DevGuard guard(h->device);
  cudaStreamSynchronize(h->stream);
  void *ptrs[] = {h->d_gamma, h->d_lambda, h->d_elogbeta, h->d_elogf, h->d_epi, h->d_partS, h->d_partL, h->d_stage, h->d_wide,
                  h->d_pairs, h->d_shuffled, h->d_adj, h->d_adj_off, h->d_heldout, h->d_touched, h->d_ctrl,
                  h->d_ballots, h->d_counts};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (h->h_ctrl) cudaFreeHost(h->h_ctrl);
  if (h->h_pairs) cudaFreeHost(h->h_pairs);
  for (int i = 0; i < 2; ++i) {
    if (h->st_ctrl[i]) cudaFreeHost(h->st_ctrl[i]);
    if (h->st_pairs[i]) cudaFreeHost(h->st_pairs[i]);
    if (h->st_done[i]) cudaEventDestroy(h->st_done[i]);
  }
  delete h;
}

int svi_fa2_set_stream(svi_fa2 *h, void *cuda_stream) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  h->stream = (cudaStream_t)cuda_stream;
  return SVI_OK;
}

int svi_fa2_sync(svi_fa2 *h) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DevGuard guard(h->device);
  SVI_CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_fa2_set_state(svi_fa2 *h, const double *gamma, const double *lambda, uint64_t nodec) {
  if (!h || !gamma || !lambda) return fail(SVI_ERR_INVALID, "svi_fa2_set_state: null argument");
  DevGuard guard(h->device);
  const Fa2Params &P = h->P;
  const size_t nk = (size_t)P.n * P.k;
  {
    int rc = ensure_stage(h, nk * sizeof(double));
    if (rc) return rc;
    SVI_CK(cudaMemcpyAsync(h->d_stage, gamma, nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    // lazy: rows store u = gamma - alpha with c = 1
    svi::k_fa2_import<<<h->sms * 8, 256, 0, h->stream>>>(h->d_stage, h->d_gamma, P.n, P.k, P.ld, P.lazy ? P.alpha : 0.0);
  }
  SVI_CK(cudaMemcpyAsync(h->d_lambda, lambda, 2 * (size_t)P.k * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  memset(h->h_ctrl, 0, sizeof(Fa2Ctrl));
  h->h_ctrl->nodec = (double)nodec;
  h->h_ctrl->cscale = 1.0;
  h->log_c = 0.0;
  SVI_CK(cudaMemcpyAsync(h->d_ctrl, h->h_ctrl, sizeof(Fa2Ctrl), cudaMemcpyHostToDevice, h->stream));
  SVI_CK(cudaMemsetAsync(h->d_touched, 0, P.n, h->stream));
  h->nodec = nodec;
  SVI_CK(cudaGetLastError());
  SVI_CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_fa2_get_state(svi_fa2 *h, double *gamma, double *lambda) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DevGuard guard(h->device);
  const Fa2Params &P = h->P;
  const size_t nk = (size_t)P.n * P.k;
  if (gamma) {
    int rc = ensure_stage(h, nk * sizeof(double));
    if (rc) return rc;
    svi::k_fa2_export<<<h->sms * 8, 256, 0, h->stream>>>(P, h->d_stage);
    SVI_CK(cudaMemcpyAsync(gamma, h->d_stage, nk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  if (lambda)
    SVI_CK(cudaMemcpyAsync(lambda, h->d_lambda, 2 * (size_t)P.k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  SVI_CK(cudaGetLastError());
  SVI_CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_fa2_step(svi_fa2 *h, uint32_t iter, uint32_t type, uint32_t start, uint64_t npairs, const uint32_t *pairs) {
  if (!h || (npairs && !pairs)) return fail(SVI_ERR_INVALID, "svi_fa2_step: null argument");
  if (type > 1 || start >= h->P.n) return fail(SVI_ERR_INVALID, "svi_fa2_step: bad type %u / start %u", type, start);
  for (uint64_t i = 0; i < npairs; ++i) {
    const uint32_t p = pairs[2 * i], q = pairs[2 * i + 1];
    if (p >= h->P.n || q >= h->P.n || p == q || (p != start && q != start))
      return fail(SVI_ERR_INVALID, "svi_fa2_step: pair %llu = (%u,%u) does not contain start node %u",
                  (unsigned long long)i, p, q, start);
  }
  DevGuard guard(h->device);
  int rc = ensure_pairs(h, npairs);
  if (rc) return rc;
  // staging slot: wait only for the upload that used this slot two steps ago (not for the device to go idle)
  const int slot = h->st_slot ^= 1;
  if (!h->st_done[slot]) {
    SVI_CK(cudaEventCreateWithFlags(&h->st_done[slot], cudaEventDisableTiming));
    SVI_CK(cudaMallocHost((void **)&h->st_ctrl[slot], sizeof(Fa2Ctrl)));
  } else {
    SVI_CK(cudaEventSynchronize(h->st_done[slot]));
  }
  if (npairs > h->st_cap[slot]) {
    if (h->st_pairs[slot]) cudaFreeHost(h->st_pairs[slot]);
    h->st_pairs[slot] = nullptr;
    h->st_cap[slot] = 0;
    const size_t cap = std::max<uint64_t>(npairs + npairs / 4, 1024);
    SVI_CK(cudaMallocHost((void **)&h->st_pairs[slot], cap * 2 * sizeof(uint32_t)));
    h->st_cap[slot] = cap;
  }
  if (npairs) {
    memcpy(h->st_pairs[slot], pairs, npairs * 2 * sizeof(uint32_t));
    SVI_CK(cudaMemcpyAsync(h->d_pairs, h->st_pairs[slot], npairs * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  }
  const svi_fa2_config &c = h->cfg;
  // only the per-iteration half of the control block is written: the device keeps the counters
  Fa2Ctrl *hc = h->st_ctrl[slot];
  hc->type = type; hc->start = start; hc->npairs = (uint32_t)npairs; hc->iter = iter;
  hc->sampled_inc = npairs;
  hc->rho_node = std::pow(c.nodetau0 + (double)h->nodec, -1 * c.nodekappa);             // src/fastamm2.cc:606
  hc->rho_t = std::pow(c.tau0 + ((double)iter + 1.0), -1 * c.kappa);                    // :627
  hc->scale = type == 0 ? (double)c.n / (2 * (1 - c.inf_epsilon))                       // :591-592
                        : ((double)c.n * (double)c.m_sets) / (2 * c.inf_epsilon);
  SVI_CK(cudaMemcpyAsync(h->d_ctrl, hc, offsetof(Fa2Ctrl, nodec), cudaMemcpyHostToDevice, h->stream));
  SVI_CK(cudaEventRecord(h->st_done[slot], h->stream));
  launch_iteration(h, h->nodec);
  h->nodec++;
  SVI_CK(cudaGetLastError());
  return SVI_OK;
}

int svi_fa2_set_graph(svi_fa2 *h, uint64_t nlinks, const uint32_t *links, uint64_t nheldout, const uint32_t *heldout,
                      const uint32_t *shuffled) {
  if (!h || (nlinks && !links) || (nheldout && !heldout) || !shuffled)
    return fail(SVI_ERR_INVALID, "svi_fa2_set_graph: null argument");
  const uint32_t n = h->P.n;
  std::vector<uint64_t> off((size_t)n + 1, 0);
  for (uint64_t e = 0; e < nlinks; ++e) {
    const uint32_t p = links[2 * e], q = links[2 * e + 1];
    if (p >= n || q >= n || p == q) return fail(SVI_ERR_INVALID, "svi_fa2_set_graph: link %llu = (%u,%u)", (unsigned long long)e, p, q);
    off[p + 1]++; off[q + 1]++;
  }
  uint64_t maxdeg = 0;
  for (uint32_t v = 0; v < n; ++v) { maxdeg = std::max(maxdeg, off[v + 1]); off[v + 1] += off[v]; }
  std::vector<uint32_t> adj(std::max<uint64_t>(2 * nlinks, 1));
  {
    std::vector<uint64_t> at(off.begin(), off.end() - 1);
    for (uint64_t e = 0; e < nlinks; ++e) {
      const uint32_t p = links[2 * e], q = links[2 * e + 1];
      adj[at[p]++] = q; adj[at[q]++] = p;
    }
  }
  for (uint32_t v = 0; v < n; ++v) std::sort(adj.begin() + off[v], adj.begin() + off[v + 1]);
  std::vector<uint64_t> ho(std::max<uint64_t>(nheldout, 1));
  for (uint64_t i = 0; i < nheldout; ++i) {
    const uint32_t a = std::min(heldout[2 * i], heldout[2 * i + 1]), b = std::max(heldout[2 * i], heldout[2 * i + 1]);
    ho[i] = ((uint64_t)a << 32) | b;
  }
  std::sort(ho.begin(), ho.begin() + nheldout);
  for (uint32_t i = 0; i < n; ++i)
    if (shuffled[i] >= n) return fail(SVI_ERR_INVALID, "svi_fa2_set_graph: shuffled[%u] = %u", i, shuffled[i]);
  DevGuard guard(h->device);
  SVI_CK(cudaStreamSynchronize(h->stream));
  for (void **p : {(void **)&h->d_adj_off, (void **)&h->d_adj, (void **)&h->d_heldout, (void **)&h->d_shuffled})
    if (*p) { cudaFree(*p); *p = nullptr; }
  SVI_CK(dalloc(h, &h->d_adj_off, (size_t)n + 1));
  SVI_CK(dalloc(h, &h->d_adj, adj.size()));
  SVI_CK(dalloc(h, &h->d_heldout, ho.size()));
  SVI_CK(dalloc(h, &h->d_shuffled, n));
  SVI_CK(cudaMemcpy(h->d_adj_off, off.data(), ((size_t)n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
  SVI_CK(cudaMemcpy(h->d_adj, adj.data(), adj.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  SVI_CK(cudaMemcpy(h->d_heldout, ho.data(), ho.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
  SVI_CK(cudaMemcpy(h->d_shuffled, shuffled, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  h->P.adj_off = h->d_adj_off; h->P.adj = h->d_adj; h->P.heldout = h->d_heldout; h->P.nheldout = nheldout;
  h->P.shuffled = h->d_shuffled;
  const uint32_t setsize = (uint32_t)((double)n / (double)h->cfg.m_sets);
  int rc = ensure_pairs(h, std::max<uint64_t>(maxdeg, setsize));
  if (rc) return rc;
  // draw window: the set size plus everything a typical start node excludes (its links, itself, held-out
  // partners) plus slack; a start node that excludes more is finished by k_fa2_draw_tail
  uint64_t avgdeg = n ? 2 * nlinks / n : 0;
  uint64_t window = std::max<uint64_t>(setsize + 4 * avgdeg + 2048, std::min<uint64_t>(maxdeg, 1u << 20));
  window = std::min<uint64_t>(window, std::max<uint64_t>(n, maxdeg));
  h->draw_blocks = (uint32_t)((window + 1023) / 1024);
  for (void **p : {(void **)&h->d_ballots, (void **)&h->d_counts})
    if (*p) { cudaFree(*p); *p = nullptr; }
  SVI_CK(dalloc(h, &h->d_ballots, (size_t)h->draw_blocks * 32));
  SVI_CK(dalloc(h, &h->d_counts, h->draw_blocks));
  h->have_graph = true;
  return SVI_OK;
}

int svi_fa2_run(svi_fa2 *h, uint32_t iter0, uint32_t iters, uint64_t seed, uint64_t *pairs_sampled) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  if (!h->have_graph) return fail(SVI_ERR_INVALID, "svi_fa2_run: call svi_fa2_set_graph first");
  DevGuard guard(h->device);
  uint64_t before = 0;
  if (pairs_sampled) {
    SVI_CK(cudaMemcpyAsync(h->h_ctrl, h->d_ctrl, sizeof(Fa2Ctrl), cudaMemcpyDeviceToHost, h->stream));
    SVI_CK(cudaStreamSynchronize(h->stream));
    before = h->h_ctrl->total_sampled;
  }
  for (uint32_t i = 0; i < iters; ++i) {
    launch_draw(h, iter0 + i, seed);
    launch_iteration(h, h->nodec + i);
  }
  h->nodec += iters;
  SVI_CK(cudaGetLastError());
  if (pairs_sampled) {
    SVI_CK(cudaMemcpyAsync(h->h_ctrl, h->d_ctrl, sizeof(Fa2Ctrl), cudaMemcpyDeviceToHost, h->stream));
    SVI_CK(cudaStreamSynchronize(h->stream));
    *pairs_sampled = h->h_ctrl->total_sampled - before;
  }
  return SVI_OK;
}

int svi_fa2_draw(svi_fa2 *h, uint32_t iter, uint64_t seed, uint32_t *type, uint32_t *start, uint64_t *npairs,
                 uint32_t *pairs, uint64_t cap) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  if (!h->have_graph) return fail(SVI_ERR_INVALID, "svi_fa2_draw: call svi_fa2_set_graph first");
  DevGuard guard(h->device);
  launch_draw(h, iter, seed);
  SVI_CK(cudaGetLastError());
  SVI_CK(cudaMemcpyAsync(h->h_ctrl, h->d_ctrl, sizeof(Fa2Ctrl), cudaMemcpyDeviceToHost, h->stream));
  SVI_CK(cudaStreamSynchronize(h->stream));
  if (type) *type = h->h_ctrl->type;
  if (start) *start = h->h_ctrl->start;
  if (npairs) *npairs = h->h_ctrl->npairs;
  if (pairs && cap) {
    const uint64_t ncopy = std::min<uint64_t>(cap, h->h_ctrl->npairs);
    if (ncopy) SVI_CK(cudaMemcpy(pairs, h->d_pairs, ncopy * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  return SVI_OK;
}

int svi_fa2_heldout(svi_fa2 *h, uint64_t npairs, const uint32_t *p, const uint32_t *q, const uint8_t *y, double *loglik) {
  if (!h || (npairs && (!p || !q || !y || !loglik))) return fail(SVI_ERR_INVALID, "svi_fa2_heldout: null argument");
  if (!npairs) return SVI_OK;
  for (uint64_t i = 0; i < npairs; ++i)
    if (p[i] >= h->P.n || q[i] >= h->P.n)
      return fail(SVI_ERR_INVALID, "svi_fa2_heldout: pair %llu out of range", (unsigned long long)i);
  DevGuard guard(h->device);
  const size_t bytes = npairs * (2 * sizeof(uint32_t) + sizeof(double) + 8);
  int rc = ensure_stage(h, bytes);
  if (rc) return rc;
  char *base = (char *)h->d_stage;
  double *d_out = (double *)base;
  uint32_t *d_p = (uint32_t *)(base + npairs * sizeof(double));
  uint32_t *d_q = d_p + npairs;
  uint8_t *d_y = (uint8_t *)(d_q + npairs);
  SVI_CK(cudaMemcpyAsync(d_p, p, npairs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  SVI_CK(cudaMemcpyAsync(d_q, q, npairs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  SVI_CK(cudaMemcpyAsync(d_y, y, npairs, cudaMemcpyHostToDevice, h->stream));
  h->ops.heldout(h->P, h->stream, npairs, d_p, d_q, d_y, d_out);
  SVI_CK(cudaGetLastError());
  SVI_CK(cudaMemcpyAsync(loglik, d_out, npairs * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  SVI_CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_fa2_phi_pair(svi_fa2 *h, uint32_t p, uint32_t q, int y, double *phi1, double *phi2, uint32_t *rounds) {
  if (!h || !phi1 || !phi2) return fail(SVI_ERR_INVALID, "svi_fa2_phi_pair: null argument");
  if (p >= h->P.n || q >= h->P.n || p == q) return fail(SVI_ERR_INVALID, "svi_fa2_phi_pair: bad pair (%u,%u)", p, q);
  DevGuard guard(h->device);
  const uint32_t k = h->P.k;
  int rc = ensure_stage(h, (2 * (size_t)k + 2) * sizeof(double));
  if (rc) return rc;
  h->ops.prep(h->P, h->stream);    // refreshes Elogbeta from the current lambda (scratch only)
  uint32_t *d_rounds = (uint32_t *)(h->d_stage + 2 * (size_t)k);
  h->ops.one_pair(h->P, h->stream, p, q, y ? 1 : 0, h->d_stage, d_rounds);
  SVI_CK(cudaGetLastError());
  SVI_CK(cudaMemcpyAsync(phi1, h->d_stage, k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  SVI_CK(cudaMemcpyAsync(phi2, h->d_stage + k, k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  uint32_t r = 0;
  SVI_CK(cudaMemcpyAsync(&r, d_rounds, sizeof r, cudaMemcpyDeviceToHost, h->stream));
  SVI_CK(cudaStreamSynchronize(h->stream));
  if (rounds) *rounds = r;
  return SVI_OK;
}

int svi_fa2_get_info(svi_fa2 *h, svi_fa2_info *info) {
  if (!h || !info) return fail(SVI_ERR_INVALID, "svi_fa2_get_info: null argument");
  DevGuard guard(h->device);
  Fa2Ctrl c;
  SVI_CK(cudaMemcpyAsync(&c, h->d_ctrl, sizeof c, cudaMemcpyDeviceToHost, h->stream));
  SVI_CK(cudaStreamSynchronize(h->stream));
  memset(info, 0, sizeof *info);
  info->ld = h->P.ld; info->lanes = (uint32_t)h->ops.lanes; info->vec = (uint32_t)h->ops.vec;
  info->pair_blocks = h->P.pair_blocks; info->device_bytes = h->device_bytes;
  info->last_npairs = c.npairs; info->last_rounds = c.last_rounds;
  info->kernels_per_step = (h->have_graph ? 3 : 0) + (h->P.lazy ? 3 : 4);   // [3 draw launches] prep, pairs, [blend], lambda
  return SVI_OK;
}

}  // extern "C"
