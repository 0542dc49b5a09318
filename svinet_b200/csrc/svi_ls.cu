// svi_ls.cu -- C ABI (include/svi_ls.h) over the sm_100a kernels in svi_ls_kernels.cuh.
//
// Host side of the device path only: CSR / segment construction, buffer management, kernel
// dispatch by K, stream plumbing.  No CPU compute fallback exists: every entry point that
// needs the GPU fails with SVI_ERR_CUDA when there is none.
#include "../../include/svi_ls.h"
#include "svi_common.h"
#include "svi_ls_kernels.cuh"
#include "svi_ls_ring.cuh"
#include "svi_ls_build.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace svi {
thread_local char g_err[512] = "";

// shared with svi_fa2.cu (svi_common.h)
int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace svi

namespace {
using svi::fail;
using svi::g_err;

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(SVI_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

using svi::Params;
using svi::s3_owner;

// launchers for one (G, V, LOGDOM) tiling
struct Ops {
  // phi sweep over segments [seg_first, seg_end); comm: with the link-community tally; publish: the tally also sets the
  // neighbour's bit (one arg-max per link)
  void (*phi)(const Params &, cudaStream_t, bool sparse, bool comm, uint32_t seg_first, uint32_t seg_end, uint32_t publish);
  void (*node)(const Params &, cudaStream_t, uint32_t blocks);
  void (*s3)(const Params &, cudaStream_t, uint32_t blocks);
  void (*lambda)(const Params &, cudaStream_t, int annealing, int update);
  void (*refresh)(const Params &, cudaStream_t, bool from_gacc);
  void (*heldout)(const Params &, cudaStream_t, uint64_t, const uint32_t *, const uint32_t *, const uint8_t *,
                  double, double *);
  int (*max_blocks_node)(int sms);
  int (*max_blocks_s3)(int sms);
  int lanes, vec, logdom;
  // second-generation sweeps (svi_ls_ring.cuh), present for 32 < K <= 256
  void (*phi_ring)(const Params &, cudaStream_t, bool sparse, bool comm, uint32_t seg_first, uint32_t seg_end,
                   uint32_t publish) = nullptr;
  void (*s3_ring)(const Params &, cudaStream_t, uint32_t blocks) = nullptr;
  void (*prepare_phi_ring)() = nullptr;   // per-device function attributes (dynamic shared memory), once per handle
  void (*prepare_s3_ring)() = nullptr;
  int (*max_blocks_s3_ring)(int sms, uint32_t ld) = nullptr;
  int ring_lanes = 0, ring_vec = 0, ring_depth = 0, ring_threads = 256;   // tiling of phi_ring
  int s3_lanes = 0, s3_vec = 0, s3_threads = 256;                         // tiling of s3_ring (may differ)
};

constexpr int kThreads = 256;

template <int G, int V, bool L>
struct Tile {
  static constexpr int CAP = 2 * G * V;
  static constexpr size_t kSmem = (size_t)(kThreads / G) * CAP * sizeof(double);

  static void phi(const Params &P, cudaStream_t st, bool sparse, bool comm, uint32_t s0, uint32_t s1, uint32_t pub) {
    if (s1 <= s0) return;
    const uint32_t blocks = (uint32_t)(((uint64_t)(s1 - s0) * G + kThreads - 1) / kThreads);
    if (sparse && comm) svi::k_phi<G, V, L, true, true><<<blocks, kThreads, 0, st>>>(P, s0, s1, pub);
    else if (sparse) svi::k_phi<G, V, L, true, false><<<blocks, kThreads, 0, st>>>(P, s0, s1, pub);
    else if (comm) svi::k_phi<G, V, L, false, true><<<blocks, kThreads, 0, st>>>(P, s0, s1, pub);
    else svi::k_phi<G, V, L, false, false><<<blocks, kThreads, 0, st>>>(P, s0, s1, pub);
  }
  static void node(const Params &P, cudaStream_t st, uint32_t blocks) {
    svi::k_node<G, V><<<blocks, kThreads, kSmem, st>>>(P);
  }
  static void s3(const Params &P, cudaStream_t st, uint32_t blocks) {
    svi::k_s3<G, V><<<blocks, kThreads, kSmem, st>>>(P);
  }
  static void lambda(const Params &P, cudaStream_t st, int annealing, int update) {
    svi::k_lambda<L><<<1, kThreads, 0, st>>>(P, annealing, update);
  }
  static void refresh(const Params &P, cudaStream_t st, bool from_gacc) {
    const uint32_t rows = P.node_end - P.node_begin;
    if (!rows) return;
    const uint32_t blocks = (uint32_t)(((uint64_t)rows * G + kThreads - 1) / kThreads);
    if (from_gacc) svi::k_refresh<G, V, L, true><<<blocks, kThreads, 0, st>>>(P);
    else svi::k_refresh<G, V, L, false><<<blocks, kThreads, 0, st>>>(P);
  }
  static void heldout(const Params &P, cudaStream_t st, uint64_t np, const uint32_t *p, const uint32_t *q,
                      const uint8_t *y, double eps, double *out) {
    if (!np) return;
    const uint32_t blocks = (uint32_t)((np * G + kThreads - 1) / kThreads);
    svi::k_heldout<G, V><<<blocks, kThreads, 0, st>>>(P, np, p, q, y, eps, out);
  }
  static int occ(const void *fn, int sms) {
    int per_sm = 0;
    if (kSmem > 48 * 1024) cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kThreads, kSmem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    return per_sm * sms;
  }
  static int max_blocks_node(int sms) { return occ((const void *)svi::k_node<G, V>, sms); }
  static int max_blocks_s3(int sms) { return occ((const void *)svi::k_s3<G, V>, sms); }
  static Ops ops() {
    return Ops{phi, node, s3, lambda, refresh, heldout, max_blocks_node, max_blocks_s3, G, V, L ? 1 : 0};
  }
};

// ring sweeps: G lanes per segment, V double2 per lane, R rows in flight per group, T threads per block
template <int G, int V, int R, int T, int MINB = 1>
struct RingTile {
  static constexpr int GPB = T / G;
  static constexpr int CAP = 2 * G * V;
  static constexpr size_t kSmem = (size_t)GPB * R * (CAP * 8 + 8) + (size_t)GPB * CAP * 4;   // ring + barriers + one-hot tallies
  template <class K>
  static void prep(K kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
  }
  using Sweep = svi::Sweep;
  static void prepare_phi() {
    prep(svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, false, false>);
    prep(svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, false, true>);
    prep(svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, true, false>);
    prep(svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, true, true>);
  }
  static void prepare_s3() { prep(svi::k_sweep_ring<G, V, R, T, MINB, Sweep::S3, false, false>); }
  static void phi(const Params &P, cudaStream_t st, bool sparse, bool comm, uint32_t s0, uint32_t s1, uint32_t pub) {
    if (s1 <= s0) return;
    const uint32_t blocks = (uint32_t)(((uint64_t)(s1 - s0) * G + T - 1) / T);
    if (sparse && comm) svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, true, true><<<blocks, T, kSmem, st>>>(P, s0, s1, pub);
    else if (sparse) svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, true, false><<<blocks, T, kSmem, st>>>(P, s0, s1, pub);
    else if (comm) svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, false, true><<<blocks, T, kSmem, st>>>(P, s0, s1, pub);
    else svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, false, false><<<blocks, T, kSmem, st>>>(P, s0, s1, pub);
  }
  static void s3(const Params &P, cudaStream_t st, uint32_t blocks) {
    if (P.nseg <= P.nseg_lo) return;
    svi::k_sweep_ring<G, V, R, T, MINB, Sweep::S3, false, false><<<blocks, T, kSmem, st>>>(P, P.nseg_lo, P.nseg, 0u);
  }
  static int max_blocks_s3(int sms, uint32_t) {
    auto kern = svi::k_sweep_ring<G, V, R, T, MINB, Sweep::S3, false, false>;
    prep(kern);
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, kSmem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    return per_sm * sms;
  }
  static void attach_phi(Ops *o) {
    o->phi_ring = phi;
    o->prepare_phi_ring = prepare_phi;
    o->ring_lanes = G;
    o->ring_vec = V;
    o->ring_depth = R;
    o->ring_threads = T;
  }
  static void attach_s3(Ops *o) {
    o->s3_ring = s3;
    o->prepare_s3_ring = prepare_s3;
    o->max_blocks_s3_ring = max_blocks_s3;
    o->s3_lanes = G;
    o->s3_vec = V;
    o->s3_threads = T;
  }
  static void attach(Ops *o) {
    attach_phi(o);
    attach_s3(o);
  }
};

#ifndef SVI_RING16_MINB
#define SVI_RING16_MINB 1
#endif
#ifndef SVI_RING16_R
#define SVI_RING16_R 4
#endif
#ifndef SVI_RING_T
#define SVI_RING_T 128
#endif
#ifndef SVI_RING_MINB
#define SVI_RING_MINB 3
#endif
void pick_ring(uint32_t k, Ops *o) {
  const char *off = getenv("SVI_LS_DISABLE_RING");
  if (off && off[0] == '1') return;
  const char *gsel = getenv("SVI_LS_RING_G");   // development A/B switch: 8 or 16
  const int want_g = gsel ? atoi(gsel) : 0;
  const uint32_t ld = (k + 3u) & ~3u;
  if (k <= 32 || k > 256) return;
  if (ld <= 104 && ld > 56 && want_g != 8) {
    // phi: G = 4, V = ceil(ld/8) in 8..13 -- eight neighbours per warp pass; at K = 100 the sweep is bound by
    // the per-neighbour overhead, not by bytes (the state fits L2), and measured 1.5x the G = 8 tile.
    // s3 : the G = 8 tile (its body is a plain row sum; G = 4 measured 1.5x SLOWER there)
    switch ((ld + 7) / 8) {
      case 8: RingTile<4, 8, 2, 128, SVI_RING_MINB>::attach_phi(o); break;
      case 9: RingTile<4, 9, 2, 128, SVI_RING_MINB>::attach_phi(o); break;
      case 10: RingTile<4, 10, 2, 128, SVI_RING_MINB>::attach_phi(o); break;
      case 11: RingTile<4, 11, 2, 128, SVI_RING_MINB>::attach_phi(o); break;
      case 12: RingTile<4, 12, 2, 128, SVI_RING_MINB>::attach_phi(o); break;
      default: RingTile<4, 13, 2, 128, SVI_RING_MINB>::attach_phi(o); break;
    }
    switch ((ld + 15) / 16) {
      case 4: RingTile<8, 4, 2, 256>::attach_s3(o); break;
      case 5: RingTile<8, 5, 2, 256>::attach_s3(o); break;
      case 6: RingTile<8, 6, 2, 256>::attach_s3(o); break;
      default: RingTile<8, 7, 2, 256>::attach_s3(o); break;
    }
  } else if (ld <= 112) {   // G = 8, V = ceil(ld/16)
    switch ((ld + 15) / 16) {
      case 3: RingTile<8, 3, 2, 256>::attach(o); break;
      case 4: RingTile<8, 4, 2, 256>::attach(o); break;
      case 5: RingTile<8, 5, 2, 256>::attach(o); break;
      case 6: RingTile<8, 6, 2, 256>::attach(o); break;
      default: RingTile<8, 7, 2, 256>::attach(o); break;
    }
  } else if (ld <= 208 && want_g != 16) {   // G = 8, V = ceil(ld/16) in 8..13
    switch ((ld + 15) / 16) {
      case 8: RingTile<8, 8, 2, SVI_RING_T, SVI_RING_MINB>::attach(o); break;
      case 9: RingTile<8, 9, 2, SVI_RING_T, SVI_RING_MINB>::attach(o); break;
      case 10: RingTile<8, 10, 2, SVI_RING_T, SVI_RING_MINB>::attach(o); break;
      case 11: RingTile<8, 11, 2, SVI_RING_T, SVI_RING_MINB>::attach(o); break;
      case 12: RingTile<8, 12, 2, SVI_RING_T, SVI_RING_MINB>::attach(o); break;
      default: RingTile<8, 13, 2, SVI_RING_T, SVI_RING_MINB>::attach(o); break;
    }
  } else {   // G = 16, V = ceil(ld/32) in 4..8
    switch ((ld + 31) / 32) {
      case 4: RingTile<16, 4, SVI_RING16_R, 256, SVI_RING16_MINB>::attach(o); break;
      case 5: RingTile<16, 5, SVI_RING16_R, 256, SVI_RING16_MINB>::attach(o); break;
      case 6: RingTile<16, 6, SVI_RING16_R, 256, SVI_RING16_MINB>::attach(o); break;
      case 7: RingTile<16, 7, SVI_RING16_R, 256, SVI_RING16_MINB>::attach(o); break;
      default: RingTile<16, 8, SVI_RING16_R, 256, SVI_RING16_MINB>::attach(o); break;
    }
  }
}

// K -> tiling.  Factorised (exp-domain) rows up to K = 256; log-domain above (underflow, see .cuh)
bool pick_ops(uint32_t k, Ops *o) {
  if (k == 0) return false;
  if (k <= 4) *o = Tile<2, 1, false>::ops();
  else if (k <= 8) *o = Tile<4, 1, false>::ops();
  else if (k <= 16) *o = Tile<8, 1, false>::ops();
  else if (k <= 32) *o = Tile<16, 1, false>::ops();
  else if (k <= 64) *o = Tile<32, 1, false>::ops();
  else if (k <= 128) *o = Tile<32, 2, false>::ops();
  else if (k <= 192) *o = Tile<32, 3, false>::ops();
  else if (k <= 256) *o = Tile<32, 4, false>::ops();
  else if (k <= 384) *o = Tile<32, 6, true>::ops();
  else if (k <= 512) *o = Tile<32, 8, true>::ops();
  else if (k <= 768) *o = Tile<32, 12, true>::ops();
  else if (k <= 1024) *o = Tile<32, 16, true>::ops();
  else return false;
  pick_ring(k, o);
  return true;
}

template <class T>
cudaError_t dalloc(T **p, size_t count, uint64_t *total) {
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void **)p, bytes);
  if (e == cudaSuccess) {
    *total += bytes;
    e = cudaMemset(*p, 0, bytes);
  }
  return e;
}

}  // namespace

struct svi_ls {
  svi_ls_config cfg{};
  int device = 0, sms = 0;
  cudaStream_t stream = nullptr;
  Ops ops{};
  Params P{};
  uint32_t nlocal = 0, blocks_node = 0, blocks_s3 = 0, kpart_blocks = 0;
  uint64_t he_phi = 0, he_s3 = 0, device_bytes = 0;
  uint32_t seg_len = 0;
  // owned device memory
  uint32_t *d_col = nullptr, *d_seg_node = nullptr, *d_seg_beg = nullptr, *d_seg_cnt = nullptr, *d_seg_nnc = nullptr;
  uint32_t *d_node_seg_lo = nullptr, *d_node_seg_up = nullptr, *d_conv_dirty = nullptr;
  bool force_partition = false;
  bool shard = false;            // the handle owns a proper node block of the graph
  bool partition_every_sweep = false;   // converged flags of other shards arrive by exchange: no local dirty flag
  double *d_tl = nullptr, *d_b = nullptr, *d_mphi = nullptr, *d_gamma = nullptr, *d_gacc = nullptr;
  double *d_part = nullptr, *d_kvec = nullptr, *d_kpart = nullptr, *d_lambda = nullptr, *d_eb = nullptr;
  double *d_scale = nullptr, *d_stage = nullptr;
  uint32_t *d_conv = nullptr, *d_active = nullptr, *d_abits = nullptr, *d_mbits = nullptr;
  uint32_t *d_conv_snap = nullptr;   // `converged` as the s3 sweep must see it when the refresh ran first
  bool conv_snap_valid = false;
  size_t stage_elems = 0;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

void free_all(svi_ls *h) {
  void *ptrs[] = {h->d_col, h->d_seg_node, h->d_seg_beg, h->d_seg_cnt, h->d_seg_nnc, h->d_node_seg_lo,
                  h->d_node_seg_up, h->d_conv_dirty, h->d_tl, h->d_b, h->d_mphi, h->d_gamma, h->d_gacc, h->d_part,
                  h->d_kvec, h->d_kpart, h->d_lambda, h->d_eb, h->d_scale, h->d_stage, h->d_conv, h->d_active,
                  h->d_abits, h->d_mbits, h->d_conv_snap};
  for (void *p : ptrs)
    if (p) cudaFree(p);
}

int ensure_stage(svi_ls *h, size_t elems) {
  if (h->stage_elems >= elems) return SVI_OK;
  if (h->d_stage) cudaFree(h->d_stage);
  h->d_stage = nullptr;
  h->stage_elems = 0;
  CK(cudaMalloc((void **)&h->d_stage, std::max<size_t>(elems, 1) * sizeof(double)));
  h->stage_elems = elems;
  return SVI_OK;
}

// balanced split of `deg` neighbours into chunks of at most seg_len
inline void push_segments(uint32_t node, uint32_t beg, uint32_t deg, uint32_t seg_len, std::vector<uint32_t> &sn,
                          std::vector<uint32_t> &sb, std::vector<uint32_t> &sc) {
  if (!deg) return;
  const uint32_t nch = (deg + seg_len - 1) / seg_len;
  const uint32_t base = deg / nch, extra = deg % nch;
  uint32_t at = beg;
  for (uint32_t c = 0; c < nch; ++c) {
    const uint32_t len = base + (c < extra ? 1u : 0u);
    sn.push_back(node);
    sb.push_back(at);
    sc.push_back(len);
    at += len;
  }
}

}  // namespace

extern "C" {

const char *svi_ls_last_error(void) { return g_err; }
int svi_ls_abi_version(void) { return SVI_LS_ABI_VERSION; }

int svi_ls_create(const svi_ls_config *cfg, const uint32_t *links, const double *tl, svi_ls **out) {
  if (!cfg || !out || (!links && cfg->nlinks)) return fail(SVI_ERR_INVALID, "svi_ls_create: null argument");
  *out = nullptr;
  if (cfg->n == 0 || cfg->k == 0) return fail(SVI_ERR_INVALID, "svi_ls_create: n and k must be positive");
  if (cfg->node_begin > cfg->node_end || cfg->node_end > cfg->n)
    return fail(SVI_ERR_INVALID, "svi_ls_create: bad shard [%u,%u) for n=%u", cfg->node_begin, cfg->node_end, cfg->n);
  Ops ops;
  if (!pick_ops(cfg->k, &ops)) return fail(SVI_ERR_UNSUPPORTED, "svi_ls_create: k=%u not supported (max 1024)", cfg->k);

  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return fail(SVI_ERR_CUDA, "svi_ls_create: no CUDA device");
  int dev = cfg->device;
  if (dev < 0) CK(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(SVI_ERR_INVALID, "svi_ls_create: device %d of %d", dev, ndev);

  svi_ls *h = new (std::nothrow) svi_ls();
  if (!h) return fail(SVI_ERR_NOMEM, "svi_ls_create: host allocation failed");
  h->cfg = *cfg;
  h->device = dev;
  h->ops = ops;
  DeviceGuard guard(dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    delete h;
    return fail(SVI_ERR_CUDA, "svi_ls_create: cudaGetDeviceProperties failed");
  }
  h->sms = prop.multiProcessorCount;

  const uint32_t n = cfg->n, k = cfg->k, nb = cfg->node_begin, ne = cfg->node_end;
  const uint32_t nlocal = ne - nb, ld = (k + 3u) & ~3u, words = (k + 31u) / 32u;
  h->nlocal = nlocal;
  h->shard = nb != 0 || ne != n;
  h->partition_every_sweep = h->shard;
  if (n >= 0x80000000u) {
    delete h;
    return fail(SVI_ERR_UNSUPPORTED, "svi_ls_create: n=%u exceeds the 31-bit node ids of the CSR build", n);
  }

  // ---- CSR of the shard's half-edges, built on the device (svi_ls_build.cuh); the neighbours a node OWNS for
  //      the s3 sweep are stored last in its list ----
  std::vector<uint32_t> deg_lo(nlocal, 0), deg_up(nlocal, 0);
  uint64_t he = 0, he3 = 0;
  {
    const uint64_t nl = cfg->nlinks;
    uint32_t *d_links = nullptr, *d_dlo = nullptr, *d_dup = nullptr;
    uint64_t *d_keys = nullptr, *d_keys2 = nullptr;
    unsigned long long *d_err = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    auto release = [&]() {
      for (void *q : {(void *)d_links, (void *)d_dlo, (void *)d_dup, (void *)d_keys, (void *)d_keys2, (void *)d_err, d_tmp})
        if (q) cudaFree(q);
    };
    A(cudaMalloc((void **)&d_links, std::max<uint64_t>(nl, 1) * 2 * sizeof(uint32_t)));
    A(cudaMalloc((void **)&d_keys, std::max<uint64_t>(nl, 1) * 2 * sizeof(uint64_t)));
    A(cudaMalloc((void **)&d_keys2, std::max<uint64_t>(nl, 1) * 2 * sizeof(uint64_t)));
    A(cudaMalloc((void **)&d_dlo, std::max<uint32_t>(nlocal, 1) * sizeof(uint32_t)));
    A(cudaMalloc((void **)&d_dup, std::max<uint32_t>(nlocal, 1) * sizeof(uint32_t)));
    A(cudaMalloc((void **)&d_err, 2 * sizeof(unsigned long long)));
    int end_bit = 34;
    while (end_bit < 64 && ((uint64_t)nlocal >> (end_bit - 33)) != 0) ++end_bit;
    cub::DoubleBuffer<uint64_t> dbuf(d_keys, d_keys2);
    if (e == cudaSuccess) A(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, dbuf, (int64_t)(2 * nl), 0, end_bit));
    A(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 1)));
    if (e == cudaSuccess) {
      A(cudaMemcpy(d_links, links, nl * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice));
      A(cudaMemset(d_dlo, 0, std::max<uint32_t>(nlocal, 1) * sizeof(uint32_t)));
      A(cudaMemset(d_dup, 0, std::max<uint32_t>(nlocal, 1) * sizeof(uint32_t)));
      A(cudaMemset(d_err, 0, 2 * sizeof(unsigned long long)));
    }
    unsigned long long err[2] = {0, 0};
    if (e == cudaSuccess && nl) {
      svi::k_build_keys<<<h->sms * 8, 256>>>(d_links, nl, n, nb, ne, d_keys, d_dlo, d_dup, d_err);
      A(cudaGetLastError());
      A(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, dbuf, (int64_t)(2 * nl), 0, end_bit));
    }
    if (e == cudaSuccess) {
      A(cudaMemcpy(err, d_err, sizeof err, cudaMemcpyDeviceToHost));
      A(cudaMemcpy(deg_lo.data(), d_dlo, nlocal * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      A(cudaMemcpy(deg_up.data(), d_dup, nlocal * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    if (e == cudaSuccess && err[0]) {
      release();
      delete h;
      const uint64_t bad = err[1];
      return fail(SVI_ERR_INVALID, "svi_ls_create: link %llu = (%u,%u) out of range (%llu such links)",
                  (unsigned long long)bad, links[2 * bad], links[2 * bad + 1], err[0]);
    }
    for (uint32_t v = 0; v < nlocal; ++v) {
      he += (uint64_t)deg_lo[v] + deg_up[v];
      he3 += deg_up[v];
    }
    if (e == cudaSuccess && he > 0xffffffffull) {
      release();
      delete h;
      return fail(SVI_ERR_UNSUPPORTED, "svi_ls_create: %llu half-edges exceed the 32-bit CSR of one shard",
                  (unsigned long long)he);
    }
    if (e == cudaSuccess) {
      A(dalloc(&h->d_col, he, &h->device_bytes));
      if (e == cudaSuccess && he) {
        svi::k_keys_to_col<<<h->sms * 8, 256>>>(dbuf.Current(), he, h->d_col);
        A(cudaGetLastError());
        A(cudaDeviceSynchronize());
      }
    }
    release();
    if (e != cudaSuccess) {
      const int rc = fail(e == cudaErrorMemoryAllocation ? SVI_ERR_NOMEM : SVI_ERR_CUDA, "svi_ls_create (graph build): %s",
                          cudaGetErrorString(e));
      free_all(h);
      delete h;
      return rc;
    }
  }
  h->he_phi = he;
  h->he_s3 = he3;
  std::vector<double> tl_host(n, 0.0);
  if (tl) std::copy(tl, tl + n, tl_host.begin());
  else   // Q3: both adjacency directions count each link for both endpoints.  A shard knows its own nodes' degrees,
    for (uint32_t v = 0; v < nlocal; ++v) tl_host[nb + v] = 2.0 * ((double)deg_lo[v] + deg_up[v]);   // and reads no others

  // ---- work segments: every node's "lo" part and "up" (owned) part are cut separately ----
  uint32_t seg_len = cfg->seg_len;
  if (!seg_len) {
    // enough segments to fill the machine several times over, long enough to amortise the
    // per-segment row load/store
    const uint64_t target = (uint64_t)h->sms * 64 * (32 / ops.lanes > 0 ? 32 / ops.lanes : 1);
    seg_len = 256;
    while (seg_len > 16 && he / seg_len < target) seg_len >>= 1;
  }
  seg_len = std::min(seg_len, svi::kMaxSegLen);
  h->seg_len = seg_len;
  std::vector<uint32_t> sn, sb, sc, nlo(nlocal + 1, 0), nup(nlocal + 1, 0);
  sn.reserve(he / seg_len + 2 * (size_t)nlocal);
  sb.reserve(he / seg_len + 2 * (size_t)nlocal);
  sc.reserve(he / seg_len + 2 * (size_t)nlocal);
  {
    uint64_t at = 0;
    for (uint32_t v = 0; v < nlocal; ++v) {
      push_segments(nb + v, (uint32_t)at, deg_lo[v], seg_len, sn, sb, sc);
      nlo[v + 1] = (uint32_t)sn.size();
      at += (uint64_t)deg_lo[v] + deg_up[v];
    }
    const uint32_t nseg_lo = (uint32_t)sn.size();
    at = 0;
    nup[0] = nseg_lo;
    for (uint32_t v = 0; v < nlocal; ++v) {
      push_segments(nb + v, (uint32_t)(at + deg_lo[v]), deg_up[v], seg_len, sn, sb, sc);
      nup[v + 1] = (uint32_t)sn.size();
      at += (uint64_t)deg_lo[v] + deg_up[v];
    }
  }
  const uint32_t nseg_lo = nlo[nlocal], nseg = (uint32_t)sn.size(), nseg3 = nseg - nseg_lo;

  // ---- device memory ----
  uint64_t &tot = h->device_bytes;
  const size_t nld = (size_t)n * ld;
  h->blocks_node = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(ops.max_blocks_node(h->sms),
                                                                   ((int64_t)nlocal * ops.lanes + kThreads - 1) / kThreads));
  if (ops.s3_ring)
    h->blocks_s3 = (uint32_t)std::max<int64_t>(
        1, std::min<int64_t>(ops.max_blocks_s3_ring(h->sms, ld),
                             ((int64_t)nseg3 * ops.s3_lanes + ops.s3_threads - 1) / ops.s3_threads));
  else
    h->blocks_s3 = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(ops.max_blocks_s3(h->sms),
                                                                   ((int64_t)nseg3 * ops.lanes + kThreads - 1) / kThreads));
  if (ops.prepare_phi_ring) ops.prepare_phi_ring();
  if (ops.prepare_s3_ring) ops.prepare_s3_ring();
  h->kpart_blocks = std::max(h->blocks_node, h->blocks_s3);
  const size_t cap = std::max(2 * (size_t)ops.lanes * ops.vec, 2 * (size_t)ops.s3_lanes * ops.s3_vec);
  cudaError_t e = cudaSuccess;
  auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  A(dalloc(&h->d_seg_node, nseg, &tot));
  A(dalloc(&h->d_seg_beg, nseg, &tot));
  A(dalloc(&h->d_seg_cnt, nseg, &tot));
  A(dalloc(&h->d_seg_nnc, nseg, &tot));
  A(dalloc(&h->d_node_seg_lo, nlocal + 1, &tot));
  A(dalloc(&h->d_node_seg_up, nlocal + 1, &tot));
  A(dalloc(&h->d_conv_dirty, 1, &tot));
  A(dalloc(&h->d_tl, n, &tot));
  A(dalloc(&h->d_b, nld, &tot));
  A(dalloc(&h->d_mphi, nld, &tot));
  A(dalloc(&h->d_gamma, nld, &tot));
  A(dalloc(&h->d_gacc, nld, &tot));
  A(dalloc(&h->d_part, (size_t)nseg * ld, &tot));
  A(dalloc(&h->d_kvec, 4 * (size_t)ld, &tot));
  A(dalloc(&h->d_kpart, (size_t)h->kpart_blocks * 3 * cap, &tot));
  A(dalloc(&h->d_lambda, 2 * (size_t)k, &tot));
  A(dalloc(&h->d_eb, ld, &tot));
  A(dalloc(&h->d_scale, ld, &tot));
  A(dalloc(&h->d_conv, n, &tot));
  A(dalloc(&h->d_active, n, &tot));
  A(dalloc(&h->d_abits, (size_t)n * words, &tot));
  A(dalloc(&h->d_mbits, (size_t)n * words, &tot));
  auto H2D = [&](void *d, const void *s, size_t bytes) {
    if (bytes) A(cudaMemcpy(d, s, bytes, cudaMemcpyHostToDevice));
  };
  if (e == cudaSuccess) {
    H2D(h->d_seg_node, sn.data(), nseg * sizeof(uint32_t));
    H2D(h->d_seg_beg, sb.data(), nseg * sizeof(uint32_t));
    H2D(h->d_seg_cnt, sc.data(), nseg * sizeof(uint32_t));
    H2D(h->d_seg_nnc, sc.data(), nseg * sizeof(uint32_t));   // nobody has converged: every neighbour is "not converged"
    H2D(h->d_node_seg_lo, nlo.data(), (nlocal + 1) * sizeof(uint32_t));
    H2D(h->d_node_seg_up, nup.data(), (nlocal + 1) * sizeof(uint32_t));
    H2D(h->d_tl, tl_host.data(), n * sizeof(double));
  }
  if (e != cudaSuccess) {
    const int rc = fail(e == cudaErrorMemoryAllocation ? SVI_ERR_NOMEM : SVI_ERR_CUDA, "svi_ls_create: %s",
                        cudaGetErrorString(e));
    free_all(h);
    delete h;
    return rc;
  }

  Params &P = h->P;
  P.n = n; P.k = k; P.ld = ld; P.words = words;
  P.node_begin = nb; P.node_end = ne;
  P.alpha = cfg->alpha; P.eta0 = cfg->eta0; P.eta1 = cfg->eta1; P.ones_d = (double)cfg->ones;
  P.k_div10 = k / 10;
  P.col = h->d_col;
  P.seg_node = h->d_seg_node; P.seg_beg = h->d_seg_beg; P.seg_cnt = h->d_seg_cnt; P.seg_nnc = h->d_seg_nnc;
  P.nseg = nseg; P.nseg_lo = nseg_lo;
  P.node_seg_lo = h->d_node_seg_lo; P.node_seg_up = h->d_node_seg_up;
  P.conv_dirty = h->d_conv_dirty;
  P.tl = h->d_tl;
  P.b = h->d_b; P.mphi = h->d_mphi; P.gamma = h->d_gamma; P.gacc = h->d_gacc; P.part = h->d_part;
  P.kvec = h->d_kvec; P.kpart = h->d_kpart; P.lambda = h->d_lambda; P.eb = h->d_eb; P.scale = h->d_scale;
  P.conv = h->d_conv; P.active = h->d_active; P.abits = h->d_abits; P.mbits = h->d_mbits;
  *out = h;
  return SVI_OK;
}

void svi_ls_destroy(svi_ls *h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  cudaStreamSynchronize(h->stream);
  free_all(h);
  delete h;
}

int svi_ls_set_stream(svi_ls *h, void *cuda_stream) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  h->stream = (cudaStream_t)cuda_stream;
  return SVI_OK;
}

int svi_ls_sync(svi_ls *h) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_ls_set_state(svi_ls *h, const double *gamma, const double *lambda) {
  if (!h || !gamma || !lambda) return fail(SVI_ERR_INVALID, "svi_ls_set_state: null argument");
  DeviceGuard guard(h->device);
  const Params &P = h->P;
  const size_t nk = (size_t)P.n * P.k;
  if (P.ld == P.k) {
    CK(cudaMemcpyAsync(h->d_gamma, gamma, nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  } else {
    int rc = ensure_stage(h, nk);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_stage, gamma, nk * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    svi::k_pad_rows<<<h->sms * 8, 256, 0, h->stream>>>(h->d_stage, h->d_gamma, P.n, P.k, P.ld);
  }
  CK(cudaMemcpyAsync(h->d_lambda, lambda, 2 * (size_t)P.k * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  // expectations for ALL rows (a shard still needs its neighbours' factors before the first sweep)
  Params all = P;
  all.node_begin = 0;
  all.node_end = P.n;
  h->ops.refresh(all, h->stream, false);
  h->ops.lambda(P, h->stream, 0, 0);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));   // host buffers may be pageable: do not return early
  return SVI_OK;
}

int svi_ls_get_state(svi_ls *h, double *gamma, double *lambda) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  const Params &P = h->P;
  const size_t nk = (size_t)P.n * P.k;
  if (gamma) {
    if (P.ld == P.k) {
      CK(cudaMemcpyAsync(gamma, h->d_gamma, nk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    } else {
      int rc = ensure_stage(h, nk);
      if (rc) return rc;
      svi::k_unpad_rows<<<h->sms * 8, 256, 0, h->stream>>>(h->d_gamma, h->d_stage, P.n, P.k, P.ld);
      CK(cudaMemcpyAsync(gamma, h->d_stage, nk * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
  }
  if (lambda)
    CK(cudaMemcpyAsync(lambda, h->d_lambda, 2 * (size_t)P.k * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_ls_set_converged(svi_ls *h, const uint32_t *converged) {
  if (!h || !converged) return fail(SVI_ERR_INVALID, "svi_ls_set_converged: null argument");
  DeviceGuard guard(h->device);
  CK(cudaMemcpyAsync(h->d_conv, converged, (size_t)h->P.n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->force_partition = true;   // arbitrary flags (a resume may even clear some): rebuild every segment's partition
  return SVI_OK;
}

int svi_ls_get_converged(svi_ls *h, uint32_t *converged, uint32_t *active_comms) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  if (converged)
    CK(cudaMemcpyAsync(converged, h->d_conv, (size_t)h->P.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  if (active_comms)
    CK(cudaMemcpyAsync(active_comms, h->d_active, (size_t)h->P.n * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                       h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_ls_phase_phi(svi_ls *h, uint32_t iter, int write_comm) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  const Params &P = h->P;
  h->conv_snap_valid = false;   // a new iteration: a snapshot left by an unfinished refresh/lambda pair is stale
  // the ring sweeps read neighbour lists partitioned by the converged flags (svi_ls_ring.cuh: k_partition); nodes
  // converge in k_refresh, at the end of the previous iteration
  if ((h->ops.phi_ring || h->ops.s3_ring) && P.nseg) {
    const bool force = h->force_partition || h->partition_every_sweep;
    svi::k_partition<<<(uint32_t)std::min<uint64_t>((P.nseg + 7) / 8, (uint64_t)h->sms * 8), 256, 0, h->stream>>>(P, force);
    CK(cudaMemsetAsync(h->d_conv_dirty, 0, sizeof(uint32_t), h->stream));
    h->force_partition = false;
  }
  if (write_comm)  // _communities.clear(); _fmap.zero()  (src/linksampling.cc:584-587)
    CK(cudaMemsetAsync(h->d_mbits, 0, (size_t)P.n * P.words * sizeof(uint32_t), h->stream));
  auto phi = h->ops.phi_ring ? h->ops.phi_ring : h->ops.phi;
  const bool sparse = iter > 1000 && P.k_div10 > 0;
  if (!write_comm) {
    phi(P, h->stream, sparse, false, 0, P.nseg, 0);
  } else if (h->shard) {
    // a shard holds the membership words of its own nodes only: the arg-max is taken on both sides of a link
    phi(P, h->stream, sparse, true, 0, P.nseg, 0);
  } else {
    // one arg-max per LINK (src/linksampling.cc:704-717 sets fmap[p] and fmap[q] from one max_k): the owner's side
    // computes it and publishes both endpoints' bits; the other side runs without the tally
    phi(P, h->stream, sparse, false, 0, P.nseg_lo, 0);
    phi(P, h->stream, sparse, true, P.nseg_lo, P.nseg, 1);
  }
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_phase_node(svi_ls *h) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  const Params &P = h->P;
  h->ops.node(P, h->stream, h->blocks_node);
  svi::k_reduce_kpart<<<4, 256, 0, h->stream>>>(h->d_kpart, h->blocks_node, 3, 2 * h->ops.lanes * h->ops.vec,
                                               h->d_kvec, P.ld);
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_phase_s3(svi_ls *h) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  Params P = h->P;
  if (h->conv_snap_valid) P.conv = h->d_conv_snap;   // the reference's s3 loop (:731-746) runs BEFORE prune (:761)
  uint32_t cap3 = 2 * h->ops.lanes * h->ops.vec;
  if (h->ops.s3_ring) {
    h->ops.s3_ring(P, h->stream, h->blocks_s3);
    cap3 = 2 * h->ops.s3_lanes * h->ops.s3_vec;
  } else {
    h->ops.s3(P, h->stream, h->blocks_s3);
  }
  svi::k_reduce_kpart<<<2, 256, 0, h->stream>>>(h->d_kpart, h->blocks_s3, 1, cap3, h->d_kvec + 3 * (size_t)P.ld, P.ld);
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_phase_finish(svi_ls *h, int annealing) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  h->ops.lambda(h->P, h->stream, annealing, 1);
  h->ops.refresh(h->P, h->stream, true);
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_phase_refresh(svi_ls *h, int annealing) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  if (!h->d_conv_snap) CK(cudaMalloc((void **)&h->d_conv_snap, std::max<size_t>(h->P.n, 1) * sizeof(uint32_t)));
  CK(cudaMemcpyAsync(h->d_conv_snap, h->d_conv, (size_t)h->P.n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->stream));
  h->conv_snap_valid = true;
  svi::k_scale<<<1, 256, 0, h->stream>>>(h->P, annealing);
  h->ops.refresh(h->P, h->stream, true);
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_phase_lambda(svi_ls *h, int annealing) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  h->ops.lambda(h->P, h->stream, annealing, 1);
  h->conv_snap_valid = false;
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_step(svi_ls *h, uint32_t iter, int annealing, int write_comm) {
  int rc;
  if ((rc = svi_ls_phase_phi(h, iter, write_comm))) return rc;
  if ((rc = svi_ls_phase_node(h))) return rc;
  if ((rc = svi_ls_phase_s3(h))) return rc;
  return svi_ls_phase_finish(h, annealing);
}

int svi_ls_get_membership(svi_ls *h, uint32_t *bits) {
  if (!h || !bits) return fail(SVI_ERR_INVALID, "svi_ls_get_membership: null argument");
  DeviceGuard guard(h->device);
  CK(cudaMemcpyAsync(bits, h->d_mbits, (size_t)h->P.n * h->P.words * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                     h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_ls_heldout(svi_ls *h, uint64_t npairs, const uint32_t *p, const uint32_t *q, const uint8_t *y,
                   double epsilon, double *loglik) {
  if (!h || (npairs && (!p || !q || !y || !loglik))) return fail(SVI_ERR_INVALID, "svi_ls_heldout: null argument");
  if (!npairs) return SVI_OK;
  DeviceGuard guard(h->device);
  for (uint64_t i = 0; i < npairs; ++i)
    if (p[i] >= h->P.n || q[i] >= h->P.n) return fail(SVI_ERR_INVALID, "svi_ls_heldout: pair %llu out of range",
                                                      (unsigned long long)i);
  // staging: [p | q] as uint32, y as bytes, out as double -- carve from one scratch allocation
  const size_t bytes = npairs * (2 * sizeof(uint32_t) + sizeof(double)) + ((npairs + 7) & ~(size_t)7);
  int rc = ensure_stage(h, (bytes + sizeof(double) - 1) / sizeof(double));
  if (rc) return rc;
  double *d_out = h->d_stage;
  uint32_t *d_p = reinterpret_cast<uint32_t *>(d_out + npairs);
  uint32_t *d_q = d_p + npairs;
  uint8_t *d_y = reinterpret_cast<uint8_t *>(d_q + npairs);
  CK(cudaMemcpyAsync(d_p, p, npairs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_q, q, npairs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_y, y, npairs, cudaMemcpyHostToDevice, h->stream));
  h->ops.heldout(h->P, h->stream, npairs, d_p, d_q, d_y, epsilon, d_out);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(loglik, d_out, npairs * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_ls_get_kvectors(svi_ls *h, double *sum, double *s1, double *s2, double *s3) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  double *dst[4] = {sum, s1, s2, s3};
  for (int v = 0; v < 4; ++v)
    if (dst[v])
      CK(cudaMemcpyAsync(dst[v], h->d_kvec + (size_t)v * h->P.ld, h->P.k * sizeof(double), cudaMemcpyDeviceToHost,
                         h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

int svi_ls_device_buffer(svi_ls *h, svi_buffer which, void **dev_ptr, uint64_t *ld) {
  if (!h || !dev_ptr) return fail(SVI_ERR_INVALID, "svi_ls_device_buffer: null argument");
  uint64_t l = h->P.ld;
  switch (which) {
    case SVI_BUF_EXPPI: *dev_ptr = h->d_b; break;
    case SVI_BUF_MPHI: *dev_ptr = h->d_mphi; break;
    case SVI_BUF_GAMMA: *dev_ptr = h->d_gamma; break;
    case SVI_BUF_KVEC: *dev_ptr = h->d_kvec; break;
    case SVI_BUF_CONVERGED: *dev_ptr = h->d_conv; l = 1; break;
    case SVI_BUF_LAMBDA: *dev_ptr = h->d_lambda; l = 2; break;
    case SVI_BUF_ACTIVE: *dev_ptr = h->d_active; l = 1; break;
    case SVI_BUF_ACTIVE_BITS: *dev_ptr = h->d_abits; l = h->P.words; break;
    case SVI_BUF_MEMBER_BITS: *dev_ptr = h->d_mbits; l = h->P.words; break;
    default: return fail(SVI_ERR_INVALID, "svi_ls_device_buffer: unknown buffer %d", (int)which);
  }
  if (ld) *ld = l;
  return SVI_OK;
}

int svi_ls_get_info(svi_ls *h, svi_ls_info *info) {
  if (!h || !info) return fail(SVI_ERR_INVALID, "svi_ls_get_info: null argument");
  info->half_edges_phi = h->he_phi;
  info->half_edges_s3 = h->he_s3;
  info->segments_phi = h->P.nseg;
  info->segments_s3 = h->P.nseg - h->P.nseg_lo;
  info->ld = h->P.ld;
  info->seg_len = h->seg_len;
  info->lanes = (uint32_t)(h->ops.phi_ring ? h->ops.ring_lanes : h->ops.lanes);
  info->vec = (uint32_t)(h->ops.phi_ring ? h->ops.ring_vec : h->ops.vec);
  info->ring_depth = (uint32_t)h->ops.ring_depth;
  info->device_bytes = h->device_bytes;
  // partition (ring tilings), phi (two launches with the tally on a whole graph), node, reduce, s3, reduce, lambda, refresh
  info->kernels_per_step = 7 + ((h->ops.phi_ring || h->ops.s3_ring) ? 1 : 0) + (h->shard ? 0 : 1);
  return SVI_OK;
}

}  // extern "C"
