// svi_ls.cu -- C ABI (include/svi_ls.h) over the sm_100a kernels in svi_ls_kernels.cuh.
//
// Host side of the device path only: CSR / segment construction, buffer management, kernel
// dispatch by K, stream plumbing.  No CPU compute fallback exists: every entry point that
// needs the GPU fails with SVI_ERR_CUDA when there is none.
#include "../../include/svi_ls.h"
#include "svi_common.h"
#include "svi_ls_kernels.cuh"
#include "svi_ls_ring.cuh"
#include "svi_ls_wide.cuh"
#include "svi_ls_build.cuh"
#include "svi_ls_mg.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
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
  void (*phi)(const Params &, cudaStream_t, bool sparse, bool comm, uint32_t seg_first, uint32_t seg_end,
              uint32_t seg_first2, uint32_t seg_end2, uint32_t publish);
  void (*node)(const Params &, cudaStream_t, uint32_t blocks);
  void (*s3)(const Params &, cudaStream_t, uint32_t blocks);
  void (*lambda)(const Params &, cudaStream_t, int annealing, int update);
  void (*refresh)(const Params &, cudaStream_t, bool from_gacc);
  void (*heldout)(const Params &, cudaStream_t, uint64_t, const uint32_t *, const uint32_t *, const uint8_t *,
                  double, double *, unsigned long long *bad, const svi::Peers *peer_rows);
  int (*max_blocks_node)(int sms);
  int (*max_blocks_s3)(int sms);
  int lanes, vec, logdom;
  // second-generation sweeps (svi_ls_ring.cuh), present for 32 < K <= 256
  void (*phi_ring)(const Params &, cudaStream_t, bool sparse, bool comm, uint32_t seg_first, uint32_t seg_end,
                   uint32_t seg_first2, uint32_t seg_end2, uint32_t publish) = nullptr;
  void (*s3_ring)(const Params &, cudaStream_t, uint32_t blocks) = nullptr;
  void (*preload)() = nullptr;            // force-load every kernel of the tile (see preload_common)
  void (*prepare_phi_ring)() = nullptr;   // per-device function attributes (dynamic shared memory), once per handle
  void (*prepare_s3_ring)() = nullptr;
  int (*max_blocks_s3_ring)(int sms, uint32_t ld) = nullptr;
  int ring_lanes = 0, ring_vec = 0, ring_depth = 0, ring_threads = 256;   // tiling of phi_ring
  int s3_lanes = 0, s3_vec = 0, s3_threads = 256;                         // tiling of s3_ring (may differ)
};

// CUDA loads kernels lazily, on their first launch, and that load can wait for running kernels to finish.  A shard
// whose stream sits in a flag wait (svi_ls_mg.cuh) while its host thread launches a never-loaded kernel would then
// dead-lock with its peers, so every kernel is loaded when the handle is created.
template <class K>
void touch(K kern) {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, (const void *)kern);
}

constexpr int kThreads = 256;
constexpr uint32_t kMaxChunks = 8;   // pipeline chunks of a shard's node block in svi_ls_mg_step

template <int G, int V, bool L>
struct Tile {
  static constexpr int CAP = 2 * G * V;
  static constexpr size_t kSmem = (size_t)(kThreads / G) * CAP * sizeof(double);

  static void phi(const Params &P, cudaStream_t st, bool sparse, bool comm, uint32_t s0, uint32_t s1, uint32_t t0,
                  uint32_t t1, uint32_t pub) {
    const uint64_t cnt = (uint64_t)(s1 - s0) + (t1 - t0);
    if (!cnt) return;
    const uint32_t blocks = (uint32_t)((cnt * G + kThreads - 1) / kThreads);
    if (sparse && comm) svi::k_phi<G, V, L, true, true><<<blocks, kThreads, 0, st>>>(P, s0, s1, t0, t1, pub);
    else if (sparse) svi::k_phi<G, V, L, true, false><<<blocks, kThreads, 0, st>>>(P, s0, s1, t0, t1, pub);
    else if (comm) svi::k_phi<G, V, L, false, true><<<blocks, kThreads, 0, st>>>(P, s0, s1, t0, t1, pub);
    else svi::k_phi<G, V, L, false, false><<<blocks, kThreads, 0, st>>>(P, s0, s1, t0, t1, pub);
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
                      const uint8_t *y, double eps, double *out, unsigned long long *bad, const svi::Peers *peer_rows) {
    if (!np) return;
    const uint32_t blocks = (uint32_t)((np * G + kThreads - 1) / kThreads);
    if (peer_rows) svi::k_heldout<G, V, svi::Peers><<<blocks, kThreads, 0, st>>>(P, *peer_rows, np, p, q, y, eps, out, bad);
    else svi::k_heldout<G, V, svi::LocalRows><<<blocks, kThreads, 0, st>>>(P, svi::LocalRows{P.gamma}, np, p, q, y, eps, out, bad);
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
  static void preload() {
    touch(svi::k_phi<G, V, L, false, false>);
    touch(svi::k_phi<G, V, L, false, true>);
    touch(svi::k_phi<G, V, L, true, false>);
    touch(svi::k_phi<G, V, L, true, true>);
    touch(svi::k_node<G, V>);
    touch(svi::k_s3<G, V>);
    touch(svi::k_lambda<L>);
    touch(svi::k_refresh<G, V, L, true>);
    touch(svi::k_refresh<G, V, L, false>);
    touch(svi::k_heldout<G, V, svi::LocalRows>);
    touch(svi::k_heldout<G, V, svi::Peers>);
  }
  static Ops ops() {
    Ops o{phi, node, s3, lambda, refresh, heldout, max_blocks_node, max_blocks_s3, G, V, L ? 1 : 0};
    o.preload = preload;
    return o;
  }
};

// K > 1024: one block per work item, columns strided over its threads (svi_ls_wide.cuh); always the log domain
struct WideTile {
  static constexpr int T = (int)svi::kWideT;
  static_assert(T == kThreads, "svi_ls_create sizes the persistent grids of a tile from kThreads / Ops::lanes");
  static void phi(const Params &P, cudaStream_t st, bool sparse, bool comm, uint32_t s0, uint32_t s1, uint32_t t0,
                  uint32_t t1, uint32_t pub) {
    const uint64_t cnt = (uint64_t)(s1 - s0) + (t1 - t0);
    if (!cnt) return;
    const uint32_t blocks = (uint32_t)cnt;   // (a segment owns a K-row of `part`: 2^31 of them do not fit any device)
    if (sparse && comm) svi::k_phi_wide<true, true><<<blocks, T, 0, st>>>(P, s0, s1, t0, t1, pub);
    else if (sparse) svi::k_phi_wide<true, false><<<blocks, T, 0, st>>>(P, s0, s1, t0, t1, pub);
    else if (comm) svi::k_phi_wide<false, true><<<blocks, T, 0, st>>>(P, s0, s1, t0, t1, pub);
    else svi::k_phi_wide<false, false><<<blocks, T, 0, st>>>(P, s0, s1, t0, t1, pub);
  }
  static void node(const Params &P, cudaStream_t st, uint32_t blocks) {
    svi::k_node_wide<<<blocks, T, 0, st>>>(P, svi::wide_cap(P.ld));
  }
  static void s3(const Params &P, cudaStream_t st, uint32_t blocks) {
    svi::k_s3_wide<<<blocks, T, 0, st>>>(P, svi::wide_cap(P.ld));
  }
  static void lambda(const Params &P, cudaStream_t st, int annealing, int update) {
    svi::k_lambda<true><<<1, kThreads, 0, st>>>(P, annealing, update);
  }
  static void refresh(const Params &P, cudaStream_t st, bool from_gacc) {
    const uint32_t rows = P.node_end - P.node_begin;
    if (!rows) return;
    if (from_gacc) svi::k_refresh_wide<true><<<rows, T, 0, st>>>(P);
    else svi::k_refresh_wide<false><<<rows, T, 0, st>>>(P);
  }
  static void heldout(const Params &P, cudaStream_t st, uint64_t np, const uint32_t *p, const uint32_t *q,
                      const uint8_t *y, double eps, double *out, unsigned long long *bad, const svi::Peers *peer_rows) {
    if (!np) return;
    const uint32_t blocks = (uint32_t)std::min<uint64_t>(np, 1u << 20);   // grid-stride over the pairs
    if (peer_rows) svi::k_heldout_wide<svi::Peers><<<blocks, T, 0, st>>>(P, *peer_rows, np, p, q, y, eps, out, bad);
    else svi::k_heldout_wide<svi::LocalRows><<<blocks, T, 0, st>>>(P, svi::LocalRows{P.gamma}, np, p, q, y, eps, out, bad);
  }
  // persistent grids: a block keeps a [3][cap] (node pass) or [cap] (s3 sweep) slot of column partials in global
  // memory, so the grid is kept small -- 2 / 8 blocks per SM (K = 65 535: 3.7 GB of partials)
  static int max_blocks_node(int sms) { return 2 * sms; }
  static int max_blocks_s3(int sms) { return 8 * sms; }
  static void preload() {
    touch(svi::k_phi_wide<false, false>);
    touch(svi::k_phi_wide<false, true>);
    touch(svi::k_phi_wide<true, false>);
    touch(svi::k_phi_wide<true, true>);
    touch(svi::k_node_wide);
    touch(svi::k_s3_wide);
    touch(svi::k_lambda<true>);
    touch(svi::k_refresh_wide<true>);
    touch(svi::k_refresh_wide<false>);
    touch(svi::k_heldout_wide<svi::LocalRows>);
    touch(svi::k_heldout_wide<svi::Peers>);
  }
  static Ops ops(uint32_t k) {
    const uint32_t ld = (k + 3u) & ~3u;
    // lanes * vec * 2 = the slot width k_reduce_kpart is told (svi::wide_cap)
    Ops o{phi, node, s3, lambda, refresh, heldout, max_blocks_node, max_blocks_s3, T, (int)(svi::wide_cap(ld) / (2u * T)), 1};
    o.preload = preload;
    return o;
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
  static void phi(const Params &P, cudaStream_t st, bool sparse, bool comm, uint32_t s0, uint32_t s1, uint32_t t0,
                  uint32_t t1, uint32_t pub) {
    const uint64_t cnt = (uint64_t)(s1 - s0) + (t1 - t0);
    if (!cnt) return;
    const uint32_t blocks = (uint32_t)((cnt * G + T - 1) / T);
    if (sparse && comm) svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, true, true><<<blocks, T, kSmem, st>>>(P, s0, s1, t0, t1, pub);
    else if (sparse) svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, true, false><<<blocks, T, kSmem, st>>>(P, s0, s1, t0, t1, pub);
    else if (comm) svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, false, true><<<blocks, T, kSmem, st>>>(P, s0, s1, t0, t1, pub);
    else svi::k_sweep_ring<G, V, R, T, MINB, Sweep::Phi, false, false><<<blocks, T, kSmem, st>>>(P, s0, s1, t0, t1, pub);
  }
  static void s3(const Params &P, cudaStream_t st, uint32_t blocks) {
    if (P.nseg <= P.nseg_lo) return;
    svi::k_sweep_ring<G, V, R, T, MINB, Sweep::S3, false, false><<<blocks, T, kSmem, st>>>(P, P.nseg_lo, P.nseg, 0u, 0u, 0u);
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
  else if (k <= 65535) *o = WideTile::ops(k);   // the reference's limit: communities are uint16_t (src/env.hh:37)
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

// Layout of the exchange arena (identical on every shard of a run: it depends on n, ld, words only)
struct ArenaLayout {
  size_t b, mphi, gamma, conv[2], active, abits, mbits, kx, flags, bytes;
  ArenaLayout() = default;
  ArenaLayout(uint32_t n, uint32_t ld, uint32_t words) {
    size_t at = 0;
    auto take = [&](size_t nbytes) { const size_t o = at; at = (at + nbytes + 255) & ~(size_t)255; return o; };
    b = take((size_t)n * ld * 8);
    mphi = take((size_t)n * ld * 8);
    gamma = take((size_t)n * ld * 8);
    conv[0] = take((size_t)n * 4);
    conv[1] = take((size_t)n * 4);
    active = take((size_t)n * 4);
    abits = take((size_t)n * words * 4);
    mbits = take((size_t)n * words * 4);
    kx = take((size_t)2 * 2 * svi::kMaxWorld * 4 * ld * 8);
    flags = take((size_t)svi::kMaxWorld * svi::kFlagKinds * 4);
    bytes = std::max<size_t>(at, 256);
  }
};

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
  // exchange arena: b, mphi, gamma, converged x2, active, active bits, membership bits, K-vector slots, flags
  unsigned char *d_arena = nullptr;
  ArenaLayout lay;
  double *d_tl = nullptr, *d_b = nullptr, *d_mphi = nullptr, *d_gamma = nullptr, *d_gacc = nullptr;
  double *d_part = nullptr, *d_kvec = nullptr, *d_kpart = nullptr, *d_lambda = nullptr, *d_eb = nullptr;
  double *d_scale = nullptr, *d_stage = nullptr;
  uint32_t *d_conv2[2] = {nullptr, nullptr}, *d_active = nullptr, *d_abits = nullptr, *d_mbits = nullptr;
  int cur = 0;                   // d_conv2[cur] = `converged` as this iteration's sweeps see it
  bool conv_pending = false;     // the refresh of this iteration has written d_conv2[cur ^ 1]; flip at its end
  size_t stage_elems = 0;
  // segment ranges of the local nodes (host copies) and the prefix of their half-edge counts: chunk planning
  std::vector<uint32_t> nlo, nup;
  std::vector<uint64_t> he_prefix;
  // ---- multi-GPU (svi_ls_peer_*, svi_ls_mg_step) ----
  svi::Peers peers{};
  bool mg = false, mg_ipc = false, share_gamma = false;
  std::vector<uint32_t> bounds;          // node blocks of all shards
  std::vector<uint32_t> chunk_nodes;     // the own block cut into pipeline chunks: chunk c = [chunk_nodes[c], chunk_nodes[c+1])
  cudaStream_t side = nullptr, own_main = nullptr, aux = nullptr, up = nullptr;   // aux: node passes, up: "up" segments
  cudaEvent_t ev_phi = nullptr, ev_node = nullptr, ev_up = nullptr;
  cudaEvent_t ev_chunk = nullptr, ev_refresh = nullptr, ev_side = nullptr;
  // how rows travel to the peers: 0 = one copy-engine transfer per peer on the side stream (default: measured best
  // at 8 GPUs, 12.2 ms per iteration at config 4), 1 = the same transfers fanned out over one stream per peer (12.7-13.3),
  // 2 = a small SM kernel that reads a row once and stores it to every peer (12.9; fast pushes, but its blocks displace
  // the persistent s3 sweep's)
  int push_mode = 0;
  uint32_t push_blocks = 32;
  cudaStream_t fan[svi::kMaxWorld] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[svi::kMaxWorld] = {};
  uint32_t epoch = 0, gamma_epoch = 0;
  bool gamma_wait = false;       // svi_ls_mg_publish_gamma ran: the next svi_ls_get_state awaits the peers' rows
  uint32_t *d_mg_err = nullptr;
  // svi_ls_step as a CUDA graph, one per (active-set branch, tally, annealing, converged-buffer parity): the
  // iteration is ~10 dependent launches, which at small sizes (config 2: 17 903 nodes, K = 20) cost more than the kernels
  cudaGraphExec_t graphs[16] = {};
  cudaStream_t cap = nullptr;
  bool use_graph = true;
  // the refresh (FP64-bound: digamma, exp) beside the s3 sweep (HBM-bound, 27 % issue utilisation) on a second stream;
  // the sweep then runs with one block per SM fewer so that a refresh block fits next to it
  bool overlap_refresh = false;
  cudaStream_t rf = nullptr, cap_rf = nullptr;
  cudaEvent_t ev_fork_rf = nullptr, ev_join_rf = nullptr;
  // optional per-phase timing of svi_ls_mg_step (svi_ls_mg_timing): events on the main stream, ring of steps
  static constexpr int kTimedSteps = 32, kMarks = 9, kSideMarks = 4;
  bool timing = false;
  cudaEvent_t tev[kTimedSteps][kMarks] = {};
  cudaEvent_t sev[kTimedSteps][kSideMarks] = {};   // side stream: first mphi push .. M flag, first b push .. B flag
  uint32_t tsteps = 0;
  uint64_t mg_timeout_ns = 20ull * 1000000000ull;
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
  if (h->mg_ipc)
    for (uint32_t r = 0; r < h->peers.world; ++r)
      if (r != h->peers.rank && h->peers.arena[r]) cudaIpcCloseMemHandle(h->peers.arena[r]);
  void *ptrs[] = {h->d_col, h->d_seg_node, h->d_seg_beg, h->d_seg_cnt, h->d_seg_nnc, h->d_node_seg_lo,
                  h->d_node_seg_up, h->d_conv_dirty, h->d_tl, h->d_arena, h->d_gacc, h->d_part,
                  h->d_kvec, h->d_kpart, h->d_lambda, h->d_eb, h->d_scale, h->d_stage, h->d_mg_err};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (auto &row : h->tev)
    for (cudaEvent_t ev : row)
      if (ev) cudaEventDestroy(ev);
  for (auto &row : h->sev)
    for (cudaEvent_t ev : row)
      if (ev) cudaEventDestroy(ev);
  for (cudaGraphExec_t g : h->graphs)
    if (g) cudaGraphExecDestroy(g);
  if (h->rf) cudaStreamDestroy(h->rf);
  if (h->cap_rf) cudaStreamDestroy(h->cap_rf);
  if (h->ev_fork_rf) cudaEventDestroy(h->ev_fork_rf);
  if (h->ev_join_rf) cudaEventDestroy(h->ev_join_rf);
  if (h->cap) cudaStreamDestroy(h->cap);
  for (cudaStream_t st : h->fan)
    if (st) cudaStreamDestroy(st);
  for (cudaEvent_t ev : h->ev_join)
    if (ev) cudaEventDestroy(ev);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_chunk) cudaEventDestroy(h->ev_chunk);
  if (h->ev_refresh) cudaEventDestroy(h->ev_refresh);
  if (h->ev_side) cudaEventDestroy(h->ev_side);
  if (h->ev_phi) cudaEventDestroy(h->ev_phi);
  if (h->ev_node) cudaEventDestroy(h->ev_node);
  if (h->ev_up) cudaEventDestroy(h->ev_up);
  if (h->up) cudaStreamDestroy(h->up);
  if (h->aux) cudaStreamDestroy(h->aux);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->own_main) cudaStreamDestroy(h->own_main);
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

static void mg_await_rows(svi_ls *h);
int svi_ls_mg_error(svi_ls *h);

const char *svi_ls_last_error(void) { return g_err; }
int svi_ls_abi_version(void) { return SVI_LS_ABI_VERSION; }

int svi_ls_create(const svi_ls_config *cfg, const uint32_t *links, const double *tl, svi_ls **out) {
  if (!cfg || !out || (!links && cfg->nlinks)) return fail(SVI_ERR_INVALID, "svi_ls_create: null argument");
  *out = nullptr;
  if (cfg->n == 0 || cfg->k == 0) return fail(SVI_ERR_INVALID, "svi_ls_create: n and k must be positive");
  if (cfg->node_begin > cfg->node_end || cfg->node_end > cfg->n)
    return fail(SVI_ERR_INVALID, "svi_ls_create: bad shard [%u,%u) for n=%u", cfg->node_begin, cfg->node_end, cfg->n);
  Ops ops;
  if (!pick_ops(cfg->k, &ops)) return fail(SVI_ERR_UNSUPPORTED, "svi_ls_create: k=%u not supported (max 65535)", cfg->k);

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
  // persistent grids: no more blocks than keep every lane group busy with a few items (each block ends in a block-wide
  // reduction of its column sums, which at small sizes would otherwise outweigh the rows it summed)
  const int64_t node_groups = kThreads / ops.lanes;
  h->blocks_node = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(ops.max_blocks_node(h->sms),
                                                                   ((int64_t)nlocal + node_groups * 8 - 1) / (node_groups * 8)));
  if (ops.s3_ring)
    h->blocks_s3 = (uint32_t)std::max<int64_t>(
        1, std::min<int64_t>(ops.max_blocks_s3_ring(h->sms, ld),
                             ((int64_t)nseg3 * ops.s3_lanes + ops.s3_threads * 4 - 1) / (ops.s3_threads * 4)));
  else
    h->blocks_s3 = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(ops.max_blocks_s3(h->sms),
                                                                   ((int64_t)nseg3 * ops.lanes + kThreads * 4 - 1) / (kThreads * 4)));
  if (const char *sb = getenv("SVI_LS_S3_BLOCKS_PER_SM"))
    if (atoi(sb) > 0) h->blocks_s3 = std::min<uint32_t>(h->blocks_s3, (uint32_t)atoi(sb) * (uint32_t)h->sms);
  if (ops.prepare_phi_ring) ops.prepare_phi_ring();   // (setting the attribute loads the ring kernels too)
  if (ops.prepare_s3_ring) ops.prepare_s3_ring();
  if (ops.preload) ops.preload();
  touch(svi::k_reduce_kpart); touch(svi::k_scale); touch(svi::k_partition); touch(svi::k_fill);
  touch(svi::k_pad_rows); touch(svi::k_unpad_rows);
  touch(svi::k_mg_signal); touch(svi::k_mg_wait); touch(svi::k_mg_kx_push); touch(svi::k_mg_kx_sum);
  touch(svi::k_mg_or_rows); touch(svi::k_mg_push<uint4>); touch(svi::k_mg_push<uint32_t>);
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
  h->lay = ArenaLayout(n, ld, words);
  A(dalloc(&h->d_arena, h->lay.bytes, &tot));
  if (e == cudaSuccess) {
    h->d_b = reinterpret_cast<double *>(h->d_arena + h->lay.b);
    h->d_mphi = reinterpret_cast<double *>(h->d_arena + h->lay.mphi);
    h->d_gamma = reinterpret_cast<double *>(h->d_arena + h->lay.gamma);
    h->d_conv2[0] = reinterpret_cast<uint32_t *>(h->d_arena + h->lay.conv[0]);
    h->d_conv2[1] = reinterpret_cast<uint32_t *>(h->d_arena + h->lay.conv[1]);
    h->d_active = reinterpret_cast<uint32_t *>(h->d_arena + h->lay.active);
    h->d_abits = reinterpret_cast<uint32_t *>(h->d_arena + h->lay.abits);
    h->d_mbits = reinterpret_cast<uint32_t *>(h->d_arena + h->lay.mbits);
  }
  A(dalloc(&h->d_mg_err, 1, &tot));
  A(dalloc(&h->d_gacc, nld, &tot));
  A(dalloc(&h->d_part, (size_t)nseg * ld, &tot));
  A(dalloc(&h->d_kvec, 4 * (size_t)ld, &tot));
  A(dalloc(&h->d_kpart, (size_t)std::max<size_t>(h->kpart_blocks, (size_t)kMaxChunks * h->blocks_node) * 3 * cap, &tot));
  A(dalloc(&h->d_lambda, 2 * (size_t)k, &tot));
  A(dalloc(&h->d_eb, ld, &tot));
  A(dalloc(&h->d_scale, ld, &tot));
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
  P.node_begin = nb; P.node_end = ne; P.shard_begin = nb;
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
  P.conv = h->d_conv2[0]; P.conv_next = h->d_conv2[1]; P.active = h->d_active; P.abits = h->d_abits; P.mbits = h->d_mbits;
  h->nlo = std::move(nlo);
  h->nup = std::move(nup);
  if (const char *ng = getenv("SVI_LS_NO_GRAPH")) h->use_graph = !(ng[0] == '1');
  if (const char *ov = getenv("SVI_LS_OVERLAP_REFRESH")) h->overlap_refresh = ov[0] == '1';
  if (h->overlap_refresh) {
    if (cudaStreamCreateWithFlags(&h->rf, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork_rf, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join_rf, cudaEventDisableTiming) != cudaSuccess)
      h->overlap_refresh = false;
  }
  h->he_prefix.assign(nlocal + 1, 0);
  for (uint32_t v = 0; v < nlocal; ++v) h->he_prefix[v + 1] = h->he_prefix[v] + deg_lo[v] + deg_up[v];
  *out = h;
  return SVI_OK;
}

void svi_ls_destroy(svi_ls *h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->side) cudaStreamSynchronize(h->side);
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
  if (h->side) CK(cudaStreamSynchronize(h->side));
  if (h->mg) return svi_ls_mg_error(h);
  return SVI_OK;
}

int svi_ls_set_state(svi_ls *h, const double *gamma, const double *lambda) {
  if (!h || !gamma || !lambda) return fail(SVI_ERR_INVALID, "svi_ls_set_state: null argument");
  DeviceGuard guard(h->device);
  const Params &P = h->P;
  const size_t nk = (size_t)P.n * P.k;
  mg_await_rows(h);   // (rows the peers pushed for the abandoned state must not land on top of the new one)
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
  if (gamma && h->share_gamma) mg_await_rows(h);   // the other shards' rows of the last iteration
  if (gamma && h->gamma_wait) {                     // ... or of svi_ls_mg_publish_gamma
    svi::k_mg_wait<<<1, 32, 0, h->stream>>>(h->peers, svi::FLAG_G, h->gamma_epoch, h->d_mg_err, h->mg_timeout_ns);
    h->gamma_wait = false;
  }
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
  CK(cudaMemcpyAsync(h->d_conv2[h->cur], converged, (size_t)h->P.n * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->force_partition = true;   // arbitrary flags (a resume may even clear some): rebuild every segment's partition
  return SVI_OK;
}

int svi_ls_get_converged(svi_ls *h, uint32_t *converged, uint32_t *active_comms) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  mg_await_rows(h);   // sharded: the other shards' flags of the last iteration
  if (converged)
    CK(cudaMemcpyAsync(converged, h->d_conv2[h->cur], (size_t)h->P.n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  if (active_comms)
    CK(cudaMemcpyAsync(active_comms, h->d_active, (size_t)h->P.n * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                       h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
}

}  // extern "C"

namespace {

void flip_converged(svi_ls *h) {
  h->cur ^= 1;
  h->P.conv = h->d_conv2[h->cur];
  h->P.conv_next = h->d_conv2[h->cur ^ 1];
  h->conv_pending = false;
}

// start of an iteration: re-partition the neighbour lists if nodes converged, clear the link-community bits
int begin_iteration(svi_ls *h, cudaStream_t st, int write_comm) {
  const Params &P = h->P;
  if (h->conv_pending) flip_converged(h);   // (an iteration abandoned between its refresh and its lambda phase)
  // the ring sweeps read neighbour lists partitioned by the converged flags (svi_ls_ring.cuh: k_partition); nodes
  // converge in k_refresh, at the end of the previous iteration
  if ((h->ops.phi_ring || h->ops.s3_ring) && P.nseg) {
    const bool force = h->force_partition || h->partition_every_sweep;
    svi::k_partition<<<(uint32_t)std::min<uint64_t>((P.nseg + 7) / 8, (uint64_t)h->sms * 8), 256, 0, st>>>(h->P, force);
    CK(cudaMemsetAsync(h->d_conv_dirty, 0, sizeof(uint32_t), st));
    h->force_partition = false;
  }
  if (write_comm)  // _communities.clear(); _fmap.zero()  (src/linksampling.cc:584-587)
    CK(cudaMemsetAsync(h->d_mbits, 0, (size_t)P.n * P.words * sizeof(uint32_t), st));
  return SVI_OK;
}

// phi sweep over the "lo" segments [lo0, lo1) (on stream st) and the "up" segments [up0, up1) (on stream st_up)
void launch_phi(svi_ls *h, cudaStream_t st, cudaStream_t st_up, uint32_t iter, int write_comm, uint32_t lo0, uint32_t lo1,
                uint32_t up0, uint32_t up1) {
  const Params &P = h->P;
  auto phi = h->ops.phi_ring ? h->ops.phi_ring : h->ops.phi;
  const bool sparse = iter > 1000 && P.k_div10 > 0;
  if (!write_comm) {
    if (st == st_up) phi(P, st, sparse, false, lo0, lo1, up0, up1, 0);
    else { phi(P, st, sparse, false, lo0, lo1, 0, 0, 0); phi(P, st_up, sparse, false, up0, up1, 0, 0, 0); }
  } else if (h->shard && !h->mg) {
    // a shard driven phase by phase (svi_ls_phase_*) holds the membership words of its own nodes only: the arg-max
    // is taken on both sides of a link
    phi(P, st, sparse, true, lo0, lo1, up0, up1, 0);
  } else {
    // one arg-max per LINK (src/linksampling.cc:704-717 sets fmap[p] and fmap[q] from one max_k): the owner's side
    // computes it and sets both endpoints' bits; the "lo" segments run the kernel without the tally.  (Two launches:
    // one launch with a per-warp switch measured 58.5 ms against 54.8 ms at config 4.)  svi_ls_mg_step: the bit of a
    // neighbour owned by another shard lands in the local replica of the membership words and is merged after the
    // sweep, k_mg_or_rows.
    phi(P, st, sparse, false, lo0, lo1, 0, 0, 0);
    phi(P, st_up, sparse, true, up0, up1, 0, 0, 1);
  }
}

// mean indicators of the nodes [v0, v1) -> block partials of chunk slot `c`
void launch_node(svi_ls *h, cudaStream_t st, uint32_t v0, uint32_t v1, uint32_t c) {
  Params P = h->P;
  P.node_begin = v0;
  P.node_end = v1;
  P.kpart = h->d_kpart + (size_t)c * h->blocks_node * 3 * (2 * h->ops.lanes * h->ops.vec);
  h->ops.node(P, st, h->blocks_node);
}

void launch_s3(svi_ls *h, cudaStream_t st) {
  const Params &P = h->P;
  uint32_t cap3 = 2 * h->ops.lanes * h->ops.vec;
  if (h->ops.s3_ring) {
    h->ops.s3_ring(P, st, h->blocks_s3);
    cap3 = 2 * h->ops.s3_lanes * h->ops.s3_vec;
  } else {
    h->ops.s3(P, st, h->blocks_s3);
  }
  svi::k_reduce_kpart<<<svi::reduce_kpart_blocks(1, P.ld), 256, 0, st>>>(h->d_kpart, h->blocks_s3, 1, cap3, h->d_kvec + 3 * (size_t)P.ld, P.ld);
}

}  // namespace

extern "C" {

int svi_ls_phase_phi(svi_ls *h, uint32_t iter, int write_comm) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  const Params &P = h->P;
  int rc = begin_iteration(h, h->stream, write_comm);
  if (rc) return rc;
  launch_phi(h, h->stream, h->stream, iter, write_comm, 0, P.nseg_lo, P.nseg_lo, P.nseg);
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_phase_node(svi_ls *h) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  const Params &P = h->P;
  launch_node(h, h->stream, P.node_begin, P.node_end, 0);
  svi::k_reduce_kpart<<<svi::reduce_kpart_blocks(3, P.ld), 256, 0, h->stream>>>(h->d_kpart, h->blocks_node, 3, 2 * h->ops.lanes * h->ops.vec,
                                               h->d_kvec, P.ld);
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_phase_s3(svi_ls *h) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  // reads h->P.conv, the flags of the START of this iteration even when the refresh already ran: the reference's
  // s3 loop (:731-746) precedes prune (:761), whose result lives in P.conv_next until the iteration ends
  launch_s3(h, h->stream);
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_phase_finish(svi_ls *h, int annealing) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  h->ops.lambda(h->P, h->stream, annealing, 1);
  h->ops.refresh(h->P, h->stream, true);
  CK(cudaGetLastError());
  flip_converged(h);
  return SVI_OK;
}

int svi_ls_phase_refresh(svi_ls *h, int annealing) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  svi::k_scale<<<1, 256, 0, h->stream>>>(h->P, annealing);
  h->ops.refresh(h->P, h->stream, true);
  h->conv_pending = true;
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_phase_lambda(svi_ls *h, int annealing) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  h->ops.lambda(h->P, h->stream, annealing, 1);
  if (h->conv_pending) flip_converged(h);
  CK(cudaGetLastError());
  return SVI_OK;
}

// the whole iteration on stream `st` (== phase_phi; phase_node; phase_s3; phase_finish without the host-side flip)
static int enqueue_step(svi_ls *h, cudaStream_t st, uint32_t iter, int annealing, int write_comm) {
  const Params &P = h->P;
  int rc = begin_iteration(h, st, write_comm);
  if (rc) return rc;
  launch_phi(h, st, st, iter, write_comm, 0, P.nseg_lo, P.nseg_lo, P.nseg);
  launch_node(h, st, P.node_begin, P.node_end, 0);
  svi::k_reduce_kpart<<<svi::reduce_kpart_blocks(3, P.ld), 256, 0, st>>>(h->d_kpart, h->blocks_node, 3, 2 * h->ops.lanes * h->ops.vec, h->d_kvec, P.ld);
  if (h->overlap_refresh) {
    // fork: rescale + refresh (needs `sum` only; writes gamma, b, conv[cur^1], masks) on the second stream, the s3
    // sweep (reads mphi, conv[cur]) on this one; join before lambda
    cudaStream_t rf = st == h->cap ? h->cap_rf : h->rf;
    CK(cudaEventRecord(h->ev_fork_rf, st));
    CK(cudaStreamWaitEvent(rf, h->ev_fork_rf, 0));
    svi::k_scale<<<1, 256, 0, rf>>>(P, annealing);
    h->ops.refresh(P, rf, true);
    CK(cudaEventRecord(h->ev_join_rf, rf));
    launch_s3(h, st);
    CK(cudaStreamWaitEvent(st, h->ev_join_rf, 0));
    h->ops.lambda(P, st, annealing, 1);
  } else {
    launch_s3(h, st);
    h->ops.lambda(P, st, annealing, 1);
    h->ops.refresh(P, st, true);
  }
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_step(svi_ls *h, uint32_t iter, int annealing, int write_comm) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  const bool sparse = iter > 1000 && h->P.k_div10 > 0;
  int rc;
  if (!h->use_graph || h->force_partition || h->partition_every_sweep || h->conv_pending || h->mg) {
    if ((rc = enqueue_step(h, h->stream, iter, annealing, write_comm))) return rc;
    flip_converged(h);
    return SVI_OK;
  }
  const int key = (sparse ? 1 : 0) | (write_comm ? 2 : 0) | (annealing ? 4 : 0) | (h->cur ? 8 : 0);
  if (!h->graphs[key]) {
    if (!h->cap) CK(cudaStreamCreateWithFlags(&h->cap, cudaStreamNonBlocking));
    if (h->overlap_refresh && !h->cap_rf) CK(cudaStreamCreateWithFlags(&h->cap_rf, cudaStreamNonBlocking));
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(h->cap, cudaStreamCaptureModeThreadLocal));
    rc = enqueue_step(h, h->cap, sparse ? 1001u : 0u, annealing, write_comm);
    const cudaError_t e = cudaStreamEndCapture(h->cap, &g);
    if (rc) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    if (e != cudaSuccess) return fail(SVI_ERR_CUDA, "svi_ls_step: graph capture failed: %s", cudaGetErrorString(e));
    const cudaError_t ei = cudaGraphInstantiate(&h->graphs[key], g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) return fail(SVI_ERR_CUDA, "svi_ls_step: cudaGraphInstantiate: %s", cudaGetErrorString(ei));
  }
  CK(cudaGraphLaunch(h->graphs[key], h->stream));
  flip_converged(h);
  return SVI_OK;
}

// ---- multi-GPU over peer memory (svi_ls_mg.cuh) ------------------------------------------------------------------

struct svi_ls_peer_blob_layout {
  uint64_t magic;
  uint32_t n, ld, words, device;
  uint64_t arena_bytes, pid;
  cudaIpcMemHandle_t mem;
};
static const uint64_t kBlobMagic = 0x5356494c53424c32ull;

size_t svi_ls_peer_blob_bytes(void) { return sizeof(svi_ls_peer_blob_layout); }

int svi_ls_peer_export(svi_ls *h, void *blob, size_t blob_bytes) {
  if (!h || !blob || blob_bytes < sizeof(svi_ls_peer_blob_layout)) return fail(SVI_ERR_INVALID, "svi_ls_peer_export: bad argument");
  DeviceGuard guard(h->device);
  svi_ls_peer_blob_layout b;
  memset(&b, 0, sizeof b);
  b.magic = kBlobMagic;
  b.n = h->P.n; b.ld = h->P.ld; b.words = h->P.words; b.device = (uint32_t)h->device;
  b.arena_bytes = h->lay.bytes;
  b.pid = (uint64_t)getpid();
  CK(cudaIpcGetMemHandle(&b.mem, h->d_arena));
  memcpy(blob, &b, sizeof b);
  return SVI_OK;
}

static int mg_push_rows(svi_ls *h, cudaStream_t st, size_t off, size_t row_bytes, uint32_t v0, uint32_t v1);

static int mg_finish_attach(svi_ls *h, uint32_t world, uint32_t rank, const uint32_t *bounds, uint32_t chunks) {
  if (bounds[0] != 0 || bounds[world] != h->P.n || bounds[rank] != h->P.node_begin || bounds[rank + 1] != h->P.node_end)
    return fail(SVI_ERR_INVALID, "svi_ls_peer_attach: bounds do not match the handle's node block [%u,%u)", h->P.node_begin,
                h->P.node_end);
  h->bounds.assign(bounds, bounds + world + 1);
  h->peers.world = world;
  h->peers.rank = rank;
  h->peers.arena[rank] = h->d_arena;
  h->peers.flags_off = h->lay.flags;
  h->peers.kx_off = h->lay.kx;
  h->peers.gamma_off = h->lay.gamma;
  for (uint32_t r = 0; r <= world; ++r) h->peers.bounds[r] = bounds[r];
  h->peers.kx_stride = 4 * h->P.ld;
  if (!h->stream) {   // the legacy default stream would serialise the shards of one process: own streams
    CK(cudaStreamCreateWithFlags(&h->own_main, cudaStreamNonBlocking));
    h->stream = h->own_main;
  }
  if (!h->side) {   // highest priority: its few push blocks / flag kernels must not queue behind a whole sweep
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));
  }
  if (!h->aux) {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->aux, cudaStreamNonBlocking, hi));
    CK(cudaEventCreateWithFlags(&h->ev_phi, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_node, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&h->up, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_up, cudaEventDisableTiming));
  }
  if (const char *pm = getenv("SVI_LS_MG_PUSH")) h->push_mode = !strcmp(pm, "ce") ? 0 : !strcmp(pm, "ce_multi") ? 1 : 2;
  if (const char *pb = getenv("SVI_LS_MG_PUSH_BLOCKS")) h->push_blocks = (uint32_t)std::max(1, atoi(pb));
  if (h->push_mode == 1 && !h->ev_fork) {
    CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (uint32_t r = 0; r < world; ++r) {
      if (r == rank) continue;
      CK(cudaStreamCreateWithFlags(&h->fan[r], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&h->ev_join[r], cudaEventDisableTiming));
    }
  }
  if (!h->ev_chunk) CK(cudaEventCreateWithFlags(&h->ev_chunk, cudaEventDisableTiming));
  if (!h->ev_refresh) CK(cudaEventCreateWithFlags(&h->ev_refresh, cudaEventDisableTiming));
  if (!h->ev_side) CK(cudaEventCreateWithFlags(&h->ev_side, cudaEventDisableTiming));
  // pipeline chunks of the own block, balanced by half-edges
  const char *ce = getenv("SVI_LS_MG_CHUNKS");
  uint32_t c = chunks ? chunks : (ce ? (uint32_t)atoi(ce) : 4u);
  c = std::max(1u, std::min(c, kMaxChunks));
  if (world == 1) c = 1;
  h->chunk_nodes.assign(1, h->P.node_begin);
  const uint64_t total = h->he_prefix[h->nlocal];
  // geometrically shrinking chunks (ratio q): what stays exposed is the LAST chunk's push
  const char *qe = getenv("SVI_LS_MG_CHUNK_RATIO");
  const double q = qe && atof(qe) > 0 ? atof(qe) : 0.6;
  double wsum = 0, w = 1;
  for (uint32_t i = 0; i < c; ++i, w *= q) wsum += w;
  double acc = 0;
  w = 1;
  for (uint32_t i = 1; i < c; ++i) {
    acc += w / wsum;
    w *= q;
    const uint64_t target = (uint64_t)((double)total * acc);
    uint32_t v = (uint32_t)(std::lower_bound(h->he_prefix.begin(), h->he_prefix.end(), target) - h->he_prefix.begin());
    v = std::min(v, h->nlocal);
    h->chunk_nodes.push_back(std::max(h->chunk_nodes.back(), h->P.node_begin + v));
  }
  h->chunk_nodes.push_back(h->P.node_end);
  // first use of the copy / memset paths between the arenas happens here, not beside a waiting kernel
  for (uint32_t r = 0; r < world; ++r)
    if (r != rank) CK(cudaMemcpyAsync(h->d_mg_err, h->peers.arena[r] + h->lay.flags, sizeof(uint32_t), cudaMemcpyDeviceToDevice, h->side));
  CK(cudaMemsetAsync(h->d_mg_err, 0, sizeof(uint32_t), h->side));
  CK(cudaStreamSynchronize(h->side));
  h->epoch = 0;
  h->mg = true;
  h->partition_every_sweep = false;   // the shards all-reduce their "newly converged" flags (svi_ls_mg_step)
  const char *te = getenv("SVI_LS_MG_TIMEOUT_S");
  if (te && atof(te) > 0) h->mg_timeout_ns = (uint64_t)(atof(te) * 1e9);
  return SVI_OK;
}

int svi_ls_peer_attach(svi_ls *h, uint32_t world, uint32_t rank, const uint32_t *bounds, const void *blobs,
                       uint32_t chunks) {
  if (!h || !bounds || !blobs || world < 1 || world > svi::kMaxWorld || rank >= world)
    return fail(SVI_ERR_INVALID, "svi_ls_peer_attach: bad argument (world <= %u)", svi::kMaxWorld);
  if (h->mg) return fail(SVI_ERR_INVALID, "svi_ls_peer_attach: already attached");
  DeviceGuard guard(h->device);
  const svi_ls_peer_blob_layout *bl = reinterpret_cast<const svi_ls_peer_blob_layout *>(blobs);
  for (uint32_t r = 0; r < world; ++r) {
    if (bl[r].magic != kBlobMagic || bl[r].n != h->P.n || bl[r].ld != h->P.ld || bl[r].arena_bytes != h->lay.bytes)
      return fail(SVI_ERR_INVALID, "svi_ls_peer_attach: blob %u does not describe a shard of this problem", r);
    if (r == rank) continue;
    if (bl[r].pid == (uint64_t)getpid())
      return fail(SVI_ERR_INVALID, "svi_ls_peer_attach: shard %u lives in this process; use svi_ls_peer_attach_local", r);
    void *ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, bl[r].mem, cudaIpcMemLazyEnablePeerAccess));
    h->peers.arena[r] = static_cast<unsigned char *>(ptr);
  }
  h->mg_ipc = true;
  return mg_finish_attach(h, world, rank, bounds, chunks);
}

int svi_ls_peer_attach_local(svi_ls *h, uint32_t world, uint32_t rank, const uint32_t *bounds, svi_ls *const *handles,
                             uint32_t chunks) {
  if (!h || !bounds || !handles || world < 1 || world > svi::kMaxWorld || rank >= world || handles[rank] != h)
    return fail(SVI_ERR_INVALID, "svi_ls_peer_attach_local: bad argument (world <= %u)", svi::kMaxWorld);
  if (h->mg) return fail(SVI_ERR_INVALID, "svi_ls_peer_attach_local: already attached");
  DeviceGuard guard(h->device);
  for (uint32_t r = 0; r < world; ++r) {
    const svi_ls *o = handles[r];
    if (!o || o->P.n != h->P.n || o->P.ld != h->P.ld || o->lay.bytes != h->lay.bytes)
      return fail(SVI_ERR_INVALID, "svi_ls_peer_attach_local: handle %u is not a shard of this problem", r);
    if (o->device != h->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, h->device, o->device));
      if (!can) return fail(SVI_ERR_UNSUPPORTED, "svi_ls_peer_attach_local: device %d cannot access device %d", h->device, o->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(SVI_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", o->device, cudaGetErrorString(e));
      cudaGetLastError();
    }
    h->peers.arena[r] = o->d_arena;
  }
  return mg_finish_attach(h, world, rank, bounds, chunks);
}

int svi_ls_mg_publish_gamma(svi_ls *h) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  if (!h->mg) return fail(SVI_ERR_INVALID, "svi_ls_mg_publish_gamma: svi_ls_peer_attach[_local] first");
  DeviceGuard guard(h->device);
  ++h->gamma_epoch;
  if (h->peers.world > 1) {
    int rc = mg_push_rows(h, h->stream, h->lay.gamma, (size_t)h->P.ld * 8, h->P.node_begin, h->P.node_end);
    if (rc) return rc;
    svi::k_mg_signal<<<1, 32, 0, h->stream>>>(h->peers, svi::FLAG_G, h->gamma_epoch);
    h->gamma_wait = true;
  }
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_mg_share_gamma(svi_ls *h, int on) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  h->share_gamma = on != 0;
  return SVI_OK;
}

// copy rows [v0, v1) of an arena matrix (row = `row_bytes` bytes at arena offset `off`) into every peer's arena
static int mg_push_rows(svi_ls *h, cudaStream_t st, size_t off, size_t row_bytes, uint32_t v0, uint32_t v1) {
  if (v1 <= v0) return SVI_OK;
  const size_t at = off + (size_t)v0 * row_bytes, bytes = (size_t)(v1 - v0) * row_bytes;
  if (h->push_mode == 2) {
    if (at % 16 == 0 && bytes % 16 == 0) {
      const size_t cnt = bytes / 16;
      const uint32_t blocks = (uint32_t)std::min<size_t>(h->push_blocks, (cnt + 255) / 256);
      svi::k_mg_push<uint4><<<blocks, 256, 0, st>>>(h->peers, at, cnt);
    } else {
      const size_t cnt = bytes / 4;
      const uint32_t blocks = (uint32_t)std::min<size_t>(h->push_blocks, (cnt + 255) / 256);
      svi::k_mg_push<uint32_t><<<blocks, 256, 0, st>>>(h->peers, at, cnt);
    }
    return SVI_OK;
  }
  if (h->push_mode == 1) CK(cudaEventRecord(h->ev_fork, st));
  for (uint32_t d = 1; d < h->peers.world; ++d) {
    const uint32_t r = (h->peers.rank + d) % h->peers.world;
    if (h->push_mode == 1) {
      CK(cudaStreamWaitEvent(h->fan[r], h->ev_fork, 0));
      CK(cudaMemcpyAsync(h->peers.arena[r] + at, h->d_arena + at, bytes, cudaMemcpyDeviceToDevice, h->fan[r]));
      CK(cudaEventRecord(h->ev_join[r], h->fan[r]));
      CK(cudaStreamWaitEvent(st, h->ev_join[r], 0));
    } else {
      CK(cudaMemcpyAsync(h->peers.arena[r] + at, h->d_arena + at, bytes, cudaMemcpyDeviceToDevice, st));
    }
  }
  return SVI_OK;
}

int svi_ls_mg_step(svi_ls *h, uint32_t iter, int annealing, int write_comm) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  if (!h->mg) return fail(SVI_ERR_INVALID, "svi_ls_mg_step: svi_ls_peer_attach[_local] first");
  DeviceGuard guard(h->device);
  const svi::Peers &pr = h->peers;
  cudaStream_t mn = h->stream, side = h->side;
  const Params &P = h->P;
  const uint32_t e = ++h->epoch, par = e & 1u, ld = P.ld;
  const bool multi = pr.world > 1;
  int rc;
  cudaEvent_t *sev = h->timing ? h->sev[h->tsteps % svi_ls::kTimedSteps] : nullptr;
  cudaEvent_t *tev = h->timing ? h->tev[h->tsteps++ % svi_ls::kTimedSteps] : nullptr;
  auto mark = [&](int i) { if (tev) cudaEventRecord(tev[i], mn); };
  auto smark = [&](int i) { if (sev && multi) cudaEventRecord(sev[i], side); };
  mark(0);
  // the peers' rows for this iteration (b, converged, ...) were pushed during their previous refresh
  if (multi && e > 1) svi::k_mg_wait<<<1, 32, 0, mn>>>(pr, svi::FLAG_B, e - 1, h->d_mg_err, h->mg_timeout_ns);
  mark(1);
  if ((rc = begin_iteration(h, mn, write_comm))) return rc;
  // phi sweep + mean indicators chunk by chunk; the finished chunk's mphi rows travel to the peers (side stream,
  // copy engines) beside the next chunk's sweep
  const uint32_t nchunks = (uint32_t)h->chunk_nodes.size() - 1, nb = P.node_begin;
  for (uint32_t c = 0; c < nchunks; ++c) {
    const uint32_t v0 = h->chunk_nodes[c], v1 = h->chunk_nodes[c + 1];
    if (nchunks == 1) {
      launch_phi(h, mn, mn, iter, write_comm, h->nlo[v0 - nb], h->nlo[v1 - nb], h->nup[v0 - nb], h->nup[v1 - nb]);
      launch_node(h, mn, v0, v1, c);
      CK(cudaEventRecord(h->ev_chunk, mn));
    } else {
      // the chunk's "lo" segments on the main stream, its "up" segments on a second one: each stream's launch
      // boundaries (a partial last wave) are filled by the other stream's blocks.  The chunk's node pass follows
      // both on the auxiliary stream, beside the next chunk's sweep.
      if (c == 0) {
        CK(cudaEventRecord(h->ev_phi, mn));
        CK(cudaStreamWaitEvent(h->up, h->ev_phi, 0));
      }
      launch_phi(h, mn, h->up, iter, write_comm, h->nlo[v0 - nb], h->nlo[v1 - nb], h->nup[v0 - nb], h->nup[v1 - nb]);
      CK(cudaEventRecord(h->ev_phi, mn));
      CK(cudaStreamWaitEvent(h->aux, h->ev_phi, 0));
      CK(cudaEventRecord(h->ev_up, h->up));
      CK(cudaStreamWaitEvent(h->aux, h->ev_up, 0));
      launch_node(h, h->aux, v0, v1, c);
      CK(cudaEventRecord(h->ev_chunk, h->aux));
    }
    if (multi) {
      CK(cudaStreamWaitEvent(side, h->ev_chunk, 0));
      if (c == 0) smark(0);
      if ((rc = mg_push_rows(h, side, h->lay.mphi, (size_t)ld * 8, v0, v1))) return rc;
    }
  }
  if (nchunks > 1) CK(cudaStreamWaitEvent(mn, h->ev_chunk, 0));   // all node passes are in (the aux stream is in order)
  if (multi) svi::k_mg_signal<<<1, 32, 0, side>>>(pr, svi::FLAG_M, e);
  smark(1);
  mark(2);
  svi::k_reduce_kpart<<<svi::reduce_kpart_blocks(3, ld), 256, 0, mn>>>(h->d_kpart, nchunks * h->blocks_node, 3, 2 * h->ops.lanes * h->ops.vec, h->d_kvec, ld);
  if (multi) {   // all-reduce of sum, s1, s2 (`sum` feeds the annealing rescale, :541-542)
    svi::k_mg_kx_push<<<1, 256, 0, mn>>>(pr, h->d_kvec, 3 * ld, par, 0, svi::FLAG_KXN, e, nullptr);
    svi::k_mg_kx_sum<<<1, 256, 0, mn>>>(pr, h->d_kvec, 3 * ld, par, 0, svi::FLAG_KXN, e, h->d_mg_err, h->mg_timeout_ns, nullptr);
  }
  if (multi && write_comm && P.nseg) {   // membership words of our rows: merge the peers' replicas (their sweeps are done)
    const size_t first = (size_t)P.node_begin * P.words, cnt = (size_t)(P.node_end - P.node_begin) * P.words;
    svi::k_mg_or_rows<<<(uint32_t)std::min<size_t>((cnt + 255) / 256, (size_t)h->sms * 2), 256, 0, mn>>>(pr, h->lay.mbits, first, cnt);
  }
  mark(3);
  // refresh BEFORE the s3 sweep (it needs `sum` only); its rows travel beside the sweep
  svi::k_scale<<<1, 256, 0, mn>>>(P, annealing);
  h->ops.refresh(P, mn, true);
  h->conv_pending = true;
  mark(4);
  if (multi) {
    CK(cudaEventRecord(h->ev_refresh, mn));
    CK(cudaStreamWaitEvent(side, h->ev_refresh, 0));
    const uint32_t v0 = P.node_begin, v1 = P.node_end;
    smark(2);
    if ((rc = mg_push_rows(h, side, h->lay.b, (size_t)ld * 8, v0, v1))) return rc;
    if ((rc = mg_push_rows(h, side, h->lay.conv[h->cur ^ 1], 4, v0, v1))) return rc;
    if (iter >= 1000) {   // the active-set branch (iter > 1000, :634) reads the neighbours' counts and masks
      if ((rc = mg_push_rows(h, side, h->lay.active, 4, v0, v1))) return rc;
      if ((rc = mg_push_rows(h, side, h->lay.abits, (size_t)P.words * 4, v0, v1))) return rc;
    }
    if (h->share_gamma && (rc = mg_push_rows(h, side, h->lay.gamma, (size_t)ld * 8, v0, v1))) return rc;
    svi::k_mg_signal<<<1, 32, 0, side>>>(pr, svi::FLAG_B, e);
    smark(3);
    svi::k_mg_wait<<<1, 32, 0, mn>>>(pr, svi::FLAG_M, e, h->d_mg_err, h->mg_timeout_ns);
  }
  mark(5);
  launch_s3(h, mn);
  mark(6);
  if (multi) {
    // the shards' "a node newly converged" flags (set by this iteration's refresh) ride along: the next iteration
    // re-partitions the neighbour lists only if some shard saw one
    svi::k_mg_kx_push<<<1, 256, 0, mn>>>(pr, h->d_kvec + 3 * (size_t)ld, ld, par, 1, svi::FLAG_KXS, e, h->d_conv_dirty);
    svi::k_mg_kx_sum<<<1, 256, 0, mn>>>(pr, h->d_kvec + 3 * (size_t)ld, ld, par, 1, svi::FLAG_KXS, e, h->d_mg_err,
                                       h->mg_timeout_ns, h->d_conv_dirty);
  }
  h->ops.lambda(P, mn, annealing, 1);
  flip_converged(h);
  mark(7);
  if (multi) {   // our own pushes belong to the iteration (the peers' are awaited when they are needed)
    CK(cudaEventRecord(h->ev_side, side));
    CK(cudaStreamWaitEvent(mn, h->ev_side, 0));
  }
  mark(8);
  CK(cudaGetLastError());
  return SVI_OK;
}

int svi_ls_mg_timing(svi_ls *h, int enable, double *phase_ms, uint32_t *steps) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  if (phase_ms) {   // mean over the recorded steps (at most the last kTimedSteps), then reset
    CK(cudaStreamSynchronize(h->stream));
    const uint32_t cnt = std::min<uint32_t>(h->tsteps, svi_ls::kTimedSteps);
    if (h->side) CK(cudaStreamSynchronize(h->side));
    for (int i = 0; i < svi_ls::kMarks + 1; ++i) phase_ms[i] = 0.0;
    for (uint32_t s = 0; s < cnt; ++s) {
      for (int i = 0; i < svi_ls::kMarks - 1; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->tev[s][i], h->tev[s][i + 1]));
        phase_ms[i] += ms / cnt;
      }
      if (h->peers.world > 1)
        for (int i = 0; i < 2; ++i) {   // side stream: span of the mphi pushes, span of the b / converged / gamma pushes
          float ms = 0.f;
          CK(cudaEventElapsedTime(&ms, h->sev[s][2 * i], h->sev[s][2 * i + 1]));
          phase_ms[svi_ls::kMarks - 1 + i] += ms / cnt;
        }
    }
    if (steps) *steps = cnt;
    h->tsteps = 0;
  }
  if (enable && !h->tev[0][0]) {
    for (auto &row : h->tev)
      for (cudaEvent_t &ev : row) CK(cudaEventCreate(&ev));
    for (auto &row : h->sev)
      for (cudaEvent_t &ev : row) CK(cudaEventCreate(&ev));
  }
  h->timing = enable != 0;
  h->tsteps = 0;
  return SVI_OK;
}

// rows of the peers that the last svi_ls_mg_step's refresh pushed (gamma when shared, b, converged) are complete
// on return of the stream work queued here
static void mg_await_rows(svi_ls *h) {
  if (h->mg && h->peers.world > 1 && h->epoch > 0)
    svi::k_mg_wait<<<1, 32, 0, h->stream>>>(h->peers, svi::FLAG_B, h->epoch, h->d_mg_err, h->mg_timeout_ns);
}

int svi_ls_mg_error(svi_ls *h) {
  if (!h) return fail(SVI_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  uint32_t v = 0;
  CK(cudaMemcpy(&v, h->d_mg_err, sizeof v, cudaMemcpyDeviceToHost));
  if (v) return fail(SVI_ERR_CUDA, "multi-GPU exchange timed out: flag kind %u from shard %u never arrived", (v >> 4) & 15u, v & 15u);
  return SVI_OK;
}

int svi_ls_get_membership_rows(svi_ls *h, uint32_t first, uint32_t count, uint32_t *bits) {
  if (!h || (!bits && count)) return fail(SVI_ERR_INVALID, "svi_ls_get_membership_rows: null argument");
  if ((uint64_t)first + count > h->P.n) return fail(SVI_ERR_INVALID, "svi_ls_get_membership_rows: rows out of range");
  DeviceGuard guard(h->device);
  if (count)
    CK(cudaMemcpyAsync(bits, h->d_mbits + (size_t)first * h->P.words, (size_t)count * h->P.words * sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return SVI_OK;
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
  // staging: [p | q] as uint32, y as bytes, out as double -- carve from one scratch allocation
  const size_t bytes = npairs * (2 * sizeof(uint32_t) + sizeof(double)) + ((npairs + 7) & ~(size_t)7) + 8;
  int rc = ensure_stage(h, (bytes + sizeof(double) - 1) / sizeof(double));
  if (rc) return rc;
  double *d_out = h->d_stage;
  uint32_t *d_p = reinterpret_cast<uint32_t *>(d_out + npairs);
  uint32_t *d_q = d_p + npairs;
  uint8_t *d_y = reinterpret_cast<uint8_t *>(d_q + npairs);
  unsigned long long *d_bad = reinterpret_cast<unsigned long long *>(d_y + ((npairs + 7) & ~(size_t)7));
  CK(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), h->stream));
  // sharded: the rows of other shards are read from their arenas (peer loads), unless gamma is replicated
  const bool peer_rows = h->mg && h->peers.world > 1 && !h->share_gamma;
  if (h->share_gamma) mg_await_rows(h);
  CK(cudaMemcpyAsync(d_p, p, npairs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_q, q, npairs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(d_y, y, npairs, cudaMemcpyHostToDevice, h->stream));
  h->ops.heldout(h->P, h->stream, npairs, d_p, d_q, d_y, epsilon, d_out, d_bad, peer_rows ? &h->peers : nullptr);
  CK(cudaGetLastError());
  unsigned long long bad = 0;
  CK(cudaMemcpyAsync(loglik, d_out, npairs * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (bad) return fail(SVI_ERR_INVALID, "svi_ls_heldout: pair %llu out of range", bad - 1ull);
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
    // the flags the NEXT sweep reads: after this iteration's refresh that is the freshly pruned copy
    case SVI_BUF_CONVERGED: *dev_ptr = h->d_conv2[h->conv_pending ? h->cur ^ 1 : h->cur]; l = 1; break;
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
