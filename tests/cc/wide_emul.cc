// wide_emul.cc -- the K > 1024 kernels (svinet_b200/csrc/svi_ls_wide.cuh) executed on host threads.
//
// TEST INFRASTRUCTURE ONLY (tests/test_wide_emulated.py).  The kernel SOURCE is the product's, compiled by g++ over
// tests/cc/cuda_shim/: a CUDA block becomes a team of kWideT host threads, __syncthreads a barrier.  Around the
// kernels this file restates, for ONE unsharded handle, what svi_ls.cu does on the host: the half-edge CSR
// ([neighbours not owned | neighbours owned], each part by neighbour id; svi_ls_build.cuh), the segment table
// (push_segments), and the launch order of one iteration (enqueue_step).  The warp-shuffle kernels it needs from
// svi_ls_kernels.cuh (k_reduce_kpart) are restated as plain loops; k_lambda<true> has no shuffle on its log-domain path
// and runs as it is.  The test compares the result with the oracle exactly like the GPU parity tests do, and runs the
// whole thing under ThreadSanitizer to check the barrier placement.
#include "svi_ls_wide.cuh"

#include <pthread.h>

#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

thread_local emu_idx threadIdx, blockIdx;
emu_idx blockDim, gridDim;
static pthread_barrier_t g_sync, g_block_end;
void __syncthreads() { pthread_barrier_wait(&g_sync); }

namespace {

// <<<grid, block>>>: `block` host threads walk the blocks one after another.  The end-of-block barrier is a different
// object from the __syncthreads one, so a kernel whose threads leave a block at different barrier counts hangs here
// instead of silently pairing up unrelated barriers.
void launch(uint32_t grid, uint32_t block, const std::function<void()> &kernel) {
  if (!grid) return;
  blockDim = emu_idx{block, 1, 1};
  gridDim = emu_idx{grid, 1, 1};
  pthread_barrier_init(&g_sync, nullptr, block);
  pthread_barrier_init(&g_block_end, nullptr, block);
  std::vector<std::thread> team;
  for (uint32_t t = 0; t < block; ++t)
    team.emplace_back([&, t] {
      for (uint32_t b = 0; b < grid; ++b) {
        threadIdx = emu_idx{t, 0, 0};
        blockIdx = emu_idx{b, 0, 0};
        kernel();
        pthread_barrier_wait(&g_block_end);
      }
    });
  for (auto &th : team) th.join();
  pthread_barrier_destroy(&g_sync);
  pthread_barrier_destroy(&g_block_end);
}

struct Emu {
  svi::Params P{};
  uint32_t nseg3 = 0, blocks_node = 0, blocks_s3 = 0, cap = 0;
  int cur = 0;
  std::vector<uint32_t> col, seg_node, seg_beg, seg_cnt, seg_nnc, node_seg_lo, node_seg_up, conv[2], active, abits, mbits;
  std::vector<uint32_t> conv_dirty;
  std::vector<double> tl, b, mphi, gamma, gacc, part, kvec, kpart, lambda, eb, scale;
};

void push_segments(uint32_t node, uint32_t beg, uint32_t deg, uint32_t seg_len, Emu *h) {   // svi_ls.cu::push_segments
  if (!deg) return;
  const uint32_t nch = (deg + seg_len - 1) / seg_len, base = deg / nch, extra = deg % nch;
  uint32_t at = beg;
  for (uint32_t c = 0; c < nch; ++c) {
    const uint32_t len = base + (c < extra ? 1u : 0u);
    h->seg_node.push_back(node);
    h->seg_beg.push_back(at);
    h->seg_cnt.push_back(len);
    at += len;
  }
}

uint32_t s3_owner(uint32_t p, uint32_t q) {   // svi_ls_build.cuh
  const uint32_t lo = p < q ? p : q, hi = p < q ? q : p;
  return ((lo ^ hi) & 1u) ? lo : hi;
}

void reduce_kpart(const Emu *h, uint32_t nblocks, uint32_t nvec, double *kvec) {   // k_reduce_kpart, as a plain loop
  const uint32_t ld = h->P.ld;
  for (uint32_t v = 0; v < nvec; ++v)
    for (uint32_t c = 0; c < ld; ++c) {
      double s = 0.0;
      if (c < h->cap)
        for (uint32_t bl = 0; bl < nblocks; ++bl) s += h->kpart[((size_t)bl * nvec + v) * h->cap + c];
      kvec[(size_t)v * ld + c] = s;
    }
}

void flip(Emu *h) {
  h->cur ^= 1;
  h->P.conv = h->conv[h->cur].data();
  h->P.conv_next = h->conv[h->cur ^ 1].data();
}

}  // namespace

// (the harness is built once per emulated block size and several builds may live in one process: everything but this
// API is compiled with hidden visibility and -fno-gnu-unique so that the builds do not share statics)
#pragma GCC visibility push(default)
extern "C" {

uint32_t we_threads(void) { return svi::kWideT; }

// one unsharded handle over n nodes; links = nlinks (p, q) pairs; tl = 2 x training degree; blocks_* <= 0: the
// product's choice for `sms` multiprocessors
Emu *we_create(uint32_t n, uint32_t k, uint64_t nlinks, const uint32_t *links, const double *tl, double alpha, double eta0,
               double eta1, uint32_t ones, uint32_t seg_len, uint32_t sms) {
  Emu *h = new Emu();
  const uint32_t ld = (k + 3u) & ~3u, words = (k + 31u) / 32u;
  std::vector<std::vector<uint32_t>> lo(n), up(n);
  for (uint64_t e = 0; e < nlinks; ++e) {
    const uint32_t p = links[2 * e], q = links[2 * e + 1];
    if (p >= n || q >= n || p == q) { delete h; return nullptr; }
    const uint32_t own = s3_owner(p, q), oth = own == p ? q : p;
    up[own].push_back(oth);
    lo[oth].push_back(own);
  }
  uint64_t he = 0;
  for (uint32_t v = 0; v < n; ++v) {
    std::sort(lo[v].begin(), lo[v].end());
    std::sort(up[v].begin(), up[v].end());
    he += lo[v].size() + up[v].size();
  }
  h->node_seg_lo.assign(n + 1, 0);
  h->node_seg_up.assign(n + 1, 0);
  {
    uint64_t at = 0;
    for (uint32_t v = 0; v < n; ++v) {
      h->col.insert(h->col.end(), lo[v].begin(), lo[v].end());
      h->col.insert(h->col.end(), up[v].begin(), up[v].end());
      push_segments(v, (uint32_t)at, (uint32_t)lo[v].size(), seg_len, h);
      h->node_seg_lo[v + 1] = (uint32_t)h->seg_node.size();
      at += lo[v].size() + up[v].size();
    }
    const uint32_t nseg_lo = (uint32_t)h->seg_node.size();
    at = 0;
    h->node_seg_up[0] = nseg_lo;
    for (uint32_t v = 0; v < n; ++v) {
      push_segments(v, (uint32_t)(at + lo[v].size()), (uint32_t)up[v].size(), seg_len, h);
      h->node_seg_up[v + 1] = (uint32_t)h->seg_node.size();
      at += lo[v].size() + up[v].size();
    }
    h->P.nseg_lo = nseg_lo;
  }
  const uint32_t nseg = (uint32_t)h->seg_node.size();
  h->nseg3 = nseg - h->P.nseg_lo;
  h->seg_nnc = h->seg_cnt;
  h->cap = svi::wide_cap(ld);
  // grids as svi_ls_create sizes them for the wide tile (lanes = kWideT, one work item per block)
  h->blocks_node = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(2 * (int64_t)sms, ((int64_t)n + 7) / 8));
  h->blocks_s3 = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(8 * (int64_t)sms, ((int64_t)h->nseg3 + 3) / 4));
  const size_t nld = (size_t)n * ld;
  h->tl.assign(tl, tl + n);
  h->b.assign(nld, 0.0); h->mphi.assign(nld, 0.0); h->gamma.assign(nld, 0.0); h->gacc.assign(nld, 0.0);
  h->part.assign((size_t)std::max<uint32_t>(nseg, 1) * ld, 0.0);
  h->kvec.assign(4 * (size_t)ld, 0.0);
  h->kpart.assign((size_t)std::max(h->blocks_node, h->blocks_s3) * 3 * h->cap, 0.0);
  h->lambda.assign(2 * (size_t)k, 0.0); h->eb.assign(ld, 0.0); h->scale.assign(ld, 0.0);
  h->conv[0].assign(n, 0); h->conv[1].assign(n, 0); h->active.assign(n, 0);
  h->abits.assign((size_t)n * words, 0); h->mbits.assign((size_t)n * words, 0);
  h->conv_dirty.assign(1, 0);
  if (h->col.empty()) h->col.push_back(0);
  for (auto *v : {&h->seg_node, &h->seg_beg, &h->seg_cnt, &h->seg_nnc}) if (v->empty()) v->push_back(0);
  svi::Params &P = h->P;
  P.n = n; P.k = k; P.ld = ld; P.words = words;
  P.node_begin = 0; P.node_end = n; P.shard_begin = 0;
  P.alpha = alpha; P.eta0 = eta0; P.eta1 = eta1; P.ones_d = (double)ones;
  P.k_div10 = k / 10;
  P.col = h->col.data();
  P.seg_node = h->seg_node.data(); P.seg_beg = h->seg_beg.data(); P.seg_cnt = h->seg_cnt.data(); P.seg_nnc = h->seg_nnc.data();
  P.nseg = nseg;
  P.node_seg_lo = h->node_seg_lo.data(); P.node_seg_up = h->node_seg_up.data();
  P.conv_dirty = h->conv_dirty.data();
  P.tl = h->tl.data();
  P.b = h->b.data(); P.mphi = h->mphi.data(); P.gamma = h->gamma.data(); P.gacc = h->gacc.data(); P.part = h->part.data();
  P.kvec = h->kvec.data(); P.kpart = h->kpart.data(); P.lambda = h->lambda.data(); P.eb = h->eb.data(); P.scale = h->scale.data();
  P.conv = h->conv[0].data(); P.conv_next = h->conv[1].data();
  P.active = h->active.data(); P.abits = h->abits.data(); P.mbits = h->mbits.data();
  return h;
}

void we_destroy(Emu *h) { delete h; }

void we_info(const Emu *h, uint32_t *nseg, uint32_t *nseg_lo, uint32_t *blocks_node, uint32_t *blocks_s3) {
  *nseg = h->P.nseg; *nseg_lo = h->P.nseg_lo; *blocks_node = h->blocks_node; *blocks_s3 = h->blocks_s3;
}

void we_set_state(Emu *h, const double *gamma, const double *lambda) {   // svi_ls_set_state
  const svi::Params &P = h->P;
  for (uint32_t i = 0; i < P.n; ++i)
    for (uint32_t c = 0; c < P.ld; ++c) h->gamma[(size_t)i * P.ld + c] = c < P.k ? gamma[(size_t)i * P.k + c] : 0.0;
  std::copy(lambda, lambda + 2 * (size_t)P.k, h->lambda.begin());
  launch(P.n, svi::kWideT, [&] { svi::k_refresh_wide<false>(P); });
  launch(1, 256, [&] { svi::k_lambda<true>(P, 0, 0); });
}

void we_get_state(const Emu *h, double *gamma, double *lambda) {
  const svi::Params &P = h->P;
  for (uint32_t i = 0; i < P.n; ++i)
    for (uint32_t c = 0; c < P.k; ++c) gamma[(size_t)i * P.k + c] = h->gamma[(size_t)i * P.ld + c];
  std::copy(h->lambda.begin(), h->lambda.end(), lambda);
}

void we_set_converged(Emu *h, const uint32_t *conv) { std::copy(conv, conv + h->P.n, h->conv[h->cur].begin()); }

void we_get_converged(const Emu *h, uint32_t *conv, uint32_t *active) {
  std::copy(h->conv[h->cur].begin(), h->conv[h->cur].end(), conv);
  std::copy(h->active.begin(), h->active.end(), active);
}

void we_get_kvectors(const Emu *h, double *sum, double *s1, double *s2, double *s3) {
  const uint32_t k = h->P.k, ld = h->P.ld;
  std::copy(h->kvec.begin(), h->kvec.begin() + k, sum);
  std::copy(h->kvec.begin() + ld, h->kvec.begin() + ld + k, s1);
  std::copy(h->kvec.begin() + 2 * (size_t)ld, h->kvec.begin() + 2 * (size_t)ld + k, s2);
  std::copy(h->kvec.begin() + 3 * (size_t)ld, h->kvec.begin() + 3 * (size_t)ld + k, s3);
}

void we_get_membership(const Emu *h, uint32_t *bits) { std::copy(h->mbits.begin(), h->mbits.end(), bits); }

// svi_ls_step of an unsharded handle: begin_iteration, launch_phi (one arg-max per link: the "lo" segments without the
// tally, the "up" segments with it and publish = 1), node pass, s3 sweep, lambda, refresh, flip
void we_step(Emu *h, uint32_t iter, int annealing, int write_comm) {
  const svi::Params &P = h->P;
  const bool sparse = iter > 1000 && P.k_div10 > 0;
  const uint32_t T = svi::kWideT, lo = P.nseg_lo, ns = P.nseg;
  if (write_comm) std::fill(h->mbits.begin(), h->mbits.end(), 0u);
  if (!write_comm) {
    if (sparse) launch(ns, T, [&] { svi::k_phi_wide<true, false>(P, 0, lo, lo, ns, 0); });
    else launch(ns, T, [&] { svi::k_phi_wide<false, false>(P, 0, lo, lo, ns, 0); });
  } else {
    if (sparse) {
      launch(lo, T, [&] { svi::k_phi_wide<true, false>(P, 0, lo, 0, 0, 0); });
      launch(ns - lo, T, [&] { svi::k_phi_wide<true, true>(P, lo, ns, 0, 0, 1); });
    } else {
      launch(lo, T, [&] { svi::k_phi_wide<false, false>(P, 0, lo, 0, 0, 0); });
      launch(ns - lo, T, [&] { svi::k_phi_wide<false, true>(P, lo, ns, 0, 0, 1); });
    }
  }
  launch(h->blocks_node, T, [&] { svi::k_node_wide(P, h->cap); });
  reduce_kpart(h, h->blocks_node, 3, h->kvec.data());
  launch(h->blocks_s3, T, [&] { svi::k_s3_wide(P, h->cap); });
  reduce_kpart(h, h->blocks_s3, 1, h->kvec.data() + 3 * (size_t)P.ld);
  launch(1, 256, [&] { svi::k_lambda<true>(P, annealing, 1); });
  launch(P.n, T, [&] { svi::k_refresh_wide<true>(P); });
  flip(h);
}

// svi_ls_heldout: out[i] = log-likelihood of pair i; returns 1 + index of a bad pair, or 0
uint64_t we_heldout(Emu *h, uint64_t npairs, const uint32_t *p, const uint32_t *q, const uint8_t *y, double epsilon,
                    double *out, uint32_t blocks) {
  unsigned long long bad = 0;
  const svi::Params &P = h->P;
  const svi::LocalRows rows{P.gamma};
  launch(blocks, svi::kWideT, [&] { svi::k_heldout_wide<svi::LocalRows>(P, rows, npairs, p, q, y, epsilon, out, &bad); });
  return bad;
}

}  // extern "C"
#pragma GCC visibility pop
