// fa2_wide_emul.cc -- the K > 512 kernels of `-rnode -stratified` (svinet_b200/csrc/svi_fa2_wide.cuh) on host threads.
//
// TEST INFRASTRUCTURE ONLY (tests/test_fa2_wide_emulated.py).  Same scheme as wide_emul.cc: the kernel SOURCE is the
// product's, compiled by g++ over tests/cc/cuda_shim/ (a block = kWideT host threads, __syncthreads = a barrier); this
// file restates what svi_fa2.cu does on the host for them -- the stored-row layout (lazy: u with gamma = alpha + c*u),
// the control block of svi_fa2_step, the launch order of one iteration (prep, pairs, [blend], lambda, the re-basing
// fold) -- and k_fa2_lambda, whose warp shuffles do not run here, as a plain loop.
#include "svi_fa2_wide.cuh"

#include <pthread.h>

#include <cmath>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

thread_local emu_idx threadIdx, blockIdx;
emu_idx blockDim, gridDim;
static pthread_barrier_t g_sync, g_block_end;
void __syncthreads() { pthread_barrier_wait(&g_sync); }

namespace {

using svi::Fa2Ctrl;
using svi::Fa2Params;

void launch(uint32_t grid, uint32_t block, const std::function<void()> &kernel) {   // see wide_emul.cc
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
  Fa2Params P{};
  Fa2Ctrl ctrl{};
  uint32_t cap = 0;
  uint64_t nodec = 0;
  double log_c = 0.0, fold_below = -230.0;   // the re-basing fold of the lazy rows (svi_fa2.cu::launch_iteration)
  uint32_t folds = 0;
  std::vector<double> gamma, lambda, elogbeta, elogf, epi, partS, partL, wide;
  std::vector<uint32_t> pairs;
  std::vector<uint8_t> touched;
};

void lambda_pass(Emu *h) {   // k_fa2_lambda (groups_per_block = 1: block b of the pair kernel had pairs iff b < npairs)
  Fa2Params &P = h->P;
  const Fa2Ctrl c = h->ctrl;
  const uint32_t active = std::min(P.pair_blocks, c.npairs);
  const double coef = c.rho_node * c.scale / ((1.0 - c.rho_node) * c.cscale);
  double *urow = P.gamma + (size_t)c.start * P.ld;
  for (uint32_t z = 0; z < P.k; ++z) {
    double s = 0.0, l = 0.0;
    for (uint32_t b = 0; b < active; ++b) {
      s += P.partS[(size_t)b * h->cap + z];
      l += P.partL[(size_t)b * h->cap + z];
    }
    if (P.lazy) urow[z] = std::fma(coef, s, urow[z]);
    if (!P.nolambda)
      for (uint32_t t = 0; t < 2; ++t) {
        const double raw = t == c.type ? l : 0.0;
        const double ldt = (t == 0 ? P.eta0 : P.eta1) + c.scale * raw;
        P.lambda[2 * z + t] = (1.0 - c.rho_t) * P.lambda[2 * z + t] + c.rho_t * ldt;
      }
  }
  h->ctrl.nodec = c.nodec + 1.0;
  if (P.lazy) h->ctrl.cscale = (1.0 - c.rho_node) * c.cscale;
  h->ctrl.total_sampled = c.total_sampled + c.sampled_inc;
  h->ctrl.total_rounds = c.total_rounds + h->ctrl.last_rounds;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

uint32_t fwe_threads(void) { return svi::kWideT; }

// the reference's defaults (svi_fa2_default_config) except where given
Emu *fwe_create(uint32_t n, uint32_t k, int eager, uint32_t pair_blocks, uint32_t online_iterations, double fold_below) {
  Emu *h = new Emu();
  Fa2Params &P = h->P;
  P.n = n; P.k = k; P.ld = (k + 3u) & ~3u;
  P.alpha = 1.0 / k; P.eta0 = 1.0; P.eta1 = 1.0;
  P.epsilon = 1e-30; P.logeps = std::log(1e-30); P.thresh = 1e-5;
  P.tau0 = 1025.0; P.kappa = 0.9; P.nodetau0 = 1025.0; P.nodekappa = 0.5;
  P.inf_epsilon = 0.5; P.online_iters = online_iterations; P.m_sets = 10; P.nolambda = 0;
  P.lazy = eager ? 0u : 1u;
  P.pair_blocks = pair_blocks;
  h->cap = svi::wide_cap(P.ld);
  h->fold_below = fold_below;   // (-230 in the product; a test raises it to see a fold within a few iterations)
  h->gamma.assign((size_t)n * P.ld, 0.0);
  h->lambda.assign(2 * (size_t)k, 0.0);
  h->elogbeta.assign(2 * (size_t)P.ld, 0.0);
  h->elogf.assign(P.ld, 0.0);
  h->epi.assign(P.ld, 0.0);
  h->partS.assign((size_t)pair_blocks * h->cap, 0.0);
  h->partL.assign((size_t)pair_blocks * h->cap, 0.0);
  h->wide.assign(((size_t)pair_blocks * svi::kFa2WideRows + svi::kFa2WideOneRows) * P.ld, 0.0);
  h->pairs.assign(2, 0);
  h->touched.assign(n, 0);
  P.gamma = h->gamma.data(); P.lambda = h->lambda.data(); P.elogbeta = h->elogbeta.data(); P.elogf = h->elogf.data();
  P.epi_start = h->epi.data(); P.partS = h->partS.data(); P.partL = h->partL.data(); P.wide = h->wide.data();
  P.pairs = h->pairs.data(); P.cap_pairs = 1; P.touched = h->touched.data(); P.ctrl = &h->ctrl;
  return h;
}

void fwe_destroy(Emu *h) { delete h; }

void fwe_set_state(Emu *h, const double *gamma, const double *lambda, uint64_t nodec) {   // svi_fa2_set_state
  Fa2Params &P = h->P;
  const double a0 = P.lazy ? P.alpha : 0.0;
  for (uint32_t i = 0; i < P.n; ++i)
    for (uint32_t c = 0; c < P.ld; ++c) h->gamma[(size_t)i * P.ld + c] = c < P.k ? gamma[(size_t)i * P.k + c] - a0 : 0.0;
  std::copy(lambda, lambda + 2 * (size_t)P.k, h->lambda.begin());
  memset(&h->ctrl, 0, sizeof h->ctrl);
  h->ctrl.nodec = (double)nodec;
  h->ctrl.cscale = 1.0;
  h->log_c = 0.0;
  std::fill(h->touched.begin(), h->touched.end(), 0);
  h->nodec = nodec;
}

void fwe_get_state(const Emu *h, double *gamma, double *lambda) {   // k_fa2_export
  const Fa2Params &P = h->P;
  const svi::Fa2Map gm = svi::fa2_map(P, h->ctrl);
  for (uint32_t i = 0; i < P.n; ++i)
    for (uint32_t c = 0; c < P.k; ++c) gamma[(size_t)i * P.k + c] = gm(h->gamma[(size_t)i * P.ld + c]);
  std::copy(h->lambda.begin(), h->lambda.end(), lambda);
}

// svi_fa2_step: control block, then launch_iteration (prep, pairs, [blend], lambda, fold)
int fwe_step(Emu *h, uint32_t iter, uint32_t type, uint32_t start, uint64_t npairs, const uint32_t *pairs) {
  Fa2Params &P = h->P;
  if (type > 1 || start >= P.n) return -1;
  for (uint64_t i = 0; i < npairs; ++i) {
    const uint32_t p = pairs[2 * i], q = pairs[2 * i + 1];
    if (p >= P.n || q >= P.n || p == q || (p != start && q != start)) return -1;
  }
  h->pairs.assign(pairs, pairs + 2 * npairs);
  if (h->pairs.empty()) h->pairs.assign(2, 0);
  P.pairs = h->pairs.data();
  Fa2Ctrl &c = h->ctrl;
  c.type = type; c.start = start; c.npairs = (uint32_t)npairs; c.iter = iter;
  c.sampled_inc = npairs;
  c.rho_node = std::pow(P.nodetau0 + (double)h->nodec, -1 * P.nodekappa);
  c.rho_t = std::pow(P.tau0 + ((double)iter + 1.0), -1 * P.kappa);
  c.scale = type == 0 ? (double)P.n / (2 * (1 - P.inf_epsilon)) : ((double)P.n * (double)P.m_sets) / (2 * P.inf_epsilon);
  const uint32_t T = svi::kWideT;
  launch(1, T, [&] { svi::k_fa2_prep_wide(P); });
  launch(P.pair_blocks, T, [&] { svi::k_fa2_pairs_wide(P, h->cap); });
  if (!P.lazy) launch(std::max(1u, std::min(P.n, 5u)), T, [&] { svi::k_fa2_blend_wide(P, h->cap); });
  lambda_pass(h);
  if (P.lazy) {
    const double rho = std::pow(P.nodetau0 + (double)h->nodec, -1 * P.nodekappa);
    h->log_c += std::log1p(-rho);
    if (h->log_c < h->fold_below) {   // k_fa2_fold + k_fa2_reset_scale
      for (double &u : h->gamma) u *= c.cscale;
      c.cscale = 1.0;
      h->log_c = 0.0;
      h->folds++;
    }
  }
  h->nodec++;
  return 0;
}

uint32_t fwe_folds(const Emu *h) { return h->folds; }
uint64_t fwe_last_rounds(const Emu *h) { return h->ctrl.last_rounds; }

void fwe_heldout(Emu *h, uint64_t npairs, const uint32_t *p, const uint32_t *q, const uint8_t *y, double *out, uint32_t blocks) {
  const Fa2Params &P = h->P;
  launch(blocks, svi::kWideT, [&] { svi::k_fa2_heldout_wide(P, npairs, p, q, y, out); });
}

void fwe_phi_pair(Emu *h, uint32_t p, uint32_t q, int y, double *phi1, double *phi2, uint32_t *rounds) {   // svi_fa2_phi_pair
  Fa2Params &P = h->P;
  std::vector<double> out(2 * (size_t)P.k);
  launch(1, svi::kWideT, [&] { svi::k_fa2_prep_wide(P); });
  double *rows = P.wide + (size_t)P.pair_blocks * svi::kFa2WideRows * P.ld;
  launch(1, svi::kWideT, [&] { svi::k_fa2_one_pair_wide(P, rows, p, q, y ? 1 : 0, out.data(), rounds); });
  std::copy(out.begin(), out.begin() + P.k, phi1);
  std::copy(out.begin() + P.k, out.end(), phi2);
}

}  // extern "C"
#pragma GCC visibility pop
