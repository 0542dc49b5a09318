// linksampling.cc -- see linksampling.hh.  All file:line citations refer to the reference's
// src/linksampling.cc unless another file is named.
#include "linksampling.hh"
#include "fixed_fmt.hh"
#include "textio.hh"
#include "pool.hh"
#include "nmi.hh"
#include "mt_jump.hh"

#include <sys/mman.h>

#include <algorithm>
#include <cassert>
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <map>
#include <mutex>
#include <sstream>
#include <thread>

namespace {

[[noreturn]] void die_dev(const char *what) {
  fprintf(stderr, "svinet: %s failed: %s\n", what, svi_ls_last_error());
  exit(-1);
}
#define DEV(call) do { if ((call) != SVI_OK) die_dev(#call); } while (0)

FILE *open_or_die(const std::string &path, const char *mode, const char *what) {
  FILE *f = fopen(path.c_str(), mode);
  if (!f) {
    printf("cannot open %s file:%s\n", what, strerror(errno));
    exit(-1);
  }
  return f;
}

// Format `rows` lines in parallel (each line built by `line(i, buf)`), write them in order.
template <class F>
void write_rows(FILE *f, uint32_t rows, F line) {
  const uint32_t block = 1u << 15;
  const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  for (uint32_t r0 = 0; r0 < rows; r0 += block) {
    const uint32_t r1 = std::min(rows, r0 + block);
    std::vector<std::string> out(nt);
    std::vector<std::thread> th;
    const uint32_t per = (r1 - r0 + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        const uint32_t a = std::min(r1, r0 + t * per), b = std::min(r1, a + per);
        for (uint32_t i = a; i < b; ++i) line(i, out[t]);
      });
    for (auto &x : th) x.join();
    for (auto &s : out) fwrite(s.data(), 1, s.size(), f);
  }
}

// SVINET_TIMING=1: wall-clock laps on stderr (development aid, not in the reference)
struct Lap {
  bool on = getenv("SVINET_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void operator()(const char *what) {
    const auto now = std::chrono::steady_clock::now();
    if (on) fprintf(stderr, "[linksampling] %-30s %.3f s\n", what, std::chrono::duration<double>(now - t).count());
    t = now;
  }
};


}  // namespace

LinkSampling::LinkSampling(Env &env, Network &network)
    : env_(env), net_(network), n_(env.n), k_(env.k), rng_(0), start_time_(time(0)) {
  total_pairs_ = (double)(uint32_t)(n_ * (n_ - 1) / 2);          // 32-bit product, :37
  env_.plog("inference n", n_);
  env_.plog("total pairs", total_pairs_);
  ones_prob_ = double(net_.ones()) / total_pairs_;               // :49-50
  zeros_prob_ = 1 - ones_prob_;
  env_.plog("ones_prob", ones_prob_);
  env_.plog("zeros_prob", zeros_prob_);
  uint32_t maxdeg;
  double avgdeg;
  net_.deg_stats(maxdeg, avgdeg);
  env_.plog("avg degree", avgdeg);
  env_.plog("max degree", maxdeg);
  {
    // the K x 2 prior printed the way the reference's Matrix::s() prints it (src/matrix.hh:898-924)
    std::ostringstream sa;
    sa << "\n[ ";
    for (uint32_t i = 0; i < std::min<uint32_t>(k_, 512); ++i) {
      const double row[2] = {env_.eta0, env_.eta1};
      for (int j = 0; j < 2; ++j) {
        double u = row[j];
        if (u < 1e-05 && u > .0) u = .0;
        if (i > 0 && j == 0) sa << "  " << u << " ";
        else sa << u << " ";
      }
      sa << "\n";
    }
    sa << "]";
    env_.plog("eta", sa.str());
  }
  if (env_.seed) rng_.set((unsigned long)env_.seed);             // :74-75
  Lap lap;

  FILE *vef = open_or_die(env_.file("/validation-edges.txt"), "w", "validation edges");
  FILE *tef = open_or_die(env_.file("/test-edges.txt"), "w", "test edges");
  fclose(tef);
  if (!env_.load_heldout) {
    env_.plog("load validation from file:", false);
    init_validation();
  } else {
    env_.plog("load validation from file:", true);
    load_validation();
  }
  for (const Edge &e : validation_pairs_)                        // edgelist_s, :190-206
    fprintf(vef, "%d\t%d\t%d\n", net_.seq2id(e.first), net_.seq2id(e.second), (int)net_.y(e.first, e.second));
  fprintf(vef, "\n");
  fclose(vef);
  if (env_.load_test) fprintf(stderr, "svinet: -load-test is accepted but the test set is not used in this build\n");
  lap("held-out draw");

  if (env_.nmi) load_ground_truth();                             // network.cc:120-123

  // gamma: reserved, advised to use huge pages BEFORE the first touch (init_gamma2 and the writers walk its rows at
  // random: with 4 KB pages every row is a TLB miss), then zero-filled
  gamma_.reserve((size_t)n_ * k_);
#ifdef MADV_HUGEPAGE
  {
    const uintptr_t a = ((uintptr_t)gamma_.data() + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
    const uintptr_t e = ((uintptr_t)gamma_.data() + gamma_.capacity() * sizeof(double)) & ~(uintptr_t)((2u << 20) - 1);
    if (e > a) madvise((void *)a, e - a, MADV_HUGEPAGE);
  }
#endif
  gamma_.assign((size_t)n_ * k_, 0.0);
  lambda_.assign((size_t)k_ * 2, 0.0);
  if (env_.model_load) {
    if (load_model() < 0) exit(-1);
  } else if (env_.use_init_communities) {
    init_gamma_external();
    for (uint32_t c = 0; c < k_; ++c) {                          // init_lambda, :364-372 (:115-116)
      lambda_[2 * c] = env_.eta0;
      lambda_[2 * c + 1] = env_.eta1;
    }
  } else {
    init_gamma2();
    for (uint32_t c = 0; c < k_; ++c) {                          // init_lambda, :364-372
      lambda_[2 * c] = env_.eta0;
      lambda_[2 * c + 1] = env_.eta1;
    }
  }

  lap("init gamma/lambda");
  tf_ = open_or_die(env_.file("/test.txt"), "w", "test");
  vf_ = open_or_die(env_.file("/validation.txt"), "w", "validation");
  env_.plog("network ones", net_.ones());
  env_.plog("network singles", net_.singles());
  lf_ = open_or_die(env_.file("/logl.txt"), "w", "logl");

  assign_training_links();                                       // :566 (no RNG use, so it can run here)
  lap("assign_training_links");
  if (env_.dump_only) return;

  create_device();
  lap("svi_ls_create + set_state");

  // held-out pairs in std::map<Edge,bool> order (lexicographic), the order validation_likelihood sums in
  // (the reference holds them in a map: a -load-validation file that repeats a pair still counts it once)
  validation_sorted_ = validation_pairs_;
  std::sort(validation_sorted_.begin(), validation_sorted_.end());
  validation_sorted_.erase(std::unique(validation_sorted_.begin(), validation_sorted_.end()), validation_sorted_.end());
  for (const Edge &e : validation_sorted_) {
    hp_.push_back(e.first);
    hq_.push_back(e.second);
    hy_.push_back(net_.y(e.first, e.second) ? 1 : 0);
  }
  hll_.resize(hp_.size());
  validation_likelihood();                                       // :150
  start_time_ = time(0);
}

LinkSampling::~LinkSampling() {
  if (vf_) fclose(vf_);
  if (tf_) fclose(tf_);
  if (lf_) fclose(lf_);
  for (svi_ls *d : devs_) svi_ls_sync(d);      // no shard may vanish while a peer still pushes rows into it
  for (svi_ls *d : devs_) svi_ls_destroy(d);
}

// The device side of the object: one handle, or with -gpus N one handle per GPU, each owning an edge-balanced
// contiguous node block; the shards exchange rows over peer memory inside svi_ls_mg_step.
// The seam stays src/main.cc:337-341: one LinkSampling object driven by one host thread.
void LinkSampling::create_device() {
  const uint32_t g = (uint32_t)std::max(1, env_.ngpus);
  svi_ls_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.n = n_; cfg.k = k_; cfg.nlinks = links_.size() / 2;
  cfg.alpha = env_.alpha; cfg.eta0 = env_.eta0; cfg.eta1 = env_.eta1;
  cfg.ones = net_.ones(); cfg.device = -1; cfg.seg_len = 0;
  cfg.node_begin = 0; cfg.node_end = n_;
  devs_.assign(g, nullptr);
  if (g == 1) {
    DEV(svi_ls_create(&cfg, links_.data(), training_links_.data(), &devs_[0]));
    dev_ = devs_[0];
    DEV(svi_ls_set_state(dev_, gamma_.data(), lambda_.data()));
    return;
  }
  // node blocks holding ~1/g of the half-edges each (training_links_[v] = 2 x degree)
  bounds_.assign(g + 1, n_);
  bounds_[0] = 0;
  double total = 0, run = 0;
  for (uint32_t v = 0; v < n_; ++v) total += training_links_[v];
  uint32_t r = 1;
  for (uint32_t v = 0; v < n_ && r < g; ++v) {
    while (r < g && run >= total * r / g) bounds_[r++] = v;
    run += training_links_[v];
  }
  // the shards are built concurrently (each uploads the link list to its GPU and sorts its half-edges there)
  // (SVINET_SHARDS_ON_ONE_GPU=1: all shards on the current device -- the same code path on a one-GPU box)
  const char *same = getenv("SVINET_SHARDS_ON_ONE_GPU");
  const bool one_gpu = same && same[0] == '1';
  std::vector<int> rc(g, 0);
  std::vector<std::string> msg(g);
  std::vector<std::thread> th;
  for (uint32_t i = 0; i < g; ++i)
    th.emplace_back([&, i] {
      svi_ls_config c = cfg;
      c.device = one_gpu ? -1 : (int32_t)i;
      c.node_begin = bounds_[i];
      c.node_end = bounds_[i + 1];
      rc[i] = svi_ls_create(&c, links_.data(), training_links_.data(), &devs_[i]);
      if (rc[i]) msg[i] = svi_ls_last_error();
    });
  for (auto &t : th) t.join();
  for (uint32_t i = 0; i < g; ++i)
    if (rc[i]) {
      fprintf(stderr, "svinet: svi_ls_create on GPU %u failed: %s\n", i, msg[i].c_str());
      exit(-1);
    }
  dev_ = devs_[0];
  for (uint32_t i = 0; i < g; ++i) {
    DEV(svi_ls_peer_attach_local(devs_[i], g, i, bounds_.data(), devs_.data(), 0));
  }
  // every shard takes the whole start state (it derives the expectations of all rows itself): uploaded concurrently
  std::fill(rc.begin(), rc.end(), 0);
  th.clear();
  for (uint32_t i = 0; i < g; ++i)
    th.emplace_back([&, i] {
      rc[i] = svi_ls_set_state(devs_[i], gamma_.data(), lambda_.data());
      if (rc[i]) msg[i] = svi_ls_last_error();
    });
  for (auto &t : th) t.join();
  for (uint32_t i = 0; i < g; ++i)
    if (rc[i]) {
      fprintf(stderr, "svinet: svi_ls_set_state on GPU %u failed: %s\n", i, msg[i].c_str());
      exit(-1);
    }
}

void LinkSampling::device_step(bool write_comm) {
  if (devs_.size() == 1) {
    DEV(svi_ls_step(dev_, iter_, annealing_ ? 1 : 0, write_comm ? 1 : 0));
    return;
  }
  for (svi_ls *d : devs_) DEV(svi_ls_mg_step(d, iter_, annealing_ ? 1 : 0, write_comm ? 1 : 0));
}

void LinkSampling::device_sync() {
  for (svi_ls *d : devs_) DEV(svi_ls_sync(d));
}

bool LinkSampling::edge_ok(const Edge &e) const {
  if (e.first == e.second) return false;
  return held_keys_.find(((uint64_t)e.first << 32) | e.second) == held_keys_.end();
}

void LinkSampling::get_random_edge(bool link, Edge &e) {
  if (!link) {
    do {
      e.first = (uint32_t)rng_.uniform_int(n_);
      e.second = (uint32_t)rng_.uniform_int(n_);
      Network::order_edge(e);
    } while (!edge_ok(e));
  } else {
    do {
      e = net_.edges()[rng_.uniform_int(net_.ones())];
    } while (!edge_ok(e));
  }
}

void LinkSampling::set_validation_sample(int s) {
  int c0 = 0, c1 = 0;
  const int p = s / 2;
  while (c0 < p || c1 < p) {
    Edge e;
    get_random_edge(c0 == p, e);   // non-link candidates first, then links
    const bool y = net_.y(e.first, e.second);
    bool keep = false;
    if (!y && c0 < p) { c0++; keep = true; }
    if (y && c1 < p) { c1++; keep = true; }
    if (keep) {
      validation_pairs_.push_back(e);
      held_keys_.insert(((uint64_t)e.first << 32) | e.second);
    }
  }
}

void LinkSampling::init_validation() {
  const int s1 = (int)(env_.heldout_ratio * net_.ones());        // :167
  set_validation_sample(s1);
  env_.plog("heldout ratio", env_.heldout_ratio);
  env_.plog("validation pairs (1s and 0s)", (uint64_t)validation_pairs_.size());
}

void LinkSampling::load_validation() {
  FILE *f = fopen(env_.load_heldout_fname.c_str(), "r");
  if (!f) {
    fprintf(stderr, "error: cannot read test validation file %s\n", env_.load_heldout_fname.c_str());
    exit(-1);
  }
  uint32_t a, b, cnt = 0;
  while (fscanf(f, "%u %u", &a, &b) == 2) {
    uint32_t p, q;
    if (!net_.id2seq(a, &p) || !net_.id2seq(b, &q)) {
      fprintf(stderr, "error: id %d or id %d not found in original network\n", a, b);
      exit(-1);
    }
    Edge e(p, q);
    Network::order_edge(e);
    validation_pairs_.push_back(e);
    ++cnt;
    int c;   // the reference's pattern is "%d\t%d\n": a third column (y) on the line is not expected
    while ((c = fgetc(f)) != EOF && c != '\n') {}
  }
  fclose(f);
  validation_sorted_ = validation_pairs_;
  std::sort(validation_sorted_.begin(), validation_sorted_.end());
  validation_sorted_.erase(std::unique(validation_sorted_.begin(), validation_sorted_.end()), validation_sorted_.end());
  for (const Edge &e : validation_sorted_) held_keys_.insert(((uint64_t)e.first << 32) | e.second);
  env_.plog("link sampling: loaded validation heldout pairs:", cnt);
}

// -init-communities <file>: one community per line (external node ids), Network::load_init_communities
// (src/network.cc:374-437; it also writes init_memberships.txt), then init_gamma_external (:404-452): every node p
// adds, ONCE PER ADJACENCY ENTRY, the normalised vector phi_p[k] = alpha + [k in communities(p)] * n / |communities(p)|
// to its row (the reference's loop body does not depend on the neighbour; the repeated addition is kept, it is what
// defines the rounding).  No RNG use.
void LinkSampling::init_gamma_external() {
  FILE *f = fopen(env_.init_communities_fname.c_str(), "r");
  if (!f) {
    printf("cannot open init communities file:%s\n", strerror(errno));
    exit(-1);
  }
  printf("+ Loading init communities from %s\n", env_.init_communities_fname.c_str());
  std::vector<std::vector<uint32_t>> member(n_);        // _init_communities_seq: communities of every node, file order
  uint32_t cid = 0;
  {
    char *line = nullptr;
    size_t cap = 0;
    ssize_t len;
    while ((len = getline(&line, &cap, f)) > 0) {
      // (the reference scans "%[^\n]" first: an empty line keeps the PREVIOUS line's text in its buffer and repeats
      // that community under a new id; such files are refused here instead of reproducing the accident)
      char *p = line, *e = nullptr;
      bool any = false;
      for (;; p = e) {
        const long u = strtol(p, &e, 10);
        if (p == e) break;
        uint32_t seq;
        if (u < 0 || !net_.id2seq((uint32_t)u, &seq)) {
          fprintf(stderr, "error: init-communities id %ld not found in original network\n", u);
          exit(-1);
        }
        if (seq < n_) member[seq].push_back(cid);
        any = true;
      }
      if (!any) {
        int c = fgetc(f);
        if (c == EOF) break;                       // trailing blank line
        fprintf(stderr, "error: line %u of %s lists no node\n", cid + 1, env_.init_communities_fname.c_str());
        exit(-1);
      }
      cid++;
    }
    free(line);
  }
  fclose(f);
  printf("+ Loaded %d init communities\n", cid);
  if (FILE *g = fopen(env_.file("/init_memberships.txt").c_str(), "w")) {
    for (uint32_t i = 0; i < n_; ++i) {
      fprintf(g, "%d\t", net_.seq2id(i));
      for (uint32_t c : member[i]) fprintf(g, "%d\t", c);
      fprintf(g, "\n");
    }
    fclose(g);
  }
  for (uint32_t i = 0; i < n_; ++i)
    for (uint32_t c : member[i])
      if (c >= k_) {   // the reference only logs this and then writes past the end of phi (:437-439)
        fprintf(stderr, "error: init community %u of node %d, but -k is %u\n", c, net_.seq2id(i), k_);
        exit(-1);
      }
  const uint32_t k = k_;
  const unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([&, t] {
      std::vector<double> phi(k);
      const uint32_t v0 = (uint32_t)((uint64_t)n_ * t / nt), v1 = (uint32_t)((uint64_t)n_ * (t + 1) / nt);
      for (uint32_t p = v0; p < v1; ++p) {
        double *g = &gamma_[(size_t)p * k];
        for (uint32_t c = 0; c < k; ++c) g[c] = env_.alpha;                 // _gamma.set_elements(alpha)
        for (uint32_t c = 0; c < k; ++c) phi[c] = env_.alpha;
        for (uint32_t c : member[p]) phi[c] += (double)n_ / member[p].size();
        double s = .0;                                                       // D1Array::normalize, matrix.hh:341-346
        for (uint32_t c = 0; c < k; ++c) s += phi[c];
        for (uint32_t c = 0; c < k; ++c) phi[c] = phi[c] / s;
        const size_t deg = net_.get_edges(p).size();
        for (size_t r = 0; r < deg; ++r)                                     // _gamma.add_slice(p, phi), once per entry
          for (uint32_t c = 0; c < k; ++c) g[c] += phi[c];
      }
    });
  for (auto &x : th) x.join();
}

void LinkSampling::init_gamma2() {
  // Reference (:374-401): for every link in (p, adjacency) order draw K uniforms, normalise, add the vector to
  // both endpoints' rows -- one serial loop, 2*K*8 bytes of read-modify-write per link on random rows (122 s at
  // 1e8 links, K = 200).  Same values, same order of additions per row, spread over the host threads:
  //   producer  : the mt19937 stream for a chunk of links (inherently serial, but only the generator)
  //   phase A   : per link, the reference's sum (index order) and division          -- parallel over links
  //   phase B   : per NODE RANGE, walk the chunk's links in order and add into the rows of that range
  //               (a row only ever sees its additions in link order)                 -- parallel over ranges
  const uint32_t k = k_;
  std::vector<uint32_t> lp, lq;
  for (uint32_t p = 0; p < n_; ++p)
    for (uint32_t q : net_.get_edges(p))
      if (p < q) { lp.push_back(p); lq.push_back(q); }
  const size_t nl = lp.size();
  const size_t chunk = std::max<size_t>(256, std::min<size_t>(1u << 14, (size_t)(1u << 22) / std::max(1u, k)));
  const unsigned nt = std::max(1u, std::min(16u, std::min<unsigned>(std::thread::hardware_concurrency(),
                                                                    (unsigned)(nl / 4096 + 1))));
  Pool pool(nt);
  std::vector<std::vector<uint32_t>> todo_of(nt);
  auto work = [&](double *u, size_t l0, size_t cnt, bool normalise) {
    if (normalise) pool.run([&](unsigned t) {        // phase A
      for (size_t i = cnt * t / nt; i < cnt * (t + 1) / nt; ++i) {
        double *phi = u + i * k;
        double s = .0;
        for (uint32_t c = 0; c < k; ++c) s += phi[c];
        for (uint32_t c = 0; c < k; ++c) phi[c] = phi[c] / s;
      }
    });
    pool.run([&](unsigned t) {                       // phase B
      const uint32_t v0 = (uint32_t)((uint64_t)n_ * t / nt), v1 = (uint32_t)((uint64_t)n_ * (t + 1) / nt);
      // the additions this thread owns, in link order; then applied with the rows of the next few prefetched (the
      // rows are 8*K bytes at random places of an n*K*8-byte matrix: without the prefetch every row starts with a
      // TLB + DRAM miss the adds then wait for -- 21 s of a 39 s run at n = 1e6, K = 200)
      std::vector<uint32_t> &todo = todo_of[t];
      todo.clear();
      for (size_t i = 0; i < cnt; ++i) {
        const uint32_t p = lp[l0 + i], q = lq[l0 + i];
        if (p >= v0 && p < v1) { todo.push_back(p); todo.push_back((uint32_t)i); }
        if (q >= v0 && q < v1) { todo.push_back(q); todo.push_back((uint32_t)i); }
      }
      const size_t m2 = todo.size() / 2, ahead = 6;
      for (size_t j = 0; j < m2; ++j) {
        if (j + ahead < m2) {
          const char *nx = reinterpret_cast<const char *>(&gamma_[(size_t)todo[2 * (j + ahead)] * k]);
          for (size_t b = 0; b < (size_t)k * 8; b += 256) __builtin_prefetch(nx + b, 1, 1);
        }
        double *g = &gamma_[(size_t)todo[2 * j] * k];
        const double *phi = u + (size_t)todo[2 * j + 1] * k;
        for (uint32_t c = 0; c < k; ++c) g[c] += phi[c];
      }
    });
  };
  const size_t nchunks = (nl + chunk - 1) / chunk;
  const size_t cw = chunk * k;                        // words (uniforms) per full chunk
  // ---- many producers: disjoint pieces of the ONE mt19937 stream, reached by jump-ahead (mt_jump.hh) ----
  // Producer t makes chunks t, t+T, t+2T, ...; chunk c starts c*cw words into the stream.  The first chunk comes from
  // rng_ itself (it may sit in the middle of a block); the start state of every later chunk is a 624-word history
  // window, moved on by t^(cw) or t^(T*cw) mod phi.  The appliers still see the chunks in order, so every row sees
  // its additions in link order -- the result is bit-identical to the serial loop.
  const char *fp = getenv("SVINET_INIT_PRODUCERS");
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  unsigned T = fp ? (unsigned)std::max(1, atoi(fp)) : std::max(1u, std::min(8u, hw / 2));
  if (!fp && (double)nl * k < 4e8) T = 1;            // below ~0.5 s of generation the set-up does not pay
  if (nchunks < 2 * (size_t)T || cw < 2 * 624 || !mtjump::ready()) T = 1;
  const size_t R = 2 * (size_t)T;                    // ring of chunk buffers
  std::vector<std::vector<double>> ring(R, std::vector<double>(cw));
  std::mutex m;
  std::condition_variable cv;
  std::vector<size_t> ready(R, (size_t)-1);          // ready[b] = chunk held by buffer b
  size_t consumed = 0;                                // chunks applied so far
  Mt19937 last_state = rng_;
  std::vector<std::vector<uint32_t>> start(T, std::vector<uint32_t>(624));
  mtjump::Poly g_skip;
  Mt19937 first = rng_;                               // producer 0 generates chunk 0 from the live generator
  if (T > 1) {
    Mt19937 probe = rng_;
    const size_t r = probe.remaining_in_block();      // words before the next block boundary (< 624 <= cw)
    std::vector<double> scratch(r + 1);
    probe.uniform_fill(scratch.data(), r);            // now at a boundary: its array is the history window H0
    uint32_t h[624];
    probe.history(h);
    mtjump::Poly g_first, g_chunk;
    std::thread a([&] { g_first = mtjump::power_of_t(cw - r); });
    std::thread b([&] { g_chunk = mtjump::power_of_t(cw); });
    g_skip = mtjump::power_of_t((uint64_t)T * cw);
    a.join();
    b.join();
    mtjump::apply(g_first, h);                        // history at the start of chunk 1
    for (unsigned t = 1; t < T; ++t) {
      std::copy(h, h + 624, start[t].begin());
      mtjump::apply(g_chunk, h);                      // ... of chunk t+1
    }
    std::copy(h, h + 624, start[0].begin());          // chunk T: producer 0's second chunk
  }
  auto produce = [&](unsigned t) {
    Mt19937 gen = first;
    for (size_t c = t; c < nchunks; c += T) {
      {
        std::unique_lock<std::mutex> g(m);
        cv.wait(g, [&] { return c < consumed + R; }); // buffer c % R is free once chunk c - R has been applied
      }
      if (!(T == 1 || (t == 0 && c == 0))) gen.set_history(start[t].data());
      const size_t cnt = std::min(chunk, nl - c * chunk);
      double *u = ring[c % R].data();
      gen.uniform_fill(u, cnt * k);
      if (c == nchunks - 1) last_state = gen;         // the stream position after init_gamma2, as in the serial loop
      if (T > 1) {
        for (size_t i = 0; i < cnt; ++i) {            // phase A here: the producers are parallel, the pool only applies
          double *phi = u + i * k;
          double sum = .0;
          for (uint32_t cc = 0; cc < k; ++cc) sum += phi[cc];
          for (uint32_t cc = 0; cc < k; ++cc) phi[cc] = phi[cc] / sum;
        }
        if (c + T < nchunks && !(t == 0 && c == 0)) mtjump::apply(g_skip, start[t].data());
      }
      {
        std::lock_guard<std::mutex> g(m);
        ready[c % R] = c;
      }
      cv.notify_all();
    }
  };
  std::vector<std::thread> producers;
  for (unsigned t = 0; t < T; ++t) producers.emplace_back(produce, t);
  for (size_t c = 0; c < nchunks; ++c) {
    {
      std::unique_lock<std::mutex> g(m);
      cv.wait(g, [&] { return ready[c % R] == c; });
    }
    work(ring[c % R].data(), c * chunk, std::min(chunk, nl - c * chunk), T == 1);
    {
      std::lock_guard<std::mutex> g(m);
      consumed = c + 1;
    }
    cv.notify_all();
  }
  for (auto &p : producers) p.join();
  rng_ = last_state;
}

int LinkSampling::load_model() {
  // <dir>gamma.txt: "seq \t id \t g_0 .. g_K-1"; <dir>lambda.txt: "k \t l_0 \t l_1"  
  if (load_numeric_rows(env_.gamma_location + "gamma.txt", n_, k_, 2, gamma_.data(), "gamma.txt") < 0) return -1;
  if (load_numeric_rows(env_.gamma_location + "lambda.txt", k_, 2, 1, lambda_.data(), "lambda.txt") < 0) return -1;
  return 0;
}

void LinkSampling::assign_training_links() {
  // :493-523.  tl[v] counts v's non-held-out adjacency entries TWICE (the reference bumps both endpoints for
  // every adjacency direction); links_ lists the kept (p<q) pairs in (p, adjacency) order.  Node
  // ranges are independent, so they run on the host threads and their link lists are concatenated in order.
  training_links_.assign(n_, 0.0);
  links_.clear();
  const unsigned nt = std::max(1u, std::min(16u, std::min<unsigned>(std::thread::hardware_concurrency(), n_ / 4096 + 1)));
  std::vector<std::vector<uint32_t>> part(nt);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([&, t] {
      const uint32_t v0 = (uint32_t)((uint64_t)n_ * t / nt), v1 = (uint32_t)((uint64_t)n_ * (t + 1) / nt);
      for (uint32_t p = v0; p < v1; ++p) {
        uint32_t kept = 0;
        for (uint32_t q : net_.get_edges(p)) {
          if (!env_.accuracy) {
            Edge e(p, q);
            Network::order_edge(e);
            if (!edge_ok(e)) continue;   // held out
          }
          ++kept;
          if (p < q) { part[t].push_back(p); part[t].push_back(q); }
        }
        training_links_[p] = 2.0 * kept;
      }
    });
  for (auto &x : th) x.join();
  size_t total = 0;
  for (auto &v : part) total += v.size();
  links_.reserve(total);
  for (auto &v : part) links_.insert(links_.end(), v.begin(), v.end());
}

bool LinkSampling::validation_likelihood() {
  if (env_.accuracy) return false;
  // -gpus N: every shard evaluates a slice of the pairs (rows of other shards are peer loads inside the library)
  const size_t ng = devs_.size(), np = hp_.size();
  for (size_t i = 0; i < ng; ++i) {
    const size_t a = np * i / ng, b = np * (i + 1) / ng;
    DEV(svi_ls_heldout(devs_[i], b - a, hp_.data() + a, hq_.data() + a, hy_.data() + a, env_.epsilon, hll_.data() + a));
  }
  uint32_t k = 0, kzeros = 0, kones = 0;
  double s = .0, szeros = 0, sones = 0;
  for (size_t i = 0; i < hll_.size(); ++i) {
    const double u = hll_[i];
    s += u;
    k += 1;
    if (hy_[i]) { sones += u; kones++; } else { szeros += u; kzeros++; }
  }
  const double nshol = (zeros_prob_ * (szeros / kzeros)) + (ones_prob_ * (sones / kones));
  fprintf(vf_, "%d\t%d\t%.9f\t%d\t%.9f\t%d\t%.9f\t%d\t%.9f\t%.9f\t%.9f\n", iter_, duration(), s / k, k,
          szeros / kzeros, kzeros, sones / kones, kones, zeros_prob_ * (szeros / kzeros),
          ones_prob_ * (sones / kones), nshol);
  fflush(vf_);

  // stop state machine, :1006-1049
  const double a = nshol;
  bool stop = false;
  int why = -1;
  if (iter_ > 10) {
    if (a > prev_h_ && prev_h_ != 0 && fabs((a - prev_h_) / prev_h_) < 0.00001) {
      stop = true;
      why = 100;
    } else if (a < prev_h_) {
      nh_++;
    } else if (a > prev_h_) {
      nh_ = 0;
    }
    if (a > max_h_) { max_h_ = a; max_t_ = 0; }
    if (nh_ > 2) { why = 1; stop = true; }
  }
  prev_h_ = nshol;
  if (FILE *f = fopen(env_.file("/max.txt").c_str(), "w")) {
    fprintf(f, "%d\t%d\t%.5f\t%.5f\t%.5f\t%d\n", iter_, duration(), a, max_t_, max_h_, why);
    fclose(f);
  }
  if (annealing_ && stop) {        // the first "stop" only ends the annealing phase
    annealing_ = false;
    nh_ = 0;
    prev_h_ = 0;
    return false;
  }
  return stop && env_.use_validation_stop;
}

void LinkSampling::test_likelihood_line() {
  if (env_.accuracy) return;
  // the reference evaluates its (empty) test map every report and logs the resulting 0/0 (Q11)
  const double nan = std::nan("");
  fprintf(tf_, "%d\t%d\t%.9f\t%d\t%.9f\t%d\t%.9f\t%d\t%.9f\t%.9f\t%.9f\n", iter_, duration(), -nan, 0, -nan, 0, -nan, 0,
          -nan, -nan, -nan);
  fflush(tf_);
}

void LinkSampling::fetch_state() {
  if (devs_.size() > 1)   // every shard sends its gamma rows to the others; shard 0 then holds the whole matrix
    for (svi_ls *d : devs_) DEV(svi_ls_mg_publish_gamma(d));
  DEV(svi_ls_get_state(dev_, gamma_.data(), lambda_.data()));
}

void LinkSampling::save_model() {
  FILE *gf = open_or_die(env_.file("/gamma.txt"), "w", "gamma");
  const uint32_t k = k_;
  write_rows(gf, n_, [&](uint32_t i, std::string &s) {
    char b[48];
    s.append(b, (size_t)snprintf(b, sizeof b, "%d\t%d\t", i, net_.seq2id(i)));
    const double *g = &gamma_[(size_t)i * k];
    for (uint32_t c = 0; c < k; ++c) append_fixed(s, g[c], 5, c == k - 1 ? '\n' : '\t');
  });
  fclose(gf);
  FILE *lf = open_or_die(env_.file("/lambda.txt"), "w", "lambda");
  for (uint32_t c = 0; c < k_; ++c) fprintf(lf, "%d\t%.5f\t%.5f\n", c, lambda_[2 * c], lambda_[2 * c + 1]);
  fclose(lf);
}

void LinkSampling::write_groups() {
  FILE *f = open_or_die(env_.file("/groups.txt"), "w", "groups");
  const uint32_t k = k_;
  write_rows(f, n_, [&](uint32_t i, std::string &s) {
    char b[48];
    s.append(b, (size_t)snprintf(b, sizeof b, "%d\t%d\t", i, net_.seq2id(i)));
    const double *g = &gamma_[(size_t)i * k];
    double sum = .0;
    for (uint32_t c = 0; c < k; ++c) sum += g[c];
    for (uint32_t c = 0; c < k; ++c) append_fixed(s, g[c] / sum, 3, c == k - 1 ? '\n' : '\t');
  });
  fclose(f);
}

void LinkSampling::write_communities(const std::string &name) {
  // one line per non-empty link community, ascending k: external ids ascending, each followed by ' '.
  // Runs every report (every iteration with the default reportfreq): nodes are visited in external-id order
  // (a permutation computed once), so every community's list comes out sorted without a per-report sort.
  const uint32_t words = (k_ + 31) / 32;
  if (member_bits_.size() != (size_t)n_ * words) member_bits_.assign((size_t)n_ * words, 0);
  if (have_membership_) {
    if (devs_.size() == 1) DEV(svi_ls_get_membership(dev_, member_bits_.data()));
    else   // every shard holds the membership words of its own node block
      for (size_t i = 0; i < devs_.size(); ++i)
        DEV(svi_ls_get_membership_rows(devs_[i], bounds_[i], bounds_[i + 1] - bounds_[i],
                                       member_bits_.data() + (size_t)bounds_[i] * words));
  }
  if (by_id_.size() != n_) {
    by_id_.resize(n_);
    for (uint32_t p = 0; p < n_; ++p) by_id_[p] = p;
    std::sort(by_id_.begin(), by_id_.end(), [&](uint32_t a, uint32_t b) { return net_.seq2id(a) < net_.seq2id(b); });
    // "<id> " of every node, formatted once
    id_text_.clear();
    id_off_.assign((size_t)n_ + 1, 0);
    char b[16];
    for (uint32_t i = 0; i < n_; ++i) {
      id_text_.append(b, (size_t)snprintf(b, sizeof b, "%d ", net_.seq2id(by_id_[i])));
      id_off_[i + 1] = (uint32_t)id_text_.size();
    }
  }
  // every thread scans a slice of the id-ordered nodes into its own per-community buffers; a community's line is
  // the concatenation of the slices in order, so the ids stay ascending
  const unsigned nt = std::max(1u, std::min(16u, std::min(std::thread::hardware_concurrency(), n_ / 4096 + 1)));
  std::vector<std::vector<std::string>> part(nt, std::vector<std::string>(k_));
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        const uint32_t i0 = (uint32_t)((uint64_t)n_ * t / nt), i1 = (uint32_t)((uint64_t)n_ * (t + 1) / nt);
        std::vector<std::string> &line = part[t];
        for (uint32_t i = i0; i < i1; ++i) {
          const uint32_t *w = &member_bits_[(size_t)by_id_[i] * words];
          for (uint32_t wi = 0; wi < words; ++wi) {
            uint32_t bits = w[wi];
            while (bits) {
              const uint32_t c = wi * 32 + (uint32_t)__builtin_ctz(bits);
              bits &= bits - 1;
              if (c < k_) line[c].append(id_text_, id_off_[i], id_off_[i + 1] - id_off_[i]);
            }
          }
        }
      });
    for (auto &x : th) x.join();
  }
  FILE *f = open_or_die(env_.file(name), "w", "communities");
  for (uint32_t c = 0; c < k_; ++c) {
    size_t len = 0;
    for (unsigned t = 0; t < nt; ++t) len += part[t][c].size();
    if (!len) continue;
    for (unsigned t = 0; t < nt; ++t) fwrite(part[t][c].data(), 1, part[t][c].size(), f);
    fputc('\n', f);
  }
  fclose(f);
}

// -nmi <file>: "<node id>\t<community> <community> ..." per line (Network::load_ground_truth, network.cc:254-307);
// ground_truth.txt / ground_truth_community_sizes.txt as write_gt_communities (:508-536) writes them: communities in
// ascending id order, members in file order.
void LinkSampling::load_ground_truth() {
  FILE *f = fopen(env_.ground_truth_fname.c_str(), "r");
  if (!f) {
    fprintf(stderr, "error: cannot read ground truth file %s; check path; skipping file\n", env_.ground_truth_fname.c_str());
    return;
  }
  std::map<uint32_t, std::vector<uint32_t>> by_id, by_seq;
  char *line = nullptr;
  size_t cap = 0;
  while (getline(&line, &cap, f) > 0) {
    char *e = nullptr;
    const long nid = strtol(line, &e, 10);
    if (e == line) continue;
    uint32_t seq;
    if (nid < 0 || !net_.id2seq((uint32_t)nid, &seq)) {
      fprintf(stderr, "error: ground truth node %ld is not in the network\n", nid);
      exit(-1);
    }
    for (char *p = e;; p = e) {
      const long u = strtol(p, &e, 10);
      if (p == e) break;
      by_id[(uint32_t)u].push_back((uint32_t)nid);
      by_seq[(uint32_t)u].push_back(seq);
    }
  }
  free(line);
  fclose(f);
  printf("+ Done loading ground truth\n");
  FILE *g = open_or_die(env_.file("/ground_truth.txt"), "w", "ground truth");
  FILE *sz = open_or_die(env_.file("/ground_truth_community_sizes.txt"), "w", "ground truth sizes");
  uint32_t c = 0;
  for (const auto &kv : by_id) {
    fprintf(sz, "%d\t%ld\n", c++, (long)kv.second.size());
    for (uint32_t v : kv.second) fprintf(g, "%d ", v);
    fprintf(g, "\n");
  }
  fclose(g);
  fclose(sz);
  for (auto &kv : by_seq) gt_communities_.push_back(std::move(kv.second));
}

void LinkSampling::log_communities() {
  write_communities("/communities.txt");
  if (env_.nmi && !gt_communities_.empty()) {
    // the reference appends the output of the external `mutual` binary here (:843-851); nmi.hh is that measure
    const uint32_t words = (k_ + 31) / 32;
    std::vector<std::vector<uint32_t>> found(k_);
    if (member_bits_.size() == (size_t)n_ * words)
      for (uint32_t p = 0; p < n_; ++p)
        for (uint32_t c = 0; c < k_; ++c)
          if (member_bits_[(size_t)p * words + c / 32] >> (c % 32) & 1u) found[c].push_back(p);
    if (FILE *f = fopen(env_.file("/mutual.txt").c_str(), "a")) {
      fprintf(f, "mutual3:\t%g\n", nmi::lfk(n_, gt_communities_, found));
      fclose(f);
    }
  }
}

void LinkSampling::do_on_stop() {
  log_communities();
  fetch_state();
  save_model();
  write_groups();
}

void LinkSampling::infer() {
  bool write_comm = false;
  const uint32_t rf = (uint32_t)env_.reportfreq;
  Lap lap;
  double t_step = 0, t_report = 0, t_heldout = 0;
  auto since = [](std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
  };
  while (1) {
    if (env_.max_iterations && iter_ > env_.max_iterations) {
      printf("+ Quitting: reached max iterations.\n");
      env_.plog("maxiterations reached", true);
      env_.terminate = true;
      if (lap.on)
        fprintf(stderr, "[linksampling] %u iterations: device steps %.3f s, reports %.3f s (held-out %.3f s, communities.txt %.3f s)\n",
                iter_, t_step, t_report, t_heldout, t_report - t_heldout);
      lap("iterations");
      do_on_stop();
      lap("do_on_stop (writers)");
      exit(0);
    }
    if (env_.max_iterations == 1) write_comm = true;
    printf("\riteration %d: processing %zu links", iter_, links_.size() / 2);
    fflush(stdout);
    auto t0 = std::chrono::steady_clock::now();
    device_step(write_comm);
    if (lap.on) { device_sync(); t_step += since(t0); }
    if (write_comm) have_membership_ = true;

    if (env_.terminate) {             // SIGTERM: dump the model and carry on (:763-766)
      do_on_stop();
      env_.terminate = false;
    }
    write_comm = (iter_ % rf == rf - 1);
    if (iter_ % rf == 0) {
      t0 = std::chrono::steady_clock::now();
      if (validation_likelihood()) {
        do_on_stop();
        exit(0);
      }
      test_likelihood_line();
      t_heldout += since(t0);
      log_communities();
      t_report += since(t0);
    }
    iter_++;
  }
}

void LinkSampling::dump_init(const std::string &dir) const {
  auto dump = [&](const char *name, const void *p, size_t bytes) {
    FILE *f = open_or_die(dir + "/" + name, "wb", name);
    if (bytes) fwrite(p, 1, bytes, f);
    fclose(f);
  };
  std::vector<uint32_t> vp;
  for (const Edge &e : validation_pairs_) { vp.push_back(e.first); vp.push_back(e.second); }
  dump("gamma.f64", gamma_.data(), gamma_.size() * sizeof(double));
  dump("lambda.f64", lambda_.data(), lambda_.size() * sizeof(double));
  dump("validation.u32", vp.data(), vp.size() * sizeof(uint32_t));
  dump("links.u32", links_.data(), links_.size() * sizeof(uint32_t));
  dump("tl.f64", training_links_.data(), training_links_.size() * sizeof(double));
}
