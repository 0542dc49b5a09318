// network.cc -- see network.hh.
#include "network.hh"

#include <algorithm>
#include <chrono>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

bool Network::add(uint32_t id) {
  if (curr_seq_ >= env_.n) return false;   // more distinct ids than -n: the line is dropped
  if (id < id_table_.size()) id_table_[id] = curr_seq_;
  else id2seq_.emplace(id, curr_seq_);
  seq2id_[curr_seq_] = id;
  ++curr_seq_;
  return true;
}

bool Network::id2seq(uint32_t id, uint32_t *seq) const {
  if (id < id_table_.size()) {
    if (id_table_[id] == kNoSeq) return false;
    *seq = id_table_[id];
    return true;
  }
  auto it = id2seq_.find(id);
  if (it == id2seq_.end()) return false;
  *seq = it->second;
  return true;
}

bool Network::y(uint32_t a, uint32_t b) const {
  if (a == b) return false;
  const uint32_t *lo = adj_sorted_.data() + off_[a], *hi = adj_sorted_.data() + off_[a + 1];
  return std::binary_search(lo, hi, b);
}

namespace {

// Parse the integers of [p, e) (one chunk of the mmap'ed file, cut at line ends) into `out`.
// Same acceptance as the reference's fscanf("%d\t%d\n") on well-formed input: integers separated by
// blanks / tabs / newlines, consumed two at a time.
bool parse_chunk(const char *p, const char *e, std::vector<uint32_t> &out) {
  uint64_t v = 0;
  bool in_num = false, neg = false;
  for (; p < e; ++p) {
    const char c = *p;
    if (c >= '0' && c <= '9') {
      if (!in_num) { in_num = true; v = 0; }
      v = v * 10 + (uint64_t)(c - '0');
    } else {
      if (in_num) {
        out.push_back((uint32_t)(neg ? (uint64_t)(-(int64_t)v) : v));
        in_num = false;
        neg = false;
      }
      if (c == '-') neg = true;
      else if (c != ' ' && c != '\t' && c != '\n' && c != '\r') return false;
    }
  }
  if (in_num) out.push_back((uint32_t)(neg ? (uint64_t)(-(int64_t)v) : v));
  return true;
}

template <class F>
void parallel_for(unsigned nt, F f) {
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; ++t) th.emplace_back(f, t);
  f(0u);
  for (auto &x : th) x.join();
}

}  // namespace

int Network::read(const std::string &path) {
  fprintf(stdout, "+ Reading network from %s\n", path.c_str());
  const bool timing = getenv("SVINET_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    const auto now = std::chrono::steady_clock::now();
    if (timing) fprintf(stderr, "[ingest] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
    t_last = now;
  };
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) {
    fprintf(stderr, "error: cannot open file %s:%s\n", path.c_str(), strerror(errno));
    exit(-1);
  }
  struct stat st;
  fstat(fd, &st);
  const size_t len = (size_t)st.st_size;
  const char *data = len ? (const char *)mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0) : "";
  if (len && data == MAP_FAILED) {
    fprintf(stderr, "error: cannot map file %s:%s\n", path.c_str(), strerror(errno));
    exit(-1);
  }
  n_ = env_.n;
  seq2id_.assign(n_, 0);
  id2seq_.reserve((size_t)n_ * 2 + 16);

  // ---- 1. parse: every thread takes a slice of the file cut at line ends ----
  const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(32, std::thread::hardware_concurrency()),
                                                                       len / (1u << 20) + 1));
  std::vector<size_t> cut(nt + 1, len);
  cut[0] = 0;
  for (unsigned t = 1; t < nt; ++t) {
    size_t c = len / nt * t;
    while (c < len && data[c] != '\n') ++c;
    cut[t] = c < len ? c + 1 : len;
  }
  std::vector<std::vector<uint32_t>> ints(nt);
  std::vector<char> ok(nt, 1);
  parallel_for(nt, [&](unsigned t) {
    ints[t].reserve((cut[t + 1] - cut[t]) / 6 + 16);
    ok[t] = parse_chunk(data + cut[t], data + cut[t + 1], ints[t]) ? 1 : 0;
  });
  for (unsigned t = 0; t < nt; ++t)
    if (!ok[t]) {
      printf("error: unexpected lines in file\n");
      exit(-1);
    }
  if (len) munmap((void *)data, len);
  close(fd);
  // an odd integer count inside a slice can only come from a malformed line; pair up over the whole stream
  size_t total = 0;
  for (auto &v : ints) total += v.size();
  std::vector<uint32_t> ids;
  if (nt == 1) {
    ids.swap(ints[0]);
  } else {
    ids.resize(total);
    size_t at = 0;
    for (auto &v : ints) {
      std::copy(v.begin(), v.end(), ids.begin() + at);
      at += v.size();
      std::vector<uint32_t>().swap(v);
    }
  }
  const size_t nlines = total / 2;
  // ids below this bound are mapped through a flat table (one load instead of a hash probe); the rest,
  // and the synthetic ids of single nodes, go through the hash map
  {
    uint32_t max_id = 0;
    for (size_t i = 0; i < total; ++i) max_id = std::max(max_id, ids[i]);
    const uint64_t bound = std::min<uint64_t>((uint64_t)max_id + 1, 8 * (uint64_t)total + (1u << 20));
    id_table_.assign((size_t)std::min<uint64_t>(bound, 1ull << 30), (uint32_t)kNoSeq);
  }
  lap("parse");

  // ---- 2. ids -> sequence ids in first-appearance order (sequential by nature); drop lines whose new
  //         id does not fit under -n, and self-loops ----
  std::vector<uint64_t> key;            // (min << 32 | max) of every candidate line, in line order
  key.reserve(nlines);
  for (size_t l = 0; l < nlines; ++l) {
    const uint32_t id1 = ids[2 * l], id2 = ids[2 * l + 1];
    uint32_t p, q;
    if (!id2seq(id1, &p)) {
      if (!add(id1)) continue;
      p = curr_seq_ - 1;
    }
    if (!id2seq(id2, &q)) {
      if (!add(id2)) continue;
      q = curr_seq_ - 1;
    }
    if (p == q) continue;
    key.push_back(((uint64_t)std::min(p, q) << 32) | std::max(p, q));
  }
  std::vector<uint32_t>().swap(ids);
  lap("id mapping");

  // ---- 3. duplicates: sort (key, line position) inside buckets of the smaller endpoint; the first
  //         occurrence of a key survives, exactly what the reference's "already linked?" scan keeps ----
  const size_t m = key.size();
  std::vector<uint8_t> keep(m, 1);
  {
    const unsigned nb = (unsigned)std::max<size_t>(1, std::min<size_t>(nt, m / 65536 + 1));
    const uint64_t span = ((uint64_t)n_ + nb - 1) / nb;
    std::vector<std::vector<std::pair<uint64_t, uint32_t>>> bucket(nb);
    if (m > 0xffffffffull) {
      fprintf(stderr, "error: more than 2^32 candidate lines\n");
      exit(-1);
    }
    // scatter (sequential, keeps line order inside a bucket; the sort below does not rely on it)
    std::vector<size_t> cnt(nb, 0);
    for (size_t i = 0; i < m; ++i) cnt[(key[i] >> 32) / span]++;
    for (unsigned b = 0; b < nb; ++b) bucket[b].reserve(cnt[b]);
    for (size_t i = 0; i < m; ++i) bucket[(key[i] >> 32) / span].emplace_back(key[i], (uint32_t)i);
    parallel_for(nb, [&](unsigned b) {
      auto &v = bucket[b];
      std::sort(v.begin(), v.end());
      for (size_t i = 1; i < v.size(); ++i)
        if (v[i].first == v[i - 1].first) keep[v[i].second] = 0;
    });
  }

  lap("duplicate sort");
  // ---- 4. edge list in line order + CSR in insertion order (counting pass) ----
  off_.assign((size_t)n_ + 1, 0);
  size_t kept = 0;
  for (size_t i = 0; i < m; ++i)
    if (keep[i]) {
      ++kept;
      off_[(key[i] >> 32) + 1]++;
      off_[(key[i] & 0xffffffffu) + 1]++;
    }
  for (uint32_t v = 0; v < n_; ++v) off_[v + 1] += off_[v];
  edges_.clear();
  edges_.reserve(kept);
  adj_.assign(2 * kept, 0);
  {
    std::vector<uint64_t> at(off_.begin(), off_.end() - 1);
    for (size_t i = 0; i < m; ++i)
      if (keep[i]) {
        const uint32_t lo = (uint32_t)(key[i] >> 32), hi = (uint32_t)key[i];
        edges_.emplace_back(lo, hi);
        adj_[at[lo]++] = hi;
        adj_[at[hi]++] = lo;
      }
  }
  lap("edge list + CSR");
  adj_sorted_ = adj_;
  parallel_for(nt, [&](unsigned t) {
    for (uint32_t v = t; v < n_; v += nt) std::sort(adj_sorted_.begin() + off_[v], adj_sorted_.begin() + off_[v + 1]);
  });
  lap("sorted adjacency");

  if (curr_seq_ != env_.n) {
    singles_ = env_.n - curr_seq_;
    printf("n = %d, curr_seq = %d\n", env_.n, curr_seq_);
    printf("+ Creating ids for %d single nodes\n", singles_);
    for (uint32_t c = curr_seq_, k = 0; c < env_.n; ++c, ++k) add(SINGLE_NODE_START_ID + k);
  }
  fprintf(stdout, "\n+ Done reading network\n");
  fflush(stdout);
  set_env_variables();
  return 0;
}

void Network::deg_stats(uint32_t &max, double &avg) const {
  max = 0;
  uint32_t s = 0, k = 0;                   // 32-bit sum, as in the reference
  for (uint32_t v = 0; v < n_; ++v) {
    const uint32_t d = (uint32_t)(off_[v + 1] - off_[v]);
    if (d > max) max = d;
    s += d;
    k++;
  }
  avg = (double)s / k;
}

void Network::set_env_variables() {
  const uint32_t n = env_.n;
  env_.total_pairs = (uint32_t)(n * (n - 1) / 2);   // 32-bit product on purpose
  env_.plog("total pairs", env_.total_pairs);
  env_.ones_prob = (double)ones() / env_.total_pairs;
  env_.zeros_prob = 1 - env_.ones_prob;
  env_.plog("ones_prob", env_.ones_prob);
  env_.plog("zeros_prob", env_.zeros_prob);
  if (env_.eta_type == "fromdata") {
    env_.eta0 = env_.total_pairs * env_.ones_prob / env_.k;
    env_.eta1 = env_.total_pairs * 1.0 / (env_.k * env_.k) - env_.eta0;
    if (env_.eta1 <= 0) env_.eta1 = 1.0;
  } else if (env_.eta_type == "uniform") {
    env_.eta0 = 1;
    env_.eta1 = 1;
  } else if (env_.eta_type == "sparse") {
    env_.eta0 = env_.eta0_sparse;
    env_.eta1 = env_.eta1_sparse;
  } else if (env_.eta_type == "dense") {
    env_.eta0 = env_.eta0_dense;
    env_.eta1 = env_.eta1_dense;
  } else {
    fprintf(stderr, "unknown eta_type\n");
    exit(-1);
  }
}
