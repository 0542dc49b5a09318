// network.cc -- see network.hh.
#include "network.hh"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>

bool Network::add(uint32_t id) {
  if (curr_seq_ >= env_.n) return false;   // more distinct ids than -n: the line is dropped
  id2seq_.emplace(id, curr_seq_);
  seq2id_[curr_seq_] = id;
  ++curr_seq_;
  return true;
}

bool Network::id2seq(uint32_t id, uint32_t *seq) const {
  auto it = id2seq_.find(id);
  if (it == id2seq_.end()) return false;
  *seq = it->second;
  return true;
}

bool Network::y(uint32_t a, uint32_t b) const {
  if (a > b) std::swap(a, b);
  for (uint32_t v : adj_[a])
    if (v == b) return true;
  return false;
}

void Network::accept_pair(uint32_t id1, uint32_t id2) {
  uint32_t p, q;
  if (!id2seq(id1, &p)) {
    if (!add(id1)) return;
    p = curr_seq_ - 1;
  }
  if (!id2seq(id2, &q)) {
    if (!add(id2)) return;
    q = curr_seq_ - 1;
  }
  if (p == q || y(p, q)) return;           // self-loop, or the pair (in either direction) was seen
  Edge e(p, q);
  order_edge(e);
  edges_.push_back(e);
  adj_[e.first].push_back(e.second);
  adj_[e.second].push_back(e.first);
  if (edges_.size() % 1000000 == 0) {
    printf("\r+ %zu entries", edges_.size());
    fflush(stdout);
  }
}

int Network::read(const std::string &path) {
  fprintf(stdout, "+ Reading network from %s\n", path.c_str());
  FILE *f = fopen(path.c_str(), "r");
  if (!f) {
    fprintf(stderr, "error: cannot open file %s:%s\n", path.c_str(), strerror(errno));
    exit(-1);
  }
  adj_.assign(env_.n, std::vector<uint32_t>());
  seq2id_.assign(env_.n, 0);
  id2seq_.reserve(env_.n * 2 + 16);
  // buffered integer scanner, same acceptance as fscanf("%d\t%d\n") on well-formed input
  std::vector<char> buf(1 << 22);
  uint64_t vals[2] = {0, 0};
  int have = 0;
  bool in_num = false, neg = false;
  size_t got;
  while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) {
    for (size_t i = 0; i < got; ++i) {
      const char c = buf[i];
      if (c >= '0' && c <= '9') {
        if (!in_num) { in_num = true; vals[have] = 0; }
        vals[have] = vals[have] * 10 + (uint64_t)(c - '0');
      } else {
        if (in_num) {
          if (neg) vals[have] = (uint64_t)(-(int64_t)vals[have]);
          in_num = false; neg = false;
          if (++have == 2) { accept_pair((uint32_t)vals[0], (uint32_t)vals[1]); have = 0; }
        }
        if (c == '-') neg = true;
        else if (c != ' ' && c != '\t' && c != '\n' && c != '\r') {
          printf("error: unexpected lines in file\n");
          exit(-1);
        }
      }
    }
  }
  if (in_num && ++have == 2) accept_pair((uint32_t)vals[0], (uint32_t)vals[1]);
  fclose(f);

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
  for (const auto &v : adj_) {
    const uint32_t d = (uint32_t)v.size();
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
