// network.hh -- graph ingest for the svinet drop-in CLI.
//
// Behaviour follows the reference's Network::read (src/network.cc:11-159): tab/space separated integer
// pairs, external ids mapped to dense sequence ids in first-appearance order, self-loops and repeated
// (or reversed) pairs dropped (the FIRST occurrence is kept), per-node neighbour lists in insertion order
// (that order defines the reference's _links order and the RNG consumption of init_gamma2), synthetic ids
// 100000+k for the "single" nodes that pad the graph up to -n (src/network.cc:107-113).
//
// The reference does this with fscanf, two std::map's and an O(degree) duplicate scan per line -- minutes
// to hours at 1e8 links.  Here: the file is mmap'ed and parsed by all host
// threads; ids are mapped in one sequential hash pass (first-appearance order is inherently sequential);
// duplicates are found by SORTING the (min,max) keys of the accepted lines, partitioned by the smaller
// endpoint so that every thread sorts its own bucket; the adjacency is a CSR filled by a counting pass that
// preserves line order, with a second, per-node sorted copy for O(log deg) membership tests.
#ifndef SVINET_B200_NETWORK_HH
#define SVINET_B200_NETWORK_HH

#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "env.hh"

typedef std::pair<uint32_t, uint32_t> Edge;   // always (first < second)

// a node's neighbour list (view into the CSR)
struct NeighbourSpan {
  const uint32_t *b, *e;
  const uint32_t *begin() const { return b; }
  const uint32_t *end() const { return e; }
  size_t size() const { return (size_t)(e - b); }
  uint32_t operator[](size_t i) const { return b[i]; }
};

class Network {
 public:
  explicit Network(Env &env) : env_(env) {}

  int read(const std::string &path);                 // src/network.cc:11-159

  uint32_t n() const { return n_; }                               // the -n argument
  uint32_t ones() const { return (uint32_t)edges_.size(); }
  uint32_t singles() const { return singles_; }
  // neighbours of `a` in insertion (line) order, Network::get_edges (src/network.hh:150-156)
  NeighbourSpan get_edges(uint32_t a) const { return NeighbourSpan{adj_.data() + off_[a], adj_.data() + off_[a + 1]}; }
  const std::vector<Edge> &edges() const { return edges_; }
  bool y(uint32_t a, uint32_t b) const;              // src/network.hh:158-176
  uint32_t seq2id(uint32_t seq) const { return seq2id_[seq]; }
  bool id2seq(uint32_t id, uint32_t *seq) const;
  void deg_stats(uint32_t &max, double &avg) const;  // src/network.cc:204-220

  static void order_edge(Edge &e) { if (e.first > e.second) std::swap(e.first, e.second); }
  static const uint32_t SINGLE_NODE_START_ID = 100000;

 private:
  bool add(uint32_t id);                             // src/network.hh:134-148
  void set_env_variables();                          // src/network.cc:223-250

  Env &env_;
  uint32_t n_ = 0;
  std::vector<uint64_t> off_;                        // [n+1]
  std::vector<uint32_t> adj_, adj_sorted_;           // [2*ones] insertion order / sorted per node
  std::vector<Edge> edges_;                          // accepted links in line order
  std::vector<uint32_t> seq2id_;
  static constexpr uint32_t kNoSeq = 0xffffffffu;
  std::vector<uint32_t> id_table_;                   // external id -> seq for small ids (flat), else id2seq_
  std::unordered_map<uint32_t, uint32_t> id2seq_;
  uint32_t curr_seq_ = 0, singles_ = 0;
};

#endif
