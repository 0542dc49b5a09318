// network.hh -- graph ingest for the svinet drop-in CLI.
//
// Behaviour follows the reference's Network::read (src/network.cc:11-159): tab/space separated integer
// pairs, external ids mapped to dense sequence ids in first-appearance order, self-loops and repeated
// (or reversed) pairs dropped, per-node neighbour lists in insertion order (that order defines the
// reference's _links order and the RNG consumption of init_gamma2), synthetic ids 100000+k for the
// "single" nodes that pad the graph up to -n (src/network.cc:107-113).  Data structures are flat
// vectors + one hash map instead of the reference's std::map / vector-of-pointers.
#ifndef SVINET_B200_NETWORK_HH
#define SVINET_B200_NETWORK_HH

#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "env.hh"

typedef std::pair<uint32_t, uint32_t> Edge;   // always (first < second)

class Network {
 public:
  explicit Network(Env &env) : env_(env) {}

  int read(const std::string &path);                 // src/network.cc:11-159

  uint32_t n() const { return (uint32_t)adj_.size(); }            // the -n argument
  uint32_t ones() const { return (uint32_t)edges_.size(); }
  uint32_t singles() const { return singles_; }
  const std::vector<uint32_t> &get_edges(uint32_t a) const { return adj_[a]; }
  const std::vector<Edge> &edges() const { return edges_; }
  bool y(uint32_t a, uint32_t b) const;              // src/network.hh:158-176
  uint32_t seq2id(uint32_t seq) const { return seq2id_[seq]; }
  bool id2seq(uint32_t id, uint32_t *seq) const;
  void deg_stats(uint32_t &max, double &avg) const;  // src/network.cc:204-220

  static void order_edge(Edge &e) { if (e.first > e.second) std::swap(e.first, e.second); }
  static const uint32_t SINGLE_NODE_START_ID = 100000;

 private:
  bool add(uint32_t id);                             // src/network.hh:134-148
  void accept_pair(uint32_t id1, uint32_t id2);
  void set_env_variables();                          // src/network.cc:223-250

  Env &env_;
  std::vector<std::vector<uint32_t>> adj_;
  std::vector<Edge> edges_;
  std::vector<uint32_t> seq2id_;
  std::unordered_map<uint32_t, uint32_t> id2seq_;
  uint32_t curr_seq_ = 0, singles_ = 0;
};

#endif
