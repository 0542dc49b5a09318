// linksampling.hh -- the drop-in `LinkSampling` of the svinet CLI, B200 build.
//
// Same seam as the reference class used at src/main.cc:337-341:
//     LinkSampling ls(env, network);  ls.infer();          // infer() ends the process with exit(0)
// The constructor keeps every host responsibility of the reference's (src/linksampling.cc:5-155): RNG,
// held-out draw, gamma/lambda initialisation, output files.  infer() (src/linksampling.cc:557-790)
// hands the whole loop body to the device through the C ABI in include/svi_ls.h and keeps the stop
// state machine, the SIGTERM poll and the file writers on the host.
#ifndef SVINET_B200_LINKSAMPLING_HH
#define SVINET_B200_LINKSAMPLING_HH

#include <cstdint>
#include <cstdio>
#include <ctime>
#include <string>
#include <unordered_set>
#include <vector>

#include "env.hh"
#include "network.hh"
#include "rng.hh"
#include "svi_ls.h"

class LinkSampling {
 public:
  LinkSampling(Env &env, Network &network);
  ~LinkSampling();

  void infer();        // never returns normally (exit(0) on stop / max-iterations), like the reference
  void save_model();   // gamma.txt + lambda.txt (src/linksampling.cc:805-837)

  // test hook (not in the reference): dump the initial state instead of running, see main.cc -dump-init
  void dump_init(const std::string &dir) const;

 private:
  // --- start-up, host side ---
  void init_validation();                 // :165-188
  void load_validation();                 // :1383-1414
  void set_validation_sample(int s);      // :281-309
  void get_random_edge(bool link, Edge &e);   // src/linksampling.hh:328-349
  bool edge_ok(const Edge &e) const;      // src/linksampling.hh:296-326
  void init_gamma2();                     // :374-401
  void init_gamma_external();             // :404-452 (-init-communities) + Network::load_init_communities, network.cc:374-437
  int load_model();                       // :1266-1352
  void assign_training_links();           // :493-523
  // --- per report ---
  bool validation_likelihood();           // :966-1050, true = the run must end
  void test_likelihood_line();            // :1147-1182 on the (always empty) test set
  void log_communities();                 // :839-852
  void load_ground_truth();               // Network::load_ground_truth + write_gt_communities, network.cc:254-307,508-536
  std::vector<std::vector<uint32_t>> gt_communities_;   // -nmi: ground-truth communities as seq ids
  void write_communities(const std::string &name);   // :882-917
  void write_groups();                    // :1452-1476
  void do_on_stop();                      // :792-802
  void fetch_state();                     // device -> gamma_/lambda_
  uint32_t duration() const { return (uint32_t)(time(0) - start_time_); }

  Env &env_;
  Network &net_;
  uint32_t n_, k_;
  uint32_t iter_ = 0;                     // the reference never initialises it; observed 0
  double total_pairs_ = 0, ones_prob_ = 0, zeros_prob_ = 0;
  Mt19937 rng_;
  std::vector<double> gamma_, lambda_;    // host mirrors (n*k, k*2)
  std::vector<Edge> validation_pairs_;    // draw order
  std::vector<Edge> validation_sorted_;   // std::map<Edge,bool> iteration order
  std::unordered_set<uint64_t> held_keys_; // membership test of the held-out set (first << 32 | second)
  std::vector<uint32_t> links_;           // training links (p<q), reference order
  std::vector<double> training_links_;    // tl[p] = 2 x training degree
  std::vector<uint32_t> member_bits_;     // last write_comm tally fetched from the device
  std::vector<uint32_t> by_id_;           // nodes in ascending external-id order (communities.txt order)
  std::string id_text_;                   // their "<id> " strings, back to back
  std::vector<uint32_t> id_off_;
  bool have_membership_ = false;
  bool annealing_ = true;
  double max_t_ = -2147483647, max_h_ = -2147483647, prev_h_ = -2147483647;
  uint32_t nh_ = 0;
  time_t start_time_;
  FILE *vf_ = nullptr, *tf_ = nullptr, *lf_ = nullptr;
  svi_ls *dev_ = nullptr;                 // shard 0 (the only one without -gpus N)
  std::vector<svi_ls *> devs_;            // -gpus N: one handle per GPU, node-block shards (include/svi_ls.h, svi_ls_mg_step)
  std::vector<uint32_t> bounds_;          // [ngpus+1] node blocks of the shards
  void create_device();
  void device_step(bool write_comm);
  void device_sync();
  // held-out pairs in evaluation order, flattened for svi_ls_heldout
  std::vector<uint32_t> hp_, hq_;
  std::vector<uint8_t> hy_;
  std::vector<double> hll_;
};

#endif
