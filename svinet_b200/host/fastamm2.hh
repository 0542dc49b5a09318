// fastamm2.hh -- the drop-in `FastAMM2` of the svinet CLI (`-rnode -stratified`), B200 build.
//
// Same seam as the reference class used at src/main.cc:368-372:
//     FastAMM2 fastamm2(env, network);  fastamm2.infer();      // infer() ends the process with exit(0)
// The constructor keeps every host responsibility of the reference's (src/fastamm2.cc:8-248): GSL stream,
// shuffle_nodes, held-out draw, init_gamma / init_lambda, output files.  infer() (src/fastamm2.cc:535-702)
// draws each minibatch exactly like opt_process / opt_process_noninf (same mt19937 consumption) and hands
// the iteration body to the device through include/svi_fa2.h; with -device-draw the device draws the
// minibatches itself from a Philox stream (svi_fa2_run) and the host only reports.
#ifndef SVINET_B200_FASTAMM2_HH
#define SVINET_B200_FASTAMM2_HH

#include <cstdint>
#include <cstdio>
#include <ctime>
#include <string>
#include <unordered_set>
#include <vector>

#include "env.hh"
#include "network.hh"
#include "rng.hh"
#include "svi_fa2.h"

class FastAMM2 {
 public:
  FastAMM2(Env &env, Network &network);
  ~FastAMM2();

  void infer();        // never returns normally, like the reference
  void save_model();   // gamma.txt + lambda.txt (src/fastamm2.cc:705-739)

  void dump_init(const std::string &dir) const;   // test hook, see main.cc -dump-init

 private:
  void init_heldout();                        // :302-341
  void load_heldout();                        // :267-299
  void set_heldout_sample(int s);             // :393-421
  void get_random_edge(bool link, Edge &e);   // src/fastamm2.hh:543-565
  bool edge_ok(const Edge &e) const;          // src/fastamm2.hh:524-541
  void init_gamma();                          // :497-515
  void init_lambda();                         // :518-531
  int load_model();                           // :1717-1803
  void plan_links(std::vector<uint32_t> &pairs);      // opt_process pair selection, :936-960
  void plan_noninf(std::vector<uint32_t> &pairs);     // opt_process_noninf pair selection, :1078-1125
  void heldout_likelihood();                  // :1297-1392
  void compute_and_log_groups();              // :743-876 (after estimate_all_pi)
  void fetch_state();
  void finish_and_exit();
  uint32_t duration() const { return (uint32_t)(time(0) - start_time_); }

  Env &env_;
  Network &net_;
  uint32_t n_, k_;
  uint32_t iter_ = 0;                         // never initialised by the reference; observed 0
  uint64_t m_ = 10;                           // _m
  double inf_epsilon_ = 0.5, link_thresh_ = 0.9;
  double zeros_prob_ = 0, ones_prob_ = 0;     // never initialised by the reference; observed 0
  uint64_t total_pairs_sampled_ = 0;
  uint32_t start_node_ = 0;
  Mt19937 rng_;
  std::vector<double> gamma_, lambda_;
  std::vector<uint32_t> shuffled_;
  std::vector<Edge> heldout_pairs_, heldout_sorted_;
  std::unordered_set<uint64_t> held_keys_;    // membership test of the held-out set (first << 32 | second)
  std::vector<uint32_t> hp_, hq_;
  std::vector<uint8_t> hy_;
  std::vector<double> hll_;
  double prev_h_ = -2147483647, max_h_ = -2147483647;
  uint32_t nh_ = 0;
  time_t start_time_;
  FILE *hf_ = nullptr, *cmapf_ = nullptr;
  svi_fa2 *dev_ = nullptr;
};

#endif
