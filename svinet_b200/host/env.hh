// env.hh -- run configuration of the svinet drop-in CLI (-link-sampling and -rnode -stratified).
//
// Mirrors what the reference's Env carries for this path (reference src/env.hh:52-202,285-628 and the
// argv loop src/main.cc:114-242): same flags, same defaults, same output-directory naming, same
// param.txt lines.  Written from scratch around a plain aggregate + a parser; the reference's
// 52-argument constructor is not reproduced.
#ifndef SVINET_B200_ENV_HH
#define SVINET_B200_ENV_HH

#include <cinttypes>
#include <cstdint>
#include <cstdio>
#include <string>

struct Env {
  // ---- command line (src/main.cc:48-110 defaults) ----
  uint32_t n = 0, k = 0;
  std::string datfname = "network.dat";
  std::string label = "mmsb";
  bool link_sampling = false;
  bool batch = false, stratified = false, rnode = false, rpair = false, massive = false, single = false;
  bool gen = false, ppc = false, gml = false, findk = false, lcstats = false, orig = false, run_gap = false;
  bool model_load = false;            // -load <dir/>
  std::string gamma_location;
  bool load_heldout = false;          // -load-validation <file>
  std::string load_heldout_fname;
  bool load_test = false;             // -load-test <file>
  std::string load_test_fname;
  bool load_test_sets = false;
  double heldout_ratio = 0.01;        // -heldout-ratio
  std::string eta_type = "uniform";   // -eta-type
  bool nmi = false;
  std::string ground_truth_fname;
  int reportfreq = 1;                 // -rfreq; -link-sampling resets it to 1 (src/main.cc:149-153)
  bool accuracy = false;
  double stopthresh = 0.00001, infthresh = 0;
  bool nonuniform = false, benchmark = false, randzeros = false, preprocess = false, strid = false;
  std::string groups_file;
  bool logl = false;
  uint32_t max_iterations = 0;
  bool use_validation_stop = true;    // cleared by -no-stop
  double seed = 0;
  double link_thresh_arg = 0.5;       // logged only: the reference never stores it
  uint32_t lt_min_deg_arg = 0;
  bool use_init_communities = false;
  std::string init_communities_fname;
  uint32_t nthreads = 0, itype = 0, scale = 1;
  bool nodelay = true, disjoint = false, adamic_adar = false;
  int ngpus = 1;                      // extension: -gpus N (ignored by the reference's parser)
  bool device_draw = false;           // extension: -device-draw (with -rnode -stratified): the device draws the
                                      //            minibatches from a Philox stream keyed by -seed (svi_fa2_run)
  bool dump_only = false;             // extension: -dump-init <dir> writes the start-up state and exits
  std::string dump_dir;               //            (host-logic tests; touches no GPU)

  // ---- constants of the reference's initialiser list (src/env.hh:305-483) ----
  uint32_t t = 2;
  double alpha = 0;                   // 1/k (src/env.hh:344)
  double eta0 = 0, eta1 = 0;          // set by Network::set_env_variables (src/network.cc:223-250)
  double eta0_dense = 4700.59, eta1_dense = 0.77, eta0_sparse = 0.97, eta1_sparse = 6.33;
  double epsilon = 1e-30;             // src/env.hh:395
  double meanchangethresh = 0.00001;  // src/env.hh:337  (the -rnode -stratified path, src/fastamm2.hh:194)
  double tau0 = 1024, nodetau0 = 1024, nodekappa = 0.5, kappa = 0.9;   // src/env.hh:405-408
  uint32_t online_iterations = 50;    // src/env.hh:415
  bool deterministic = false;         // src/env.hh:446
  double precision_ratio = 0.001;
  bool undirected = true, nolambda = false;
  // effective values on this path (both members are never assigned in the reference)
  double link_thresh = 0.0, lt_min_deg = 0.0;

  // ---- derived at run time ----
  uint64_t total_pairs = 0;           // n*(n-1)/2 evaluated in 32 bits (src/network.cc:225)
  double ones_prob = 0, zeros_prob = 0;
  volatile bool terminate = false;    // set by the SIGTERM handler (src/main.cc:29-40)
  std::string prefix;                 // output directory

  // parse argv exactly like src/main.cc:114-242 (unknown flags are ignored); returns false on -help
  bool parse(int argc, char **argv);
  // directory name (src/env.hh:503-568), mkdir, infer.log, param.txt head, network.dat symlink
  void open_output();
  std::string file(const std::string &name) const { return prefix + name; }

  // param.txt writers (formats of Env::plog, src/env.hh:205-259)
  void plog(const std::string &key, double v) const;
  void plog(const std::string &key, bool v) const;
  void plog(const std::string &key, int v) const;
  void plog(const std::string &key, uint32_t v) const;
  void plog(const std::string &key, uint64_t v) const;
  void plog(const std::string &key, const std::string &v) const;
  void plog(const std::string &key, const char *v) const { plog(key, std::string(v)); }

  static void usage();

 private:
  FILE *plogf_ = nullptr;
};

#endif
