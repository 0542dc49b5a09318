// env.cc -- argv parsing, output directory, param.txt (see env.hh).
#include "env.hh"

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <sys/stat.h>
#include <unistd.h>

namespace {
const char *next_arg(int argc, char **argv, int &i) {
  if (i + 1 > argc - 1) {
    fprintf(stderr, "+ insufficient arguments!\n");
    exit(-1);
  }
  return argv[++i];
}
}  // namespace

bool Env::parse(int argc, char **argv) {
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-help") return false;
    else if (a == "-file") datfname = next_arg(argc, argv, i);
    else if (a == "-n") n = (uint32_t)atoi(next_arg(argc, argv, i));
    else if (a == "-k") k = (uint32_t)atoi(next_arg(argc, argv, i));
    else if (a == "-label") label = next_arg(argc, argv, i);
    else if (a == "-link-sampling") { link_sampling = true; batch = false; reportfreq = 1; }
    else if (a == "-batch") { batch = true; reportfreq = 1; }
    else if (a == "-stratified") { stratified = true; if (reportfreq == 1) reportfreq = 100; }
    else if (a == "-rnode") { rnode = true; if (reportfreq == 1) reportfreq = 100; }
    else if (a == "-rpair") { rpair = true; if (reportfreq == 1) reportfreq = 100; }
    else if (a == "-load") { model_load = true; gamma_location = next_arg(argc, argv, i); }
    else if (a == "-load-validation") { load_heldout = true; load_heldout_fname = next_arg(argc, argv, i); }
    else if (a == "-load-test") { load_test = true; load_test_fname = next_arg(argc, argv, i); }
    else if (a == "-load-test-sets") load_test_sets = true;
    else if (a == "-heldout-ratio") heldout_ratio = atof(next_arg(argc, argv, i));
    else if (a == "-eta-type") eta_type = next_arg(argc, argv, i);
    else if (a == "-nmi") { ground_truth_fname = next_arg(argc, argv, i); nmi = true; }
    else if (a == "-rfreq") reportfreq = atoi(next_arg(argc, argv, i));
    else if (a == "-accuracy") accuracy = true;
    else if (a == "-stopthresh") stopthresh = atof(next_arg(argc, argv, i));
    else if (a == "-inf") infthresh = atof(next_arg(argc, argv, i));
    else if (a == "-nonuniform") nonuniform = true;
    else if (a == "-bmark") benchmark = true;
    else if (a == "-randzeros") randzeros = true;
    else if (a == "-preprocess") { preprocess = true; massive = true; }
    else if (a == "-strid") strid = true;
    else if (a == "-groups-file") groups_file = next_arg(argc, argv, i);
    else if (a == "-logl") logl = true;
    else if (a == "-max-iterations") max_iterations = (uint32_t)atoi(next_arg(argc, argv, i));
    else if (a == "-no-stop") use_validation_stop = false;
    else if (a == "-seed") seed = atof(next_arg(argc, argv, i));
    else if (a == "-link-thresh") link_thresh_arg = atof(next_arg(argc, argv, i));
    else if (a == "-lt-min-deg") lt_min_deg_arg = (uint32_t)atof(next_arg(argc, argv, i));
    else if (a == "-init-communities") { use_init_communities = true; init_communities_fname = next_arg(argc, argv, i); }
    else if (a == "-nthreads") nthreads = (uint32_t)atoi(next_arg(argc, argv, i));
    else if (a == "-itype") itype = (uint32_t)atoi(next_arg(argc, argv, i));
    else if (a == "-scale") scale = (uint32_t)atoi(next_arg(argc, argv, i));
    else if (a == "-infset") massive = true;
    else if (a == "-single") single = true;
    else if (a == "-orig") orig = true;
    else if (a == "-gen") gen = true;
    else if (a == "-ppc") ppc = true;
    else if (a == "-gml") gml = true;
    else if (a == "-findk") findk = true;
    else if (a == "-lcstats") lcstats = true;
    else if (a == "-gp") run_gap = true;
    else if (a == "-disjoint") disjoint = true;
    else if (a == "-adamic-adar") adamic_adar = true;
    else if (a == "-gpus") ngpus = atoi(next_arg(argc, argv, i));
    else if (a == "-device-draw") device_draw = true;
    else if (a == "-dump-init") { dump_only = true; dump_dir = next_arg(argc, argv, i); }
    // -force, -online, -nodelay and anything unknown: accepted and ignored, like the reference
  }
  alpha = k ? 1.0 / k : 0.0;
  return true;
}

void Env::open_output() {
  // n<N>-k<K>-<label>[-seed<S>]-linksampling | -S..rnode   (src/env.hh:503-545)
  std::ostringstream sa;
  sa << "n" << n << "-" << "k" << k;
  if (label != "") {
    sa << "-" << label;
  } else if (datfname.length() > 3 && datfname.find("mmsb_gen.dat") == std::string::npos) {
    std::string q = datfname.substr(0, 2);
    if (q == "..") q = "xx";
    sa << "-" << q;
  }
  if (seed) sa << "-seed" << seed;
  if (link_sampling) {
    sa << "-linksampling";
  } else {
    // stratified || delaylearn || nolambda || undirected || randomnode: undirected is always true
    sa << "-";
    if (stratified) sa << "S";
    if (nolambda) sa << "P";
    if (rpair) sa << "rpair";
    if (rnode) sa << "rnode";
    if (nonuniform) sa << "R";
  }
  if (run_gap) sa << "-GAP";
  if (nthreads > 0) sa << "-T" << nthreads;
  if (itype > 0) sa << "-i" << itype;
  prefix = sa.str();

  fprintf(stdout, "+ Output directory: %s\n", prefix.c_str());
  fflush(stdout);
  struct stat st;
  if (stat(prefix.c_str(), &st) != 0) {
    if (errno != ENOENT || mkdir(prefix.c_str(), S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH) != 0) {
      fprintf(stderr, "Warning: could not create dir %s\n", prefix.c_str());
      exit(-1);
    }
  }  // an existing directory is reused (force_overwrite_dir = true, src/main.cc:49)
  if (FILE *lf = fopen(file("/infer.log").c_str(), "w")) fclose(lf);

  plogf_ = fopen(file("/param.txt").c_str(), "w");
  if (!plogf_) {
    printf("cannot open param file:%s\n", strerror(errno));
    exit(-1);
  }
  // src/env.hh:582-618, same keys, same order
  plog("nodes", n);
  plog("groups", k);
  plog("t", t);
  plog("minibatch (rpair or stratified rpair options only)", n / 2);
  plog("mbsize", (uint32_t)1);
  plog("alpha", alpha);
  plog("sbm_alpha", alpha);
  plog("heldout_ratio", heldout_ratio);
  plog("precision_ratio", precision_ratio);
  plog("stratified", stratified);
  plog("delaylearn", !nodelay);
  plog("nolambda", nolambda);
  plog("randomnode", rnode);
  plog("gen", gen);
  plog("undirected", undirected);
  plog("gap", run_gap);
  plog("nthreads", nthreads);
  plog("stopthresh", stopthresh);
  plog("infthresh", infthresh);
  plog("randzeros", randzeros);
  plog("benchmark", benchmark);
  plog("max iterations", max_iterations);
  plog("seed", seed);
  plog("use validation stop", use_validation_stop);
  plog("gamma location", gamma_location);
  plog("link_thresh", link_thresh_arg);
  plog("lt_min_deg", lt_min_deg_arg);
  plog("epsilon", epsilon);
  plog("sets_mini_batch", n / 100);
  plog("use_init_communities", use_init_communities);
  plog("load_test_sets", load_test_sets);
  plog("val_load", load_heldout);
  plog("val_file_location", load_heldout_fname);
  plog("test_load", load_test);
  plog("test_file_location", load_test_fname);
  plog("reportfreq", reportfreq);
  plog("eta_type", eta_type);

  // network.dat -> the input path exactly as given (src/env.hh:621-625)
  const std::string link = file("/network.dat");
  unlink(link.c_str());
  if (symlink(datfname.c_str(), link.c_str()) < 0) {
    fprintf(stderr, "cannot create %s: %s\n", link.c_str(), strerror(errno));
    exit(-1);
  }
  unlink(file("/mutual.txt").c_str());
}

void Env::plog(const std::string &key, double v) const { fprintf(plogf_, "%s: %.9f\n", key.c_str(), v); fflush(plogf_); }
void Env::plog(const std::string &key, bool v) const { fprintf(plogf_, "%s: %s\n", key.c_str(), v ? "True" : "False"); fflush(plogf_); }
void Env::plog(const std::string &key, int v) const { fprintf(plogf_, "%s: %d\n", key.c_str(), v); fflush(plogf_); }
void Env::plog(const std::string &key, uint32_t v) const { fprintf(plogf_, "%s: %d\n", key.c_str(), v); fflush(plogf_); }
void Env::plog(const std::string &key, uint64_t v) const { fprintf(plogf_, "%s: %" PRIu64 "\n", key.c_str(), v); fflush(plogf_); }
void Env::plog(const std::string &key, const std::string &v) const { fprintf(plogf_, "%s: %s\n", key.c_str(), v.c_str()); fflush(plogf_); }

void Env::usage() {
  fprintf(stdout,
          "\nSVINET (B200 build): stochastic variational inference of undirected networks\n"
          "svinet [OPTIONS]\n"
          "\t-help\t\tusage\n\n"
          "\t-file <name>\tinput tab-separated file with a list of undirected links\n\n"
          "\t-n <N>\t\tnumber of nodes in network\n\n"
          "\t-k <K>\t\tnumber of communities\n\n"
          "\t-link-sampling\tinference using link sampling\n\n"
          "\t-rnode -stratified\tinference using stratified random-node sampling (class FastAMM2)\n\n"
          "\t-device-draw\t(with -rnode -stratified) minibatches are drawn on the GPU from a Philox stream\n\n"
          "\t-load-validation <fname>\tuse the pairs in the file as the validation set for convergence\n\n"
          "\t-load <dir/>\tresume from <dir/>gamma.txt and <dir/>lambda.txt\n\n"
          "\t-label\t\ttag output directory\n\n"
          "\t-rfreq\t\tset the frequency at which convergence is estimated and statistics are logged\n\n"
          "\t-max-iterations\tmaximum number of iterations (use with -no-stop to avoid stopping earlier)\n\n"
          "\t-no-stop\tdisable stopping criteria\n\n"
          "\t-seed\t\tset the random generator seed\n\n"
          "\t-heldout-ratio\tfraction of links held out for validation (default 0.01)\n\n"
          "\t-eta-type\tuniform | fromdata | sparse | dense\n\n"
          "\t-accuracy\tno held-out set: every link is a training link\n\n");
  fflush(stdout);
}
