// main.cc -- `svinet` command line, B200 build: the reference's CLI surface (src/main.cc:43-377) for the
// two engines this repository re-implements, `-link-sampling` and `-rnode -stratified`.  Every other mode is out
// of scope and is refused with a message instead of being silently ignored.
#include <csignal>
#include <cstdio>
#include <cstdlib>

#include "env.hh"
#include "fastamm2.hh"
#include "linksampling.hh"
#include "network.hh"

static Env *env_global = nullptr;

static void term_handler(int sig) {
  if (env_global) {
    // same contract as src/main.cc:29-40: flag the run, the loop dumps the model at the next iteration
    env_global->terminate = true;
  } else {
    signal(sig, SIG_DFL);
    raise(sig);
  }
}

int main(int argc, char **argv) {
  signal(SIGTERM, term_handler);
  // -gpus N: a shard runs four streams with flag-wait kernels at their heads; give CUDA enough hardware queues that
  // no two of them alias (read when CUDA initialises; a value set by the user is kept)
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  if (argc == 1) {
    Env::usage();
    exit(-1);
  }
  Env env;
  if (!env.parse(argc, argv)) {
    Env::usage();
    exit(0);
  }
  const bool fa2 = !env.link_sampling && env.stratified && env.rnode;       // src/main.cc:368
  if ((!env.link_sampling && !fa2) || env.batch || env.massive || env.single || env.gen || env.ppc || env.gml ||
      env.findk || env.lcstats || env.orig) {
    fprintf(stderr,
            "svinet (B200 build): -link-sampling and -rnode -stratified are implemented here; the other engines and\n"
            "tools of the reference (-batch, -rpair, -infset, -single, -orig, -gen, -ppc, -gml, -findk) are unchanged\n"
            "upstream code and are not part of this build.\n");
    exit(-1);
  }
  if (env.n == 0 || env.k == 0) {
    fprintf(stderr, "svinet: -n and -k are required\n");
    exit(-1);
  }
  env.open_output();
  env_global = &env;

  Network network(env);
  if (network.read(env.datfname) < 0) {
    fprintf(stderr, "error reading %s; quitting\n", env.datfname.c_str());
    return -1;
  }
  env.n = network.n() - network.singles();   // src/main.cc:291

  if (fa2) {
    FastAMM2 fastamm2(env, network);
    if (env.dump_only) {
      fastamm2.dump_init(env.dump_dir);
      exit(0);
    }
    fastamm2.infer();
    exit(0);
  }
  LinkSampling ls(env, network);
  if (env.dump_only) {
    ls.dump_init(env.dump_dir);
    exit(0);
  }
  ls.infer();
  exit(0);
}
