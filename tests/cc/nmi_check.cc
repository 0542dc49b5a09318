// prints nmi::lfk (svinet_b200/host/nmi.hh) of two cover files (one community per line, node ids 0..n-1)
#include "nmi.hh"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>

static std::vector<std::vector<uint32_t>> read_cover(const char *path) {
  std::vector<std::vector<uint32_t>> c;
  std::ifstream f(path);
  std::string line;
  while (std::getline(f, line)) {
    std::istringstream ss(line);
    std::vector<uint32_t> v;
    uint32_t x;
    while (ss >> x) v.push_back(x);
    c.push_back(v);
  }
  return c;
}

int main(int argc, char **argv) {
  if (argc != 4) return 2;
  printf("%.15g\n", nmi::lfk((uint32_t)atoi(argv[1]), read_cover(argv[2]), read_cover(argv[3])));
  return 0;
}
