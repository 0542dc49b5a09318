// textio.hh -- reading a saved model back (gamma.txt / lambda.txt, Appendix C of SURVEY.md): whitespace separated
// numeric rows, the first `skip` columns are ids.  The reference parses them line by line with strtod
// (src/linksampling.cc:1267-1352, src/fastamm2.cc:1717-1803); gamma.txt is 1.8 GB at n=1e6, k=200, so the file is
// mapped and its lines are parsed by all host threads.
#ifndef SVINET_B200_TEXTIO_HH
#define SVINET_B200_TEXTIO_HH

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

// Fills out[rows*cols] from the file; every line must hold at least skip + cols numbers (extra ones are
// ignored) and the file must hold exactly `rows` non-empty lines.  Returns 0, or -1 (missing file), -2 (a short
// line), -3 (wrong number of rows); `what` names the file in messages.
inline int load_numeric_rows(const std::string &path, uint32_t rows, uint32_t cols, uint32_t skip, double *out,
                             const char *what) {
  const int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) {
    fprintf(stderr, "no %s found\n", what);
    return -1;
  }
  struct stat st;
  fstat(fd, &st);
  const size_t len = (size_t)st.st_size;
  const char *data = len ? (const char *)mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0) : "";
  if (len && data == MAP_FAILED) {
    close(fd);
    fprintf(stderr, "cannot map %s\n", what);
    return -1;
  }
  // line starts (sequential newline scan is memchr-fast)
  std::vector<size_t> start;
  start.reserve((size_t)rows + 2);
  for (size_t i = 0; i < len;) {
    const char *nl = (const char *)memchr(data + i, '\n', len - i);
    const size_t e = nl ? (size_t)(nl - data) : len;
    bool blank = true;
    for (size_t j = i; j < e && blank; ++j) blank = data[j] == ' ' || data[j] == '\t' || data[j] == '\r';
    if (!blank) start.push_back(i);
    i = e + 1;
  }
  int rc = 0;
  if (start.size() != rows) {
    fprintf(stderr, "%s has %zu rows, expected %u\n", what, start.size(), rows);
    rc = -3;
  } else {
    const unsigned nt = std::max(1u, std::min(16u, std::min<unsigned>(std::thread::hardware_concurrency(), rows / 1024 + 1)));
    std::vector<int> bad(nt, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t] {
        std::string line;
        for (uint32_t r = (uint32_t)((uint64_t)rows * t / nt); r < (uint32_t)((uint64_t)rows * (t + 1) / nt); ++r) {
          const char *b = data + start[r];
          const char *nl = (const char *)memchr(b, '\n', len - start[r]);
          line.assign(b, nl ? (size_t)(nl - b) : len - start[r]);     // strtod needs a terminated buffer
          char *p = &line[0];
          uint32_t col = 0;
          while (col < skip + cols) {
            char *q = nullptr;
            const double d = strtod(p, &q);
            if (q == p) break;
            p = q;
            if (col >= skip) out[(size_t)r * cols + (col - skip)] = d;
            col++;
          }
          if (col < skip + cols) bad[t] = 1;
        }
      });
    for (auto &x : th) x.join();
    for (int b : bad)
      if (b) {
        fprintf(stderr, "error parsing %s\n", what);
        rc = -2;
      }
  }
  if (len) munmap((void *)data, len);
  close(fd);
  return rc;
}

#endif
