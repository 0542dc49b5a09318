#!/usr/bin/env python3
"""INTEGRATION.md option B, executed: the reference's own LinkSampling with its loop body replaced by libsvi_ls.so.

TEST INFRASTRUCTURE.  Reads the reference sources where they lie (/root/reference/src, never copied into the
repository), writes PATCHED COPIES of linksampling.{hh,cc} plus symlinks to the other, unmodified files into
oracle/_ref/b200_src/ (a build intermediate: `make -C oracle ref_b200` compiles it against include/svi_ls.h, links
svinet_b200/lib/libsvi_ls.so -> oracle/_ref/svinet_ref_b200, and removes the directory again).

The edits are anchored on identifiers (regular expressions), not on a diff, and are exactly INTEGRATION.md B.1-B.4:
  B.1  linksampling.hh   #include "svi_ls.h", one member + four private helpers
  B.2  infer()           dev_create() right after assign_training_links()
  B.3  infer()           everything from clear() to prune() becomes dev_step(write_comm)
  B.4  validation_likelihood / do_on_stop   dev_sync_state() first (the host mirrors _gamma/_lambda are refreshed, the
       reference's own edge_likelihood / save_model / write_groups then run unchanged)
       log_communities    dev_fill_communities() first (_communities rebuilt from the membership bits)
"""
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src"
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "b200_src")

HELPERS = r'''
// ---- added by oracle/ref_b200.py (INTEGRATION.md option B) ----------------------------------------------------
#include <vector>
#include <string.h>
void
LinkSampling::dev_create()
{
  std::vector<uint32_t> links(2 * (size_t)_nlinks);
  const double **ld = _links.const_data();
  for (uint32_t e = 0; e < _nlinks; ++e) { links[2*e] = (uint32_t)ld[e][0]; links[2*e+1] = (uint32_t)ld[e][1]; }
  std::vector<double> tl(_n);
  for (uint32_t i = 0; i < _n; ++i) tl[i] = _training_links[i];
  svi_ls_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.n = _n;  cfg.k = _k;  cfg.nlinks = _nlinks;
  cfg.alpha = _env.alpha;  cfg.eta0 = _env.eta0;  cfg.eta1 = _env.eta1;
  cfg.ones = _network.ones();  cfg.device = -1;  cfg.node_begin = 0;  cfg.node_end = _n;
  if (svi_ls_create(&cfg, &links[0], &tl[0], &_dev) != SVI_OK) {
    lerr("svi_ls_create: %s", svi_ls_last_error());
    exit(-1);
  }
  std::vector<double> g((size_t)_n * _k), l((size_t)_k * 2);
  for (uint32_t i = 0; i < _n; ++i) memcpy(&g[(size_t)i * _k], _gamma.const_data()[i], _k * sizeof(double));
  for (uint32_t k = 0; k < _k; ++k) memcpy(&l[2 * k], _lambda.const_data()[k], 2 * sizeof(double));
  if (svi_ls_set_state(_dev, &g[0], &l[0]) != SVI_OK) { lerr("%s", svi_ls_last_error()); exit(-1); }
}

void
LinkSampling::dev_step(bool write_comm)
{
  if (svi_ls_step(_dev, _iter, _annealing_phase, write_comm) != SVI_OK) { lerr("%s", svi_ls_last_error()); exit(-1); }
}

void
LinkSampling::dev_sync_state()
{
  if (!_dev) return;          // the constructor's first validation_likelihood runs before the device exists
  std::vector<double> g((size_t)_n * _k), l((size_t)_k * 2);
  if (svi_ls_get_state(_dev, &g[0], &l[0]) != SVI_OK) { lerr("%s", svi_ls_last_error()); exit(-1); }
  double **gd = _gamma.data(), **ld = _lambda.data();
  for (uint32_t i = 0; i < _n; ++i) memcpy(gd[i], &g[(size_t)i * _k], _k * sizeof(double));
  for (uint32_t k = 0; k < _k; ++k) memcpy(ld[k], &l[2 * k], 2 * sizeof(double));
}

void
LinkSampling::dev_fill_communities()
{
  if (!_dev) return;
  const uint32_t words = (_k + 31) / 32;
  std::vector<uint32_t> bits((size_t)_n * words);
  if (svi_ls_get_membership(_dev, &bits[0]) != SVI_OK) { lerr("%s", svi_ls_last_error()); exit(-1); }
  _communities.clear();
  for (uint32_t p = 0; p < _n; ++p)
    for (uint32_t c = 0; c < _k; ++c)
      if ((bits[(size_t)p * words + c / 32] >> (c % 32)) & 1u) _communities[c].push_back(p);
}
'''


def insert_after(lines, pattern, text, start=0, count=1):
    rx = re.compile(pattern)
    done = 0
    i = start
    while i < len(lines):
        if rx.search(lines[i]):
            lines[i + 1:i + 1] = text
            done += 1
            if done == count:
                return i + 1 + len(text)
            i += len(text)
        i += 1
    raise SystemExit("ref_b200.py: anchor %r not found" % pattern)


def first_brace_after(lines, pattern):
    rx = re.compile(pattern)
    for i, ln in enumerate(lines):
        if rx.search(ln):
            for j in range(i, i + 4):
                if lines[j].strip() == "{" or lines[j].rstrip().endswith("{"):
                    return j
    raise SystemExit("ref_b200.py: function %r not found" % pattern)


def main():
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        os.unlink(os.path.join(OUT, f))
    for f in sorted(os.listdir(REF)):
        if f.endswith((".cc", ".hh", ".h")) and f not in ("linksampling.cc", "linksampling.hh"):
            os.symlink(os.path.join(REF, f), os.path.join(OUT, f))

    hh = open(os.path.join(REF, "linksampling.hh")).read().split("\n")
    for i, ln in enumerate(hh):                                   # B.1
        if ln.startswith("#include"):
            hh.insert(i, '#include "svi_ls.h"')
            break
    insert_after(hh, r"\bbool\s+_annealing_phase\s*;", [
        "  svi_ls *_dev = NULL;            // device-side problem (oracle/ref_b200.py)",
        "  void dev_create();", "  void dev_step(bool write_comm);", "  void dev_sync_state();",
        "  void dev_fill_communities();"])
    open(os.path.join(OUT, "linksampling.hh"), "w").write("\n".join(hh))

    cc = open(os.path.join(REF, "linksampling.cc")).read().split("\n")
    infer = next(i for i, ln in enumerate(cc) if re.match(r"LinkSampling::infer\(\)", ln))
    at = insert_after(cc, r"^\s*assign_training_links\(\);\s*$", ["  dev_create();"], start=infer)          # B.2
    a = next(i for i in range(at, len(cc)) if re.match(r"^\s*clear\(\);\s*$", cc[i]))                        # B.3
    b = next(i for i in range(a, len(cc)) if re.match(r"^\s*prune\(\);\s*$", cc[i]))
    cc[a:b + 1] = ["    dev_step(write_comm);"]
    for fn in (r"^LinkSampling::validation_likelihood\(", r"^LinkSampling::do_on_stop\(\)"):                 # B.4
        j = first_brace_after(cc, fn)
        cc.insert(j + 1, "  dev_sync_state();")
    j = first_brace_after(cc, r"^LinkSampling::log_communities\(\)")
    cc.insert(j + 1, "  dev_fill_communities();")
    open(os.path.join(OUT, "linksampling.cc"), "w").write("\n".join(cc) + HELPERS)
    print("ref_b200.py: patched sources in", OUT)


if __name__ == "__main__":
    main()
