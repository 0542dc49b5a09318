#!/usr/bin/env python3
"""Generate tests/golden/* by running the UNMODIFIED reference (oracle/_ref/svinet_ref).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs oracle/_ref, i.e.
`make -C oracle ref`, which needs /root/reference):

    python oracle/make_golden.py

Each case runs `svinet_ref -file G -n N -k K -link-sampling <flags>` (or `-rnode -stratified <flags>` for
the fa2_* cases, class FastAMM2) in a scratch
directory (the reference writes its output directory into the cwd, env.hh:503-568) and
copies the files that pin the path's results into tests/golden/<case>/.  The wall-clock
"secs" column (col 2 of validation.txt / max.txt, linksampling.cc:996-1002,1030-1034) is
zeroed so the fixtures are reproducible.  Large gamma.txt files are stored gzip'ed.
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_BIN = os.path.join(HERE, "_ref", "svinet_ref")
DATA = os.path.join(HERE, "_ref", "data")
GOLD = os.path.join(REPO, "tests", "golden")

# name -> (input file, n, k, extra flags, output dir the reference creates)
CASES = {
    "c1_m30": ("assort-75-4.txt", 75, 4, ["-max-iterations", "30", "-no-stop"], "n75-k4-mmsb-linksampling"),
    "c1_natural": ("assort-75-4.txt", 75, 4, [], "n75-k4-mmsb-linksampling"),
    "c1_m1": ("assort-75-4.txt", 75, 4, ["-max-iterations", "1", "-no-stop"], "n75-k4-mmsb-linksampling"),
    "c1_seed7_m12": ("assort-75-4.txt", 75, 4, ["-max-iterations", "12", "-no-stop", "-seed", "7"],
                     "n75-k4-mmsb-seed7-linksampling"),
    "c1_accuracy_m8": ("assort-75-4.txt", 75, 4, ["-max-iterations", "8", "-no-stop", "-accuracy"],
                       "n75-k4-mmsb-linksampling"),
    "c1_k7_m15": ("assort-75-4.txt", 75, 7, ["-max-iterations", "15", "-no-stop"], "n75-k7-mmsb-linksampling"),
    "lfr_k28_m20": ("LFR-network-n1000-k28.txt", 1000, 28, ["-max-iterations", "20", "-no-stop"],
                    "n1000-k28-mmsb-linksampling"),
    "c2_m12": ("ca-AstroPh.csv", 17903, 20, ["-max-iterations", "12", "-no-stop"], "n17903-k20-mmsb-linksampling"),
    "c2_m25": ("ca-AstroPh.csv", 17903, 20, ["-max-iterations", "25", "-no-stop"], "n17903-k20-mmsb-linksampling"),
    "c1_etasparse_m10": ("assort-75-4.txt", 75, 4, ["-max-iterations", "10", "-no-stop", "-eta-type", "sparse"],
                         "n75-k4-mmsb-linksampling"),
    # -init-communities (init_gamma_external, linksampling.cc:404-452): gamma starts from given communities, no RNG
    "c1_initcomm_m10": ("assort-75-4.txt", 75, 4, ["-max-iterations", "10", "-no-stop", "-init-communities",
                                                   "assort-75-4-init-communities.txt"], "n75-k4-mmsb-linksampling"),
    # the natural run: the validation stop ends it (iteration 30 here); small files only
    "c2_natural": ("ca-AstroPh.csv", 17903, 20, [], "n17903-k20-mmsb-linksampling"),
}
# extra input files a case needs in its cwd (committed under tests/golden/inputs/)
EXTRA_INPUTS = {"c1_initcomm_m10": ["assort-75-4-init-communities.txt"]}
KEEP_EXTRA = {"c1_initcomm_m10": ["init_memberships.txt"]}
KEEP_ONLY = {"c2_natural": ["lambda.txt", "communities.txt", "validation.txt", "max.txt", "validation-edges.txt", "param.txt"]}

# `-rnode -stratified` (class FastAMM2): name -> (input, n, k, flags, outdir)
FA2_CASES = {
    "fa2_c1_m200": ("assort-75-4.txt", 75, 4, ["-max-iterations", "200", "-rfreq", "50"], "n75-k4-mmsb-Srnode"),
    "fa2_c1_k6_seed9_m500": ("assort-75-4.txt", 75, 6, ["-max-iterations", "500", "-rfreq", "100", "-seed", "9"],
                             "n75-k6-mmsb-seed9-Srnode"),
    "fa2_lfr_k28_m300": ("LFR-network-n1000-k28.txt", 1000, 28, ["-max-iterations", "300", "-rfreq", "100"],
                         "n1000-k28-mmsb-Srnode"),
    "fa2_c2_m120": ("ca-AstroPh.csv", 17903, 20, ["-max-iterations", "120", "-rfreq", "40"],
                    "n17903-k20-mmsb-Srnode"),
}
FA2_KEEP = ["gamma.txt", "lambda.txt", "heldout.txt", "heldout-pairs.txt", "groups.txt", "communities.txt",
            "communities_size.txt", "summary.txt", "param.txt"]

KEEP = ["gamma.txt", "lambda.txt", "communities.txt", "groups.txt", "validation.txt", "max.txt",
        "validation-edges.txt", "param.txt"]
GZIP_OVER = 256 * 1024


def zero_secs_column(path):
    out = []
    with open(path) as f:
        for line in f:
            parts = line.rstrip("\n").split("\t")
            if len(parts) > 1:
                parts[1] = "0"
            out.append("\t".join(parts))
    with open(path, "w") as f:
        f.write("\n".join(out) + ("\n" if out else ""))


def main():
    if not os.path.exists(REF_BIN):
        sys.exit("oracle/_ref/svinet_ref missing: run `make -C oracle ref` first")
    only = set(sys.argv[1:])
    manifest_path = os.path.join(GOLD, "MANIFEST.json")
    manifest = json.load(open(manifest_path)) if os.path.exists(manifest_path) else {}
    jobs = [(name, c, "-link-sampling", KEEP) for name, c in CASES.items()]
    jobs += [(name, c, "-rnode -stratified", FA2_KEEP) for name, c in FA2_CASES.items()]
    for name, (fname, n, k, flags, outdir), mode, keep in jobs:
        if only and name not in only:
            continue
        scratch = tempfile.mkdtemp(prefix="golden_")
        shutil.copy(os.path.join(DATA, fname), os.path.join(scratch, fname))
        for extra in EXTRA_INPUTS.get(name, []):
            shutil.copy(os.path.join(GOLD, "inputs", extra), os.path.join(scratch, extra))
        cmd = [REF_BIN, "-file", fname, "-n", str(n), "-k", str(k)] + mode.split() + flags
        with open(os.path.join(scratch, "stdout.log"), "w") as log:
            rc = subprocess.call(cmd, cwd=scratch, stdout=log, stderr=subprocess.STDOUT)
        if rc != 0:
            sys.exit("reference failed (%d) for %s" % (rc, name))
        src = os.path.join(scratch, outdir)
        dst = os.path.join(GOLD, name)
        shutil.rmtree(dst, ignore_errors=True)
        os.makedirs(dst)
        entry = {"input": fname, "n": n, "k": k, "flags": flags, "outdir": outdir, "md5": {}, "mode": mode}
        if name in EXTRA_INPUTS:
            entry["extra_inputs"] = EXTRA_INPUTS[name]
        for f in KEEP_ONLY.get(name, keep) + KEEP_EXTRA.get(name, []):
            p = os.path.join(src, f)
            if not os.path.exists(p):
                continue
            if f in ("validation.txt", "max.txt", "heldout.txt"):
                zero_secs_column(p)
            data = open(p, "rb").read()
            entry["md5"][f] = hashlib.md5(data).hexdigest()
            if len(data) > GZIP_OVER:
                with gzip.GzipFile(os.path.join(dst, f + ".gz"), "wb", compresslevel=9, mtime=0) as g:
                    g.write(data)
            else:
                with open(os.path.join(dst, f), "wb") as o:
                    o.write(data)
        manifest[name] = entry
        shutil.rmtree(scratch, ignore_errors=True)
        print("golden %-16s ok  (%s)" % (name, " ".join(cmd[1:])))
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
