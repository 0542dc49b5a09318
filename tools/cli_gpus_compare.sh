#!/bin/bash
# development: `svinet -link-sampling -gpus G` against `-gpus 1` on a synthetic graph written to disk: same output
# directory up to the last printed digit?   usage: tools/cli_gpus_compare.sh <workload> <iterations> <G>
set -e
cd "$(dirname "$0")/.."
REPO=$PWD
WL=${1:-c3}; IT=${2:-8}; G=${3:-2}
D=$(mktemp -d)
python - "$WL" "$D" <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from bench import WORKLOADS
from svinet_b200 import synth
n, k, target = WORKLOADS[sys.argv[1]]
links = synth.mmsb_links(n, k, target, seed=1234, device="cuda:0")
used = np.unique(links)
remap = np.zeros(n, dtype=np.int64); remap[used] = np.arange(used.size)
import pandas as pd
pd.DataFrame(remap[links.astype(np.int64)]).to_csv(sys.argv[2] + "/g.txt", sep="\t", header=False, index=False)
open(sys.argv[2] + "/nk", "w").write("%d %d\n" % (used.size, k))
PY
read N K < $D/nk
echo "graph: n=$N k=$K $(wc -l < $D/g.txt) lines"
cd $D
for g in 1 $G; do
  mkdir run$g; ln -s ../g.txt run$g/g.txt
  ( cd run$g; T0=$(date +%s.%N)
    SVINET_TIMING=1 $REPO/svinet_b200/lib/svinet -file g.txt -n $N -k $K -link-sampling -max-iterations $IT -no-stop -gpus $g > out.log 2> err.log || { tail -5 err.log; exit 1; }
    echo "gpus=$g wall $(python -c "print(round($(date +%s.%N) - $T0, 2))") s"; grep -E "iterations:" err.log | head -2 )
done
python - "$D" "$N" "$K" <<'PY'
import sys, os, glob
sys.path.insert(0, os.path.join(os.environ.get("REPO", "."), "tests"))
d = sys.argv[1]
a = glob.glob(d + "/run1/n*-linksampling")[0]
b = [p for p in glob.glob(d + "/run*/n*-linksampling") if not p.startswith(d + "/run1/")][0]
from decimal import Decimal
tot = off = 0
for f in ("gamma.txt", "lambda.txt", "groups.txt", "validation.txt", "max.txt"):
    la, lb = open(os.path.join(a, f)).read().split("\n"), open(os.path.join(b, f)).read().split("\n")
    assert len(la) == len(lb), f
    for x, y in zip(la, lb):
        fx, fy = x.split(), y.split()
        assert len(fx) == len(fy), (f, x[:80], y[:80])
        for i, (u, v) in enumerate(zip(fx, fy)):
            tot += 1
            if u == v or (f in ("validation.txt", "max.txt") and i == 1):
                continue
            dec = len(v.split(".")[1]) if "." in v else 0
            assert "." in v and abs(Decimal(u) - Decimal(v)).scaleb(dec) <= 1, (f, u, v)
            off += 1
same = open(os.path.join(a, "communities.txt")).read() == open(os.path.join(b, "communities.txt")).read()
print("fields compared: %d, last-digit differences: %d, communities.txt identical: %s" % (tot, off, same))
assert same
PY
rm -rf $D
