#!/bin/bash
# development: the drop-in CLI end to end on a synthetic graph written to disk (ingest, init, iterations, writers)
# usage: tools/cli_at_scale.sh <workload c3|c2s|tiny> <iterations> [extra svinet flags]
set -e
cd "$(dirname "$0")/.."
WL=${1:-c3}; IT=${2:-10}; shift 2 || true
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
( time SVINET_TIMING=1 $OLDPWD/svinet_b200/lib/svinet -file g.txt -n $N -k $K -link-sampling -max-iterations $IT -no-stop "$@" > out.log 2> err.log ) 2>&1 | grep real
grep -E "^\[" err.log | head -40
ls -la n$N-k$K-*/ | head -20
rm -rf $D
