#!/bin/bash
# development: host start-up of the CLI (-dump-init, no device work) at a workload, one producer against the default
set -e
cd "$(dirname "$0")/.."
REPO=$PWD
WL=${1:-c4}
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
cd $D
nproc
for p in 1 default; do
  mkdir -p dump_$p
  if [ "$p" = "1" ]; then export SVINET_INIT_PRODUCERS=1; else unset SVINET_INIT_PRODUCERS; fi
  echo "producers=$p"
  SVINET_TIMING=1 $REPO/svinet_b200/lib/svinet -file g.txt -n $N -k $K -link-sampling -max-iterations 2 -no-stop -dump-init dump_$p 2>&1 | grep -E "init gamma|held-out"
  md5sum dump_$p/gamma.f64 | cut -c1-16
  rm -rf dump_$p
done
rm -rf $D
