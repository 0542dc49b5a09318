#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_ring.py tests/test_gpu_mg.py -x -q 2>&1 | tail -3
for wl in c2s tiny; do
timeout 300 python bench.py --workload $wl --steps 100 --warmup 10 --no-cpu-baseline --no-fa2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl', round(d['ms_per_step'],4), d['phase_ms'], d['verify']['ok'], d['late_run']['ms_per_step'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_c2s.csv python bench.py --workload c2s --steps 2 --warmup 1 --no-cpu-baseline --no-verify --no-fa2 --converged-frac 0 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_c2s.csv')) if len(r)>5]
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[start+1:]:
    if 'svi::' in r[ki]: agg[r[ki][:60]].append(float(r[vi].replace(',','')))
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print("%-62s n=%3d mean=%9.1f ns" % (k, len(v), sum(v)/len(v)))
PY
