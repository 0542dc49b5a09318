#!/usr/bin/env python3
"""Development probe: time the phi sweep alone (write_comm on/off) at a workload.  Not part of the product."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, fast_state
from svinet_b200 import synth
from svinet_b200.engine import LinkSamplingEngine

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n, k, target = WORKLOADS[wl]
links = synth.mmsb_links(n, k, target, seed=1234, device="cuda:0")
g0, l0 = fast_state(n, k, links)
stream = torch.cuda.current_stream()
eng = LinkSamplingEngine(n, k, links, device=0, stream=stream.cuda_stream)
eng.set_state(g0, l0)
for it in range(2):
    eng.step(it, True, True)
out = {}
for comm in (0, 1):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record(stream)
    for r in range(reps):
        eng.phase_phi(2, comm); ev[r + 1].record(stream)
    torch.cuda.synchronize()
    ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
    out["phi_comm%d_ms" % comm] = (float(np.median(ms)), float(np.min(ms)))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record(stream)
for r in range(reps):
    eng.phase_s3(); ev[r + 1].record(stream)
torch.cuda.synchronize()
ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
out["s3_ms"] = (float(np.median(ms)), float(np.min(ms)))
for name, fn in (("node", eng.phase_node), ("finish", lambda: eng.phase_finish(True)), ("refresh_only", lambda: eng.phase_refresh(True))):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record(stream)
    for r in range(reps):
        fn(); ev[r + 1].record(stream)
    torch.cuda.synchronize()
    ms = [ev[r].elapsed_time(ev[r + 1]) for r in range(reps)]
    out[name + "_ms"] = (float(np.median(ms)), float(np.min(ms)))
# the same kernels inside a whole iteration (after the two sweeps, under the power cap)
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(reps)]
for r in range(reps):
    ev[r][0].record(stream); eng.phase_phi(5 + r, 1); ev[r][1].record(stream); eng.phase_node(); ev[r][2].record(stream)
    eng.phase_s3(); ev[r][3].record(stream); eng.phase_finish(True); ev[r][4].record(stream)
torch.cuda.synchronize()
out["in_step_ms"] = [round(float(np.median([ev[r][i].elapsed_time(ev[r][i + 1]) for r in range(reps)])), 3) for i in range(4)]
print(os.environ.get("SVI_LS_LIB", "default"), eng.info().get("lanes"), eng.info().get("vec"), out, flush=True)
