#!/usr/bin/env python3
"""Development probe: phi / s3 sweep times with a fraction of the nodes already converged (late-run regime)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, fast_state
from svinet_b200 import synth
from svinet_b200.engine import LinkSamplingEngine

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
n, k, target = WORKLOADS[wl]
links = synth.mmsb_links(n, k, target, seed=1234, device="cuda:0")
g0, l0 = fast_state(n, k, links)
stream = torch.cuda.current_stream()
eng = LinkSamplingEngine(n, k, links, device=0, stream=stream.cuda_stream)
rng = np.random.default_rng(0)
for frac in (0.0, 0.35, 0.8):
    eng.set_state(g0, l0)
    conv = np.zeros(n, dtype=np.uint32)
    who = rng.random(n) < frac
    conv[who] = rng.integers(1, k + 1, who.sum())
    eng.set_converged(conv)
    out = {}
    for name, fn in (("phi_comm", lambda: eng.phase_phi(2, 1)), ("s3", lambda: eng.phase_s3())):
        fn(); torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        ev[0].record(stream)
        for r in range(6):
            fn(); ev[r + 1].record(stream)
        torch.cuda.synchronize()
        out[name] = round(float(np.median([ev[r].elapsed_time(ev[r + 1]) for r in range(6)])), 2)
    print("converged fraction %.2f:" % frac, out, flush=True)
