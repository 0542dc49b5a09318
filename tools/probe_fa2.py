#!/usr/bin/env python3
"""Development probe: a few eager-mode FastAMM2 iterations at a workload (for ncu captures of k_fa2_pairs / k_fa2_blend)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS
from bench_fa2 import make_problem
from svinet_b200.fa2_engine import Fa2Engine

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
eager = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n, k, target = WORKLOADS[wl]
links, heldout, hy, shuffled, gamma, lam = make_problem(n, k, target, "cuda:0")
eng = Fa2Engine(n, k, device=0, eager_blend=eager)
eng.set_state(gamma, lam)
eng.set_graph(links, heldout, shuffled)
seed = 20261017
done = 0
for it in range(64):
    typ, start, pr = eng.draw(it, seed)
    if typ == 1:                      # non-informative sets only: n/10 pairs
        eng.step(it, typ, start, pr)
        done += 1
        if done == 4:
            break
eng.sync()
print("probe_fa2 done", eng.info())
