#!/usr/bin/env python3
"""Development: a small workload that drives every ring-sweep instantiation the benchmark uses (K = 200: G8 x V13;
K = 100: G4 x V13 phi + G8 x V7 s3; K = 256: G16), with long segments, a hub and 35 % converged nodes, for
compute-sanitizer (memcheck / racecheck / synccheck).  Not part of the product.

    compute-sanitizer --tool racecheck --num-cuda-barriers 64 python tools/sanitize_ring.py [K ...]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from svinet_b200 import synth
from svinet_b200.engine import LinkSamplingEngine

ks = [int(a) for a in sys.argv[1:]] or [200, 100, 256]
for k in ks:
    n = 600
    base = synth.mmsb_links(n, k, n * 60, seed=k)
    hub = np.stack([np.zeros(n - 1, dtype=np.uint32), np.arange(1, n, dtype=np.uint32)], 1)
    links = np.unique(np.concatenate([base, hub]), axis=0).astype(np.uint32)
    rng = np.random.default_rng(k)
    deg = np.bincount(links.ravel().astype(np.int64), minlength=n).astype(np.float64)
    gamma = (deg[:, None] / k) * (1.0 + 0.5 * rng.random((n, k))) + 1.0 / k
    conv = np.zeros(n, dtype=np.uint32)
    who = rng.random(n) < 0.35
    conv[who] = rng.integers(1, k + 1, who.sum())
    eng = LinkSamplingEngine(n, k, links, seg_len=64)
    eng.set_state(gamma, np.ones((k, 2)))
    eng.set_converged(conv)
    for it in range(2):
        eng.step(it, True, True)
    g, lam = eng.get_state()
    print("k=%d ring_depth=%d segs=%d gamma checksum %.12e" % (k, eng.info()["ring_depth"], eng.info()["segments_phi"], g.sum()), flush=True)
    eng.close()
