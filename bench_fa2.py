#!/usr/bin/env python3
"""bench_fa2.py -- secondary bench: the `-rnode -stratified` path (class FastAMM2) on synthetic MMSB graphs.

    python bench_fa2.py [--workload c3|c4|c2s|tiny] [--steps K] [--warmup W] [--no-cpu-baseline]

bench.py (the driver's contract) measures the headline `-link-sampling` path; this file measures the second
path of SURVEY.md section 8 (rows a9-a11, f4) the same way.  One "step" = one iteration of FastAMM2::infer
(src/fastamm2.cc:566-640): minibatch draw (device Philox stream, svi_fa2_run), per-pair two-phi coordinate
ascent, Robbins-Monro blend of ALL N gamma rows and of lambda.  Reported: iterations/s, pair-updates/s
(pairs x iterations / device time), mean coordinate-ascent rounds per pair.  The default handle carries the
decay of the untouched rows as one scalar (lazy mode); a second handle with eager_blend = 1 runs the
reference's explicit pass and gives `eager_blend_ms_per_step` and the roofline of that blend (the
bandwidth-bound kernel: it reads and writes every gamma row, 2 x N x ld x 8 bytes per iteration; timed as a
step with an empty minibatch, i.e. prep + an idle pair kernel + the blend + the lambda kernel).
`e2e` = the same iterations driven with HOST-chosen minibatches through svi_fa2_step (pair list uploaded
every iteration) plus a held-out evaluation and a state download every `--report` iterations, which is what
the reference-facing CLI does.  `cpu_baseline` = the oracle (oracle/oracle_fa2.c, 1 thread) on a bounded
sample of the workload, through bench.py's cpu_baseline leg (this file never touches oracle/ itself).
`python bench.py --path fa2 ...` forwards here.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
from bench import WORKLOADS, ClockSampler, cpu_baseline_fa2_port, fast_state, measured_peak_hbm   # noqa: E402

METRIC = "fa2_pair_updates_per_sec"
UNIT = "pair-updates/s"


def make_problem(n, k, target, device):
    from svinet_b200 import synth
    links = synth.mmsb_links(n, k, target, seed=1234, device=device)
    rng = np.random.default_rng(7)
    nh = max(2, links.shape[0] // 100)
    ho_links = links[rng.choice(links.shape[0], size=nh // 2, replace=False)]
    p = rng.integers(0, n, nh // 2).astype(np.uint32)
    q = ((p.astype(np.int64) + 1 + rng.integers(0, n - 1, nh // 2)) % n).astype(np.uint32)
    ho_non = np.stack([np.minimum(p, q), np.maximum(p, q)], axis=1).astype(np.uint32)
    heldout = np.concatenate([ho_links, ho_non])
    hy = np.concatenate([np.ones(len(ho_links), np.uint8), np.zeros(len(ho_non), np.uint8)])
    shuffled = rng.permutation(n).astype(np.uint32)
    gamma, _ = fast_state(n, k, links)
    gamma = gamma / np.maximum(gamma.sum(axis=1, keepdims=True), 1e-300) * k * 1.0 + 0.01   # ~Gamma(100,.01)-scale rows
    lam = 1.0 + rng.gamma(100.0, 0.01, size=(k, 2))
    return links, heldout, hy, shuffled, gamma, lam


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--report", type=int, default=100, help="e2e: held-out evaluation cadence")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    print(json.dumps(run(args)))


def run(args):
    """One measurement of the -rnode -stratified path; returns the JSON record (bench.py embeds it as
    `secondary_path_fa2`)."""
    import torch
    from svinet_b200.fa2_engine import Fa2Engine
    if not torch.cuda.is_available():
        raise SystemExit("bench_fa2.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(0)
    n, k, target = WORKLOADS[args.workload]
    t0 = time.time()
    links, heldout, hy, shuffled, gamma, lam = make_problem(n, k, target, "cuda:0")
    torch.cuda.empty_cache()
    t_gen = time.time() - t0
    stream = torch.cuda.current_stream()
    t0 = time.time()
    eng = Fa2Engine(n, k, device=0, stream=stream.cuda_stream)
    eng.set_state(gamma, lam)
    eng.set_graph(links, heldout, shuffled)
    t_create = time.time() - t0
    seed = 20261017

    # ---- device-resident: svi_fa2_run, minibatches drawn on the device ----
    eng.run(0, args.warmup, seed, count=False)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    info0 = None
    e0.record(stream)
    t_wall0 = time.perf_counter()
    eng.run(args.warmup, args.steps, seed, count=False)
    e1.record(stream)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    total_ms = e0.elapsed_time(e1)
    # pairs / rounds of the timed region: replay the draws (cheap, no state change) for the exact count
    pairs = rounds_pairs = 0
    types = [0, 0]
    for it in range(args.warmup, args.warmup + args.steps):
        typ, start, pr = eng.draw(it, seed, cap=1)
        types[typ] += 1
        pairs += eng.info()["last_npairs"]
    value = pairs / (total_ms * 1e-3)

    # ---- eager mode (svi_fa2_config.eager_blend = 1): the reference's explicit O(N*K) pass every iteration ----
    # timed on a second handle with the same draws; its blend kernel is the bandwidth-bound kernel of this path
    eager = Fa2Engine(n, k, device=0, stream=stream.cuda_stream, eager_blend=1)
    eager.set_state(gamma, lam)
    eager.set_graph(links, heldout, shuffled)
    eager.run(0, args.warmup, seed, count=False)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    eager.run(args.warmup, args.steps, seed, count=False)
    g1.record(stream)
    torch.cuda.synchronize()
    eager_ms = g0.elapsed_time(g1) / args.steps
    it_e = args.warmup + args.steps
    empty = np.zeros((0, 2), np.uint32)
    for _ in range(3):
        eager.step(it_e, 0, 0, empty); it_e += 1
    torch.cuda.synchronize()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nb = 20
    b0.record(stream)
    for _ in range(nb):
        eager.step(it_e, 0, 0, empty); it_e += 1
    b1.record(stream)
    torch.cuda.synchronize()
    blend_ms = b0.elapsed_time(b1) / nb
    eager.close()
    it_next = args.warmup + args.steps
    info = eng.info()
    blend_bytes = 2 * n * info["ld"] * 8
    peak, peak_src = measured_peak_hbm()

    # ---- mean rounds per pair on one non-informative minibatch ----
    rounds_per_pair = None
    for it in range(1000, 1064):
        typ, start, pr = eng.draw(it, seed)
        if typ == 1 and len(pr):
            eng.step(it_next, typ, start, pr); it_next += 1
            i2 = eng.info()
            rounds_per_pair = i2["last_rounds"] / max(1, i2["last_npairs"])
            break

    # ---- end to end: host-chosen minibatches through svi_fa2_step, periodic report ----
    e2e_steps = min(args.steps, 100)
    plans = [eng.draw(it, seed) for it in range(2000, 2000 + e2e_steps)]
    hp = np.ascontiguousarray(heldout[:, 0]); hq = np.ascontiguousarray(heldout[:, 1])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_pairs = 0
    for i, (typ, start, pr) in enumerate(plans):
        eng.step(it_next, typ, start, pr); it_next += 1
        h2d += pr.nbytes
        e2e_pairs += len(pr)
        if (i + 1) % args.report == 0 or i + 1 == len(plans):
            # the reference's report: heldout_likelihood (src/fastamm2.cc:652-671); the model is written when the
            # run ends (save_model on terminate, :673-686), not per report
            ll = eng.heldout(hp, hq, hy)
            h2d += hp.nbytes + hq.nbytes + hy.nbytes
            d2h += ll.nbytes
    eng.sync()
    dt = time.perf_counter() - t0
    e2e = {"value": e2e_pairs / dt, "unit": UNIT, "steps": e2e_steps, "iterations_per_s": e2e_steps / dt,
           "h2d_bytes_per_step": h2d / e2e_steps, "d2h_bytes_per_step": d2h / e2e_steps,
           "what": "svi_fa2_step with host pair lists (double-buffered pinned staging); svi_fa2_heldout every %d "
                   "iterations (host pairs in, host log-likelihoods out)" % args.report,
           "heldout_mean_loglik": float(ll.mean())}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": total_ms / args.steps, "iterations_per_s": args.steps / (total_ms * 1e-3),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "%s: synthetic MMSB n=%d k=%d links=%d, -rnode -stratified (FastAMM2)" % (
                          args.workload, n, k, links.shape[0]), "n": n, "k": k, "links": int(links.shape[0]),
                      "minibatches": "device Philox draws: %d link sets, %d non-informative sets of n/10" % tuple(types),
                      "pairs_in_timed_region": int(pairs), "rounds_per_pair_noninf": rounds_per_pair,
                      "l2": "gamma %.2f GB %s L2" % (n * info["ld"] * 8 / 1e9, "exceeds" if n * info["ld"] * 8 > 126e6 else "fits"),
                      "tile": "G%d x V%d" % (info["lanes"], info["vec"]), "pair_blocks": info["pair_blocks"],
                      "decay": "lazy (scalar decay factor, no O(N*K) pass; svi_fa2_config.eager_blend = 0)",
                      "eager_blend_ms_per_step": eager_ms},
           "clocks": clocks, "e2e": e2e, "gpu_launches": info["kernels_per_step"] * args.steps,
           "roofline": {"bound": "hbm", "kernel": "k_fa2_blend", "achieved": blend_bytes / (blend_ms * 1e-3) / 1e9,
                        "peak": peak, "unit": "GB/s", "frac": blend_bytes / (blend_ms * 1e-3) / 1e9 / peak,
                        "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(blend_bytes),
                        "note": "eager mode only (the default lazy mode never launches it); timed as a whole "
                                "empty-minibatch step (prep + idle pair kernel + blend + lambda): %.3f ms" % blend_ms},
           "setup_s": {"generate": t_gen, "create+upload": t_create}, "wall_s_timed_region": t_wall}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_fa2_port(k, args.cpu_budget)      # the oracle leg lives in bench.py
    eng.close()
    return out


if __name__ == "__main__":
    main()
