#!/usr/bin/env python3
"""bench.py -- link-sampling edge-updates/sec on synthetic MMSB graphs (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2s|tiny] [--impl ours|reference]

One "step" = one full variational iteration (phi sweep + mean indicators + s3 sweep + lambda finish +
expectation refresh + prune == the loop body src/linksampling.cc:584-761) over ALL training links of the
workload.  `value` = links x steps / device time, state resident in HBM.  `e2e` = the same step driven
through the C ABI with HOST buffers the way the reference-facing caller drives it every iteration:
svi_ls_step, the held-out likelihood (svi_ls_heldout: host pair lists in, host log-likelihoods out -- the
"loss" the reference computes every iteration, src/linksampling.cc:778-780) and the link-community
membership of the sweep (svi_ls_get_membership, host bits out; log_communities, :785).
`e2e.state_roundtrip` is the pessimistic variant that also uploads and downloads the whole gamma/lambda
state around every step (svi_ls_set_state / svi_ls_get_state).

Under torchrun (N > 1) every rank owns an edge-balanced node block (svinet_b200/sharded.py); timing is
CUDA events on the launching stream, max over ranks.

`--impl reference` times the UNMODIFIED reference binary (oracle/_ref/svinet_ref, 1 thread -- the
reference path is serial) on a bounded sample of the same workload; if the binary is absent it falls
back to the oracle port.  That arm and the `cpu_baseline` leg are the only places this file touches
oracle/.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (n, k, links)          BASELINE.json configs[3] is the one the metric is quoted on
    "c4": (1_000_000, 200, 100_000_000),
    "c3": (100_000, 100, 5_000_000),
    "c2s": (17_903, 20, 196_972),          # AstroPh-shaped synthetic
    "tiny": (2_000, 20, 20_000),
    "widek": (50_000, 2048, 1_000_000),    # K > 1024: the block-per-row kernels (svi_ls_wide.cuh); not a BASELINE config
}
METRIC = "link_sampling_edge_updates_per_sec"
UNIT = "edge-updates/s"


def measured_peak_hbm():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
        os.close(fd)
        self.f = open(self.path, "w")
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                     stderr=subprocess.DEVNULL)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, reasons, smax, pw = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out


# --------------------------------------------------------------------------------------------
def fast_state(n, k, links, seed=5):
    """Synthetic gamma shaped like init_gamma2's (every link adds a normalised uniform K-vector to both
    endpoints, src/linksampling.cc:374-401): row p ~ deg(p)/K * (1 + noise).  lambda = eta = (1,1)."""
    deg = np.bincount(links.ravel().astype(np.int64), minlength=n).astype(np.float64)
    rng = np.random.default_rng(seed)
    gamma = np.empty((n, k), dtype=np.float64)
    step = max(1, (1 << 25) // k)
    for s in range(0, n, step):
        e = min(n, s + step)
        gamma[s:e] = (deg[s:e, None] / k) * (1.0 + 0.2 * rng.random((e - s, k))) + 1.0 / k
    return gamma, np.ones((k, 2), dtype=np.float64)


def heldout_pairs(n, links, count, seed=11):
    """`count` held-out pairs, half links half random non-link candidates (like the reference's 1% set)."""
    rng = np.random.default_rng(seed)
    half = max(1, count // 2)
    li = links[rng.integers(0, links.shape[0], half)]
    p = rng.integers(0, n, half).astype(np.uint32)
    q = ((p.astype(np.int64) + 1 + rng.integers(0, n - 1, half)) % n).astype(np.uint32)
    pp = np.concatenate([li[:, 0], np.minimum(p, q)]).astype(np.uint32)
    qq = np.concatenate([li[:, 1], np.maximum(p, q)]).astype(np.uint32)
    yy = np.concatenate([np.ones(half, np.uint8), np.zeros(half, np.uint8)])
    return pp, qq, yy


def phi_kernel_bytes(info, k):
    """Algorithmic bytes of ONE phi sweep of this implementation, no node converged:
    per half-edge one neighbour row (ld*8) + its column index (4); per segment one self row read (ld*8), one
    partial row written (ld*8) and the segment descriptor (16)."""
    ld = info["ld"]
    return info["half_edges_phi"] * (ld * 8 + 4) + info["segments_phi"] * (2 * ld * 8 + 16)


def step_bytes_survey(nlinks, n, k, s=8):
    """bytes_iter = nlinks*(6*K*s+16) + 7*N*K*s (push-form accounting)."""
    return nlinks * (6 * k * s + 16) + 7 * n * k * s


def verify_sampled_rows(step_once, get_state, get_converged, n, k, links, alpha, sample=512, seed=99, compute=True):
    """Parity spot-check AT THE BENCHMARKED SIZE: run one more iteration (annealing off, tally on) and recompute
    the new gamma rows of `sample` random nodes on the host in the REFERENCE's formulation -- Elogpi = psi(gamma) -
    psi(sum gamma) with scipy's digamma (independent of both the device's and the oracle's), per link the running
    log-sum-exp of src/linksampling.cc:685-694 and exp(phi - r) (src/matrix.hh:320-325), the one-hot shortcut of
    :619-631, then compute_mean_indicators (:526-545) -- from the state downloaded before the iteration.
    A stale / wrong neighbour row shifts a gamma row by O(1/degree); the bar is 1e-9 relative."""
    from scipy.special import digamma
    rng = np.random.default_rng(seed)
    deg = np.bincount(links.ravel().astype(np.int64), minlength=n)
    cand = np.flatnonzero(deg > 0)
    pick = np.sort(rng.choice(cand, size=min(sample, cand.size), replace=False))
    is_s = np.zeros(n, dtype=bool)
    is_s[pick] = True
    m0, m1 = is_s[links[:, 0]], is_s[links[:, 1]]
    src = np.concatenate([links[m0, 0], links[m1, 1]]).astype(np.int64)
    dst = np.concatenate([links[m0, 1], links[m1, 0]]).astype(np.int64)
    rows = np.unique(np.concatenate([src, dst]))
    g0, lam0 = get_state()
    gr = g0[rows].copy()
    del g0
    conv0 = get_converged().astype(np.int64)
    step_once()
    g1, _ = get_state()
    got = g1[pick].copy()
    del g1
    if not compute:
        return None
    elogpi = digamma(gr) - digamma(gr.sum(1, keepdims=True))
    elogbeta0 = digamma(lam0[:, 0]) - digamma(lam0.sum(1))
    si, di = np.searchsorted(rows, src), np.searchsorted(rows, dst)
    pc, qc = conv0[src], conv0[dst]
    short = (pc != 0) != (qc != 0)
    full = ~short
    ep, eq = elogpi[si[full]], elogpi[di[full]]
    r = None
    with np.errstate(over="ignore"):
        for kk in range(k):
            x = ep[:, kk] + eq[:, kk] + elogbeta0[kk]
            r = x if kk == 0 else np.where(x < r, r + np.log(1 + np.exp(x - r)),
                                           x + np.log(1 + np.exp(r - x)))
    phi = np.exp(ep + eq + elogbeta0[None, :] - r[:, None]) if full.any() else np.zeros((0, k))
    out_idx = np.searchsorted(pick, src)
    acc = np.zeros((pick.size, k))
    np.add.at(acc, out_idx[full], phi)
    np.add.at(acc, (out_idx[short], np.where(pc[short] != 0, pc[short], qc[short]) - 1), 1.0)
    tl = 2.0 * deg[pick].astype(np.float64)[:, None]
    gn = alpha + acc
    mphi = (gn - alpha) / tl
    want = gn + (n - tl - 1.0) * mphi
    err = float(np.max(np.abs(got - want) / np.abs(want)))
    return {"rows": int(pick.size), "half_edges": int(src.size), "shortcut_half_edges": int(short.sum()),
            "max_rel_err": err, "tol": 1e-9, "ok": bool(err <= 1e-9),
            "what": "gamma rows of sampled nodes after one more iteration (annealing off) vs a host recomputation in "
                    "the reference's formulation (scipy digamma, running log-sum-exp per link, compute_mean_indicators)"}


# --------------------------------------------------------------------------------------------
def run_ours(args):
    # a shard runs four streams of its own beside torch's, with flag-wait kernels at their heads: enough hardware queues
    # that no two of them alias (include/svi_ls.h, svi_ls_mg_step).  Read when CUDA initialises.
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import torch
    import torch.distributed as dist
    from svinet_b200 import synth
    from svinet_b200.engine import LinkSamplingEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout must carry exactly one JSON line: NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) goes there
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    n, k, target = WORKLOADS[args.workload]
    t0 = time.time()
    links = synth.mmsb_links(n, k, target, seed=1234, device=str(dev))      # identical on every rank
    torch.cuda.empty_cache()
    nlinks = links.shape[0]
    gamma0, lam0 = fast_state(n, k, links)
    t_gen = time.time() - t0

    stream = torch.cuda.current_stream()
    if world > 1:
        # the shards wait for each other's flags on their streams: a stream of its own, not the legacy default one
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
    t0 = time.time()
    peer, mg_ms = False, None
    if world == 1:
        eng = LinkSamplingEngine(n, k, links, device=local_rank, stream=stream.cuda_stream)
        runner = None
        eng.set_state(gamma0, lam0)
        step = lambda it: eng.step(it, True, it > 0)
    else:
        from svinet_b200.sharded import ShardedLinkSampling
        runner = ShardedLinkSampling(n, k, links, rank=rank, world=world, device=local_rank,
                                     stream=stream.cuda_stream, exchange=args.exchange, chunks=args.chunks)
        eng = runner.eng
        peer = runner.exchange == "peer"
        runner.set_state(gamma0, lam0)
        step = lambda it: runner.step(it, True, it > 0)
    info = eng.info()
    t_create = time.time() - t0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    it = 0
    for _ in range(args.warmup):
        step(it); it += 1
    barrier()
    if world > 1 and peer:
        eng.mg_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        ev[s][0].record(stream)
        if world == 1:
            eng.step(it, True, it > 0)            # svi_ls_step: the iteration as one CUDA graph launch
        else:
            runner.step(it, True, it > 0, events=ev[s])
        ev[s][4].record(stream)
        it += 1
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    total_ms = ev[0][0].elapsed_time(ev[-1][4])
    if world == 1:
        # per-phase times: the same iteration through the phase-level entry points, outside the timed region
        pev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(min(args.steps, 10))]
        for s in range(len(pev)):
            pev[s][0].record(stream)
            eng.phase_phi(it, it > 0); pev[s][1].record(stream)
            eng.phase_node(); pev[s][2].record(stream)
            eng.phase_s3(); pev[s][3].record(stream)
            eng.phase_finish(True); pev[s][4].record(stream)
            it += 1
        torch.cuda.synchronize()
        ev = pev
    if world == 1:
        phase_ms = [float(np.mean([ev[s][i].elapsed_time(ev[s][i + 1]) for s in range(len(ev))])) for i in range(4)]
    elif peer:
        mg_ms, _ = eng.mg_timing(False, read=True)
        phase_ms = [mg_ms["wait_b_rows"] + mg_ms["phi+node"], mg_ms["allreduce_sum_s1_s2"] + mg_ms["refresh"] + mg_ms["wait_mphi_rows"],
                    mg_ms["s3"], mg_ms["allreduce_s3+lambda"] + mg_ms["drain_own_pushes"]]
    else:
        phase_ms = runner.phase_ms(ev)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = nlinks * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the C ABI with HOST buffers -------------------------------------------------
    # What the reference-facing caller (svinet_b200/host/linksampling.cc::infer, the drop-in for
    # src/linksampling.cc:571-789 with its default reportfreq = 1) does every iteration: svi_ls_step, then the
    # held-out likelihood of the validation pairs (pair lists host -> device, log-likelihoods device -> host,
    # src/linksampling.cc:778-780) and the link-community membership of the sweep (device -> host,
    # log_communities :785).  The graph and the variational state stay resident, as they do in that caller.
    # `state_roundtrip` additionally re-uploads and downloads the WHOLE state (gamma, lambda) around every
    # step -- a pattern no caller has, kept as the pessimistic bound the previous profiles quoted.
    e2e = None
    if world == 1:
        pin_g = torch.empty((n, k), dtype=torch.float64).pin_memory()
        pin_l = torch.empty((k, 2), dtype=torch.float64).pin_memory()
        eng.get_state_ptr(pin_g.data_ptr(), pin_l.data_ptr())
        hp, hq, hy = heldout_pairs(n, links, max(2, min(nlinks // 100, 2_000_000)))

        def pinned(a):          # the caller's host buffers of the e2e loop live in pinned memory
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t, t.numpy()
        keep = [pinned(a) for a in (hp, hq, hy, np.empty(hp.shape[0], np.float64),
                                    np.empty((n, (k + 31) // 32), np.uint32))]
        hp, hq, hy, ll, bits = [b for _, b in keep]
        e2e_steps = max(1, min(args.steps, 10))
        eng.step(it, True, True); it += 1
        eng.heldout(hp, hq, hy, out=ll); eng.membership_bits(out=bits)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(e2e_steps):
            eng.step(it, True, True); it += 1
            eng.heldout(hp, hq, hy, out=ll)
            eng.membership_bits(out=bits)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": nlinks * e2e_steps / dt, "unit": UNIT, "steps": e2e_steps,
               "h2d_bytes_per_step": hp.nbytes + hq.nbytes + hy.nbytes,
               "d2h_bytes_per_step": ll.nbytes + bits.nbytes,
               "what": "per iteration, as the drop-in CLI does: svi_ls_step + svi_ls_heldout(pinned host pairs -> pinned "
                       "host log-likelihoods) + svi_ls_get_membership(pinned host bits); graph and state resident",
               "heldout_mean_loglik": float(ll.mean())}
        rt_steps = max(1, min(args.steps, 5))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(rt_steps):
            eng.set_state_ptr(pin_g.data_ptr(), pin_l.data_ptr())
            eng.step(it, True, True); it += 1
            eng.heldout(hp, hq, hy, out=ll)
            eng.get_state_ptr(pin_g.data_ptr(), pin_l.data_ptr())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        state_bytes = (n * k + 2 * k) * 8
        e2e["state_roundtrip"] = {
            "value": nlinks * rt_steps / dt, "unit": UNIT, "steps": rt_steps,
            "h2d_bytes_per_step": state_bytes + hp.nbytes + hq.nbytes + hy.nbytes,
            "d2h_bytes_per_step": state_bytes + ll.nbytes,
            "what": "svi_ls_set_state(pinned host) + svi_ls_step + svi_ls_heldout + svi_ls_get_state(pinned host)"}
        del pin_g, pin_l
    else:
        e2e_steps = max(1, min(args.steps, 10))
        e2e = runner.e2e(step_fn=lambda i: runner.step(i, True, True), it0=it, steps=e2e_steps, nlinks=nlinks, unit=UNIT,
                         heldout=heldout_pairs(n, links, max(2, min(nlinks // 100, 2_000_000))))
        it += e2e_steps + 1

    # ---- second regime: a fraction of the nodes already converged (where real runs live: the recorded reference run
    # handles 35 % of its links on the one-hot shortcut).  Roofline on the bytes of the rows
    # actually fetched.
    late = None
    if world == 1 and args.converged_frac > 0:
        rng = np.random.default_rng(0)
        conv = np.zeros(n, dtype=np.uint32)
        who = rng.random(n) < args.converged_frac
        conv[who] = rng.integers(1, k + 1, int(who.sum()))
        eng.set_state(gamma0, lam0)
        eng.set_converged(conv)
        for i in range(3):
            eng.step(i, True, True)
        lsteps = max(1, min(args.steps, 10))
        lev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(lsteps)]
        for s_ in range(lsteps):
            lev[s_][0].record(stream)
            eng.phase_phi(3 + s_, True); lev[s_][1].record(stream)
            eng.phase_node(); lev[s_][2].record(stream)
            eng.phase_s3(); lev[s_][3].record(stream)
            eng.phase_finish(True); lev[s_][4].record(stream)
        torch.cuda.synchronize()
        lms = [float(np.mean([lev[s_][i].elapsed_time(lev[s_][i + 1]) for s_ in range(lsteps)])) for i in range(4)]
        ltot = lev[0][0].elapsed_time(lev[-1][4]) / lsteps
        cv = eng.get_converged()[0] != 0                 # flags after the timed sweeps (more nodes may have converged)
        full_links = int((cv[links[:, 0]] == cv[links[:, 1]]).sum())
        ld_ = info["ld"]
        lbytes = 2 * full_links * (ld_ * 8 + 4) + 2 * (nlinks - full_links) * 4 + info["segments_phi"] * (2 * ld_ * 8 + 16)
        lgbs = lbytes / (lms[0] * 1e-3) / 1e9
        late = {"converged_frac_set": args.converged_frac, "converged_frac_end": float(cv.mean()),
                "full_phi_link_frac": full_links / nlinks, "ms_per_step": ltot, "value": nlinks / (ltot * 1e-3), "unit": UNIT,
                "phase_ms": {"phi": lms[0], "node": lms[1], "s3": lms[2], "finish": lms[3]},
                "roofline": {"bound": "hbm", "achieved": lgbs, "peak": measured_peak_hbm()[0], "unit": "GB/s",
                             "frac": lgbs / measured_peak_hbm()[0], "algorithmic_bytes_per_launch": int(lbytes),
                             "note": "phi sweep; bytes of the rows actually fetched (shortcut links fetch no row)"}}
        it = 0

    verify = None
    if not args.no_verify and (world == 1 or peer):
        t0 = time.time()
        if world == 1:
            verify = verify_sampled_rows(lambda: eng.step(it, False, True), eng.get_state,
                                         lambda: eng.get_converged()[0], n, k, links, 1.0 / k)
        else:
            # collective: every rank steps and publishes its gamma rows; rank 0 does the arithmetic
            def all_step():
                runner.step(it, False, True)
                eng.sync()
                dist.barrier()
            verify = verify_sampled_rows(all_step, runner.gather_state, lambda: eng.get_converged()[0], n, k, links,
                                         1.0 / k, compute=(rank == 0))
        if verify is not None:
            verify["seconds"] = time.time() - t0
        it += 1
    checksum = None
    if args.checksum:
        # the same S iterations from the same start state at any GPU count: the sums agree to rounding (the per-row
        # summation order does not depend on the sharding; the K-vector reductions do)
        if world == 1:
            eng.set_state(gamma0, lam0); eng.set_converged(np.zeros(n, dtype=np.uint32))
            for i in range(args.checksum):
                eng.step(i, True, i > 0)
        else:
            eng.sync(); dist.barrier()
            runner.set_state(gamma0, lam0); eng.set_converged(np.zeros(n, dtype=np.uint32))
            eng.sync(); dist.barrier()
            for i in range(args.checksum):
                runner.step(i, True, i > 0)
            eng.sync(); dist.barrier()
        g, lam = eng.get_state() if world == 1 else runner.gather_state()
        cv = eng.get_converged()[0]
        checksum = {"steps": args.checksum, "gamma_sum": float(g.sum()), "gamma_sq_sum": float((g * g).sum()),
                    "lambda_sum": float(lam.sum()), "converged_nodes": int((cv != 0).sum()),
                    "converged_label_sum": int(cv.astype(np.int64).sum())}
        del g

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return
    peak, peak_src = measured_peak_hbm()
    phi_bytes = phi_kernel_bytes(info, k)
    phi_gbs = phi_bytes / (phase_ms[0] * 1e-3) / 1e9
    survey_gbs = step_bytes_survey(nlinks, n, k) * args.steps / (total_ms * 1e-3) / 1e9 / max(world, 1)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: synthetic MMSB n=%d k=%d links=%d (BASELINE.json configs[3])" % (
                       args.workload, n, k, nlinks) if args.workload == "c4" else
                   "%s: synthetic MMSB n=%d k=%d links=%d" % (args.workload, n, k, nlinks),
                   "n": n, "k": k, "links": int(nlinks), "annealing": True, "write_comm": True,
                   "parallelism": "node-block shards x%d" % world if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (state %.2f GB per matrix), no flush" % (n * info["ld"] * 8 / 1e9)
                   if n * info["ld"] * 8 > 200e6 else "state fits L2; not flushed between steps (iterative workload)",
                   "seg_len": info["seg_len"], "tile": "G%d x V%d" % (info["lanes"], info["vec"])},
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": info["kernels_per_step"] * args.steps,
        "phase_ms": {"phi": phase_ms[0], "node": phase_ms[1], "s3": phase_ms[2], "finish": phase_ms[3]},
        "roofline": {"bound": "hbm", "kernel": "k_sweep_ring<Phi> (phi sweep: the two launches of a tally sweep + k_partition)"
                     if info["ring_depth"] else "k_phi", "achieved": phi_gbs, "peak": peak, "unit": "GB/s",
                     "frac": phi_gbs / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(phi_bytes),
                     "note": "pull-form bytes of this kernel (DESIGN.md section 4)",
                     "step_gbs_survey_8d_formula": survey_gbs,
                     "step_frac_survey_8d_formula": survey_gbs / peak},
        "verify": verify, "checksum": checksum, "late_run": late,
        "exchange": (runner.exchange if world > 1 else None), "mg_phase_ms": (mg_ms if world > 1 else None),
        "setup_s": {"generate": t_gen, "create+upload": t_create},
        "wall_s_timed_region": t_wall,
    }
    # DRAM traffic of the phi sweep: from the committed `ncu --set full` capture of the same command (it cannot be
    # measured inside an un-profiled run); the file names its source
    traffic_file = os.path.join(REPO, "profiles", "k_phi_traffic.json")
    if os.path.exists(traffic_file) and world == 1:
        try:
            tf = json.load(open(traffic_file))
            if tf.get("workload") == args.workload:
                out["roofline"]["traffic"] = tf.get("dram_bytes_per_launch")
                out["roofline"]["traffic_source"] = tf.get("source")
        except Exception:
            pass
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_port(k, budget_s=args.cpu_budget)
    if world == 1 and not args.no_fa2:
        # the declared secondary path (-rnode -stratified, class FastAMM2): its own
        # measurement (bench_fa2.py) at config 3, embedded so that the driver's run carries it
        try:
            eng.close()
            del eng
            torch.cuda.empty_cache()
            import bench_fa2
            fa = bench_fa2.run(argparse.Namespace(workload="c3", steps=200, warmup=10, report=100, no_cpu_baseline=True,
                                                  cpu_budget=0.0))
            out["secondary_path_fa2"] = {key: fa[key] for key in ("metric", "value", "unit", "ms_per_step", "iterations_per_s",
                                                                  "config", "e2e", "roofline", "gpu_launches")}
        except Exception as exc:          # never lose the main line over the secondary measurement
            out["secondary_path_fa2"] = {"error": repr(exc)}
    print(json.dumps(out))
    sys.stdout.flush()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
def sample_graph(k, nlinks_sample, avg_deg=200):
    """A bounded sample of the workload for CPU timing: same generator, same K, same average degree."""
    from svinet_b200 import synth
    n_s = max(64, int(2 * nlinks_sample / avg_deg))
    links = synth.mmsb_links(n_s, k, nlinks_sample, seed=4321, device="cpu")
    return n_s, links


def cpu_baseline_port(k, budget_s=20.0):
    """The oracle (C restatement) on a bounded sample of the workload, on this box's host cores: the serial sweep
    (oracle/oracle_ls.c -- the reference path itself is serial) and the all-cores sweep (oracle/oracle_ls_omp.c,
    OpenMP).  `value` (in the bench's unit) is the all-cores figure, `cores` the threads it used; the one-thread figure
    is reported beside it."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_py as orc
    per_edge_k = 6e-8                                   # ~54-60 ns per (edge,k) on the survey's Xeon
    sweeps = 2
    threads = max(1, min(orc.lib().orc_omp_max_threads(), os.cpu_count() or 1))
    # half of the budget for the serial sweeps, half for the parallel ones (the same number of sweeps, more links)
    ns1 = int(max(2000, min(2_000_000, 0.5 * budget_s / (sweeps * per_edge_k * k))))
    nsp = int(max(2000, min(8_000_000, ns1 * max(1, threads // 2))))

    def run(nlinks_sample, nthreads):
        n_s, links = sample_graph(k, nlinks_sample)
        gamma, lam = fast_state(n_s, k, links)
        st = orc.State.alloc(n_s, k, links.shape[0])
        c = st.c
        c.alpha, c.eta0, c.eta1, c.ones = 1.0 / k, 1.0, 1.0, links.shape[0]
        st.arr("links")[:] = links
        tl = np.zeros(n_s); np.add.at(tl, links.ravel().astype(np.int64), 2.0)
        st.arr("tl")[:] = tl
        st.arr("gamma")[:] = gamma; st.arr("gammanext")[:] = c.alpha
        st.arr("lambda_")[:] = lam; st.arr("lambdanext")[:] = lam
        st.refresh_expectations()
        step = (lambda i, wc: st.step(i, 1, wc)) if nthreads == 1 else (lambda i, wc: st.step_omp(1, wc, nthreads))
        step(0, 0)                                     # warm
        t0 = time.perf_counter()
        for i in range(sweeps):
            step(1 + i, 1)
        dt = time.perf_counter() - t0
        st.free()
        return links.shape[0] * sweeps / dt, n_s, links.shape[0]

    v1, n1, l1 = run(ns1, 1)
    out = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": "oracle/oracle_ls.c, %d sweeps over a synthetic MMSB sample n=%d k=%d links=%d "
                     "(same generator and average degree as the workload)" % (sweeps, n1, k, l1),
           "host_cpus": os.cpu_count()}
    if threads > 1:
        vp, np_, lp = run(nsp, threads)
        out = {"value": vp, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "oracle/oracle_ls_omp.c (OpenMP, %d threads), %d sweeps over a synthetic MMSB sample n=%d k=%d "
                         "links=%d (same generator and average degree as the workload)" % (threads, sweeps, np_, k, lp),
               "single_thread": {"value": v1, "cores": 1, "sample": out["sample"]},
               "host_cpus": os.cpu_count()}
    return out


def cpu_baseline_fa2_port(k, budget_s=15.0):
    """cpu_baseline leg of bench_fa2.py (the -rnode -stratified path): oracle/oracle_fa2.c, 1 thread, bounded sample."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import oracle_py as orc
    from svinet_b200 import synth
    # one non-informative iteration costs ~ (n/10) pairs x rounds x 2K x 3 transcendentals (~20 ns each)
    per_pair = 50 * 2 * k * 3 * 20e-9
    n_s = int(max(400, min(40000, 10 * budget_s / (6 * per_pair))))
    links = synth.mmsb_links(n_s, k, n_s * 20, seed=4321, device="cpu")
    used = np.unique(links)
    remap = np.zeros(n_s, dtype=np.int64); remap[used] = np.arange(used.size)
    g = orc.Graph.from_pairs(remap[links.astype(np.int64)].astype(np.uint32), used.size)
    m = orc.Fa2Model(g, k, max_iterations=0, reportfreq=1 << 30)
    m.run(2)
    pairs = iters = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s and iters < 64:
        before = m.total_pairs_sampled
        m.run(1)
        pairs += m.total_pairs_sampled - before
        iters += 1
    dt = time.perf_counter() - t0
    out = {"value": pairs / dt, "unit": "pair-updates/s", "cores": 1, "kind": "port", "iterations_per_s": iters / dt,
           "sample": "oracle/oracle_fa2.c, %d iterations (reference mt19937 minibatches) on a synthetic MMSB sample "
                     "n=%d k=%d links=%d" % (iters, g.n, k, g.ones), "host_cpus": os.cpu_count()}
    m.close(); g.close()
    return out



def run_reference(args):
    """The reference's own CPU implementation on this box's host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, k, _ = WORKLOADS[args.workload]
    ref_bin = os.path.join(REPO, "oracle", "_ref", "svinet_ref")
    budget = args.ref_budget
    sweeps_total = 2 * (2 * args.warmup + args.steps + 2)   # runs M1 = warmup and M2 = warmup + steps (+1 each), twice
    per_edge_k = 6e-8
    ns = int(max(2000, min(1_000_000, budget / (sweeps_total * per_edge_k * k))))
    n_s, links = sample_graph(k, ns)
    if os.path.exists(ref_bin):
        d = tempfile.mkdtemp(prefix="refarm_")
        try:
            # ids are written 1-based in link order so the reference sees exactly n_s distinct nodes
            used = np.unique(links)
            remap = np.zeros(n_s, dtype=np.int64); remap[used] = np.arange(used.size)
            np.savetxt(os.path.join(d, "g.txt"), remap[links.astype(np.int64)], fmt="%d", delimiter="\t")
            def run(m):
                t0 = time.perf_counter()
                subprocess.check_call([ref_bin, "-file", "g.txt", "-n", str(used.size), "-k", str(k), "-link-sampling",
                                       "-rfreq", "100000", "-accuracy", "-max-iterations", str(m), "-no-stop"],
                                      cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                return time.perf_counter() - t0
            m1 = max(1, args.warmup)
            m2 = m1 + args.steps
            # the constructor (ingest + init_gamma2, seconds) is in both runs and jitters: best of two each
            t1 = min(run(m1), run(m1))
            t2 = min(run(m2), run(m2))
            per_step = max(t2 - t1, 1e-9) / args.steps
            kind, cores = "reference", 1
            sample = ("oracle/_ref/svinet_ref (unmodified reference, g++ -O2, 1 thread: the path is serial) "
                      "-link-sampling -accuracy -rfreq 100000 on a synthetic MMSB sample n=%d k=%d links=%d; "
                      "per-step = (wall(M=%d) - wall(M=%d)) / %d (best of two runs each), cancelling the constructor"
                      % (used.size, k, links.shape[0], m2, m1, args.steps))
        finally:
            shutil.rmtree(d, ignore_errors=True)
    else:
        cb = cpu_baseline_port(k, budget_s=budget / 4)
        per_step = links.shape[0] / cb["value"]
        kind, cores, sample = "port", 1, cb["sample"] + " (oracle/_ref absent: oracle port)"
        links = links[: int(cb["value"] * per_step)]
    value = links.shape[0] / per_step
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": "%s: synthetic MMSB n=%d k=%d (bounded sample of it: links=%d)" % (
               args.workload, n, k, links.shape[0]), "k": k},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                            "host_cpus": os.cpu_count()},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="(default on at 1 GPU) recompute sampled gamma rows on the host")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-fa2", action="store_true", help="skip the embedded bench_fa2 record (secondary path)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = svi_ls_mg_step (rows pushed over peer memory inside the library); "
                         "nccl = torch.distributed collectives between the phases (the library baseline)")
    ap.add_argument("--chunks", type=int, default=0, help="N > 1, peer exchange: pipeline chunks of a shard (0 = default)")
    ap.add_argument("--converged-frac", type=float, default=0.35,
                    help="1 GPU: also time the iteration with this fraction of the nodes converged (0 = skip)")
    ap.add_argument("--checksum", type=int, default=0, metavar="S",
                    help="also run S iterations from the start state and print sums of the resulting state")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds for the --impl reference arm")
    ap.add_argument("--path", default="ls", choices=["ls", "fa2"],
                    help="ls = -link-sampling (the BASELINE.json metric); fa2 = -rnode -stratified, forwarded to bench_fa2.py")
    args, rest = ap.parse_known_args()
    if args.path == "fa2":
        import bench_fa2
        sys.argv = [sys.argv[0]] + rest + (["--workload", args.workload] if "--workload" in sys.argv else []) + \
                   (["--steps", str(args.steps)] if "--steps" in sys.argv else []) + \
                   (["--warmup", str(args.warmup)] if "--warmup" in sys.argv else []) + \
                   (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
        return bench_fa2.main()
    if rest:
        ap.error("unrecognized arguments: %s" % " ".join(rest))
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
