"""Parity of the TMA ring sweeps (svi_ls_ring.cuh) in the regime the benchmark runs them in -- needs a B200.

The small-graph tests of test_gpu_parity.py collapse the automatic segment length to 16, i.e. they check
the ring kernels with at most two chunks of neighbours per segment.  Here: K in the ring range (32 < K <=
256, incl. the benchmarked K = 200 tile G8 x V13 and the K = 100 tile G4 x V13), average degree ~300 so that
a segment spans dozens of ring laps, one hub whose list splits into many segments (or one 2000-neighbour
segment), 35 % of the nodes converged (shortcut links interleaved with full-phi links inside every chunk),
and the segment lengths 256 (the benchmark's), 37 (ragged) and 4096 (clamped to 1024).  Every sweep is compared with the
oracle: gamma / lambda / K-vectors to 1e-9 relative, converged / active_comms / link-community membership
bit-exact.  A wrong or stale ring slot changes gamma rows by O(1/deg) and cannot hide under 1e-9.

`test_full_sweeps_at_config3_against_oracle` is BASELINE.json configs[2] (n = 1e5, K = 100, 5e6 links) run
sweep for sweep against the oracle (~30 s of oracle per sweep).
"""
import numpy as np
import pytest

import oracle_py as orc
from svinet_b200 import synth
from svinet_b200.engine import LinkSamplingEngine
from test_gpu_parity import TOL, compare_sweep, engine_from_state, rel_err

pytestmark = pytest.mark.gpu


def oracle_state(n, k, links, gamma, lam, conv=None):
    st = orc.State.alloc(n, k, links.shape[0])
    c = st.c
    c.alpha, c.eta0, c.eta1, c.ones = 1.0 / k, 1.0, 1.0, links.shape[0]
    st.arr("links")[:] = links
    tl = np.zeros(n)
    np.add.at(tl, links.ravel().astype(np.int64), 2.0)
    st.arr("tl")[:] = tl
    st.arr("gamma")[:] = gamma
    st.arr("gammanext")[:] = c.alpha
    st.arr("lambda_")[:] = lam
    st.arr("lambdanext")[:] = lam
    if conv is not None:
        st.arr("converged")[:] = conv
    st.refresh_expectations()
    return st


def dense_graph_with_hub(n, k, avg_deg, seed):
    base = synth.mmsb_links(n, k, n * avg_deg // 2, seed=seed)
    hub = np.stack([np.zeros(n - 1, dtype=np.uint32), np.arange(1, n, dtype=np.uint32)], 1)
    return np.unique(np.concatenate([base, hub]), axis=0).astype(np.uint32)


@pytest.mark.parametrize("k", [64, 100, 200, 256])
def test_ring_sweeps_long_segments_hub_and_converged(k):
    n = 2000
    links = dense_graph_with_hub(n, k, 300, seed=900 + k)
    rng = np.random.default_rng(k)
    deg = np.bincount(links.ravel().astype(np.int64), minlength=n).astype(np.float64)
    gamma = (deg[:, None] / k) * (1.0 + 0.5 * rng.random((n, k))) + 1.0 / k
    hot = rng.integers(0, k, n)
    gamma[np.arange(n), hot] *= 4.0                      # a clear arg-max per node, ties stay possible elsewhere
    conv = np.zeros(n, dtype=np.uint32)
    who = rng.random(n) < 0.35
    conv[who] = rng.integers(1, k + 1, who.sum())
    conv[1:8] = k                                        # the pc == K corner of the s3 shortcut (Q4)
    st = oracle_state(n, k, links, gamma, np.ones((k, 2)), conv)
    engines = [engine_from_state(st, links.shape[0], seg_len=s) for s in (256, 37, 4096)]
    for e in engines:
        info = e.info()
        assert info["ring_depth"] > 0, "K=%d must run the ring sweeps" % k
    assert engines[1].info()["segments_phi"] > 3 * n     # degrees in the hundreds in ragged pieces of <= 37 neighbours
    assert engines[2].info()["seg_len"] == 1024          # clamped: the hub's two parts hold ~1000 rows each
    for it, ann, wc in [(0, 1, 0), (1, 1, 1), (2, 0, 1)]:
        st.step(it, ann, wc)
        for e in engines:
            e.step(it, ann, wc)
            compare_sweep(e, st, "k=%d seg_len=%d iter %d" % (k, e.info()["seg_len"], it), check_member=bool(wc))
    assert st.c.cnt_shortcut > 0 and st.c.cnt_dense > 0
    for e in engines:
        e.close()
    st.free()


@pytest.mark.parametrize("k", [20, 40, 64, 100, 200])
def test_active_set_branch_on_a_settled_state(k):
    """iter > 1000 (:634-681) on a state where the branch really fires: the oracle first runs 20 sweeps from a
    random start, so that many rows have settled on < K/10 active communities (a random concentrated state loses
    its concentration in one sweep, and the branch would silently never run).  Ring tilings with long segments for
    K > 32 (the SPARSE instantiation of k_sweep_ring), k_phi<SPARSE> for K = 20."""
    n = 1500
    links = dense_graph_with_hub(n, k, 120, seed=400 + k)
    gamma, lam = synth.random_state(n, k, links, seed=k)
    st = oracle_state(n, k, links, gamma, lam)
    for it in range(20):
        st.step_omp(it < 10, 0, 0)           # all-cores oracle leg: only used to reach a settled state quickly
    eng = engine_from_state(st, links.shape[0], seg_len=256)
    assert (eng.info()["ring_depth"] > 0) == (k > 32)
    st.step(20, 0, 0); eng.step(20, 0, 0)    # one ordinary sweep from the same start brings the device's masks in line
    compare_sweep(eng, st, "warm")
    sparse = dense = 0
    for it, wc in [(1001, 1), (1002, 0), (1003, 1)]:
        st.step(it, 0, wc); eng.step(it, 0, wc)
        compare_sweep(eng, st, "k=%d iter %d" % (k, it), check_member=bool(wc))
        sparse += st.c.cnt_sparse
        dense += st.c.cnt_dense
    assert sparse > 0 and dense > 0, (sparse, dense)
    eng.close(); st.free()


@pytest.mark.slow
def test_full_sweeps_at_config3_against_oracle():
    """BASELINE.json configs[2]: n = 100000, K = 100, 5e6 links, the bench's own generator, seed and start
    state; two full sweeps (annealing + tally, then a sweep with 20 % of the nodes forced converged) against
    the oracle."""
    import torch
    n, k = 100_000, 100
    links = synth.mmsb_links(n, k, 5_000_000, seed=1234, device="cuda")
    torch.cuda.empty_cache()
    deg = np.bincount(links.ravel().astype(np.int64), minlength=n).astype(np.float64)
    rng = np.random.default_rng(5)
    gamma = (deg[:, None] / k) * (1.0 + 0.2 * rng.random((n, k))) + 1.0 / k
    st = oracle_state(n, k, links, gamma, np.ones((k, 2)))
    eng = engine_from_state(st, links.shape[0])
    info = eng.info()
    assert info["ring_depth"] > 0 and info["seg_len"] == 256
    st.step(0, 1, 1); eng.step(0, 1, 1)
    compare_sweep(eng, st, "config 3 sweep 0", check_member=True)
    who = rng.random(n) < 0.2
    conv = st.arr("converged")
    conv[who & (conv == 0)] = rng.integers(1, k + 1, n)[who & (conv == 0)]
    eng.set_converged(conv)
    st.step(1, 0, 1); eng.step(1, 0, 1)
    compare_sweep(eng, st, "config 3 sweep 1 (20% converged)", check_member=True)
    assert st.c.cnt_shortcut > 0
    eng.close(); st.free()
