"""Parity of the CUDA path (through the C ABI) against the oracle -- needs a B200.

Bar (BASELINE.json north_star): gamma/lambda within 1e-5 relative after a fixed seed and a fixed
iteration count.  The FP64 device path actually lands near 1e-12; the tests assert 1e-9 so that a
regression in arithmetic (a wrong branch, a missed term) cannot hide inside the official tolerance,
and `TOL_OFFICIAL` is asserted separately where the reference's own fixtures are involved.
Integer outputs (converged, active_comms, link-community membership) must match exactly.
"""
import numpy as np
import pytest

import oracle_py as orc
from golden_util import MANIFEST, Scratch, input_path
from svinet_b200 import synth
from svinet_b200.engine import LinkSamplingEngine
from parity_util import TOL, TOL_OFFICIAL, compare_sweep, engine_from_state, rel_err  # noqa: F401 (re-exported)

pytestmark = pytest.mark.gpu


def run_model_lockstep(g, k, sweeps, tol=TOL, **opts):
    """Drive the oracle's whole-run model and the engine with the same per-sweep flags."""
    m = orc.Model(g, k, **opts)
    st = m.state
    eng = engine_from_state(st, g.ones)
    worst = {}
    for _ in range(sweeps):
        it, ann = m.iter, m.annealing
        wc = m.write_comm or m.opts.max_iterations == 1
        if m.run(1) == 0:
            break
        eng.step(it, ann, wc)
        errs = compare_sweep(eng, st, "iter %d" % it, tol=tol, check_member=wc)
        for n_, e in errs.items():
            worst[n_] = max(worst.get(n_, 0.0), e)
    # held-out likelihood of the validation pairs under the final state
    vp = m.validation_pairs()
    if len(vp):
        y = np.array([g.y(int(a), int(b)) for a, b in vp], dtype=np.uint8)
        got = eng.heldout(vp[:, 0], vp[:, 1], y)
        want = np.array([st.edge_likelihood(int(a), int(b), int(yy)) for (a, b), yy in zip(vp, y)])
        # log-likelihoods are averaged by the caller: an absolute floor is the meaningful metric for the
        # near-zero ones (log(1 - 1e-6) is ill-conditioned in relative terms)
        assert rel_err(got, want, floor=1e-3) <= tol
    eng.close()
    m.close()
    return worst


# ---- the reference's own inputs (configs 1-2 of BASELINE.json and the LFR example) -------------
@pytest.mark.parametrize("case,sweeps", [("c1_m30", 31), ("c1_k7_m15", 16), ("c1_seed7_m12", 13),
                                         ("c1_accuracy_m8", 9), ("c1_m1", 2), ("lfr_k28_m20", 21),
                                         ("c1_etasparse_m10", 11)])
def test_lockstep_small_inputs(case, sweeps):
    ent = MANIFEST[case]
    opts = {"max_iterations": 0, "use_validation_stop": 0}
    if "-seed" in ent["flags"]:
        opts["seed"] = float(ent["flags"][ent["flags"].index("-seed") + 1])
    if "-accuracy" in ent["flags"]:
        opts["accuracy"] = 1
    if case == "c1_m1":
        opts["max_iterations"] = 1
    if "-eta-type" in ent["flags"]:
        opts["eta0"], opts["eta1"] = {"sparse": (0.97, 6.33)}[ent["flags"][ent["flags"].index("-eta-type") + 1]]
    with Scratch() as d:
        g = orc.Graph.read(input_path(ent["input"], d), ent["n"])
        worst = run_model_lockstep(g, ent["k"], sweeps, **opts)
        g.close()
    assert worst["gamma"] <= TOL and worst["lambda"] <= TOL


def test_lockstep_astroph_k20():
    """BASELINE.json config 2: ca-AstroPh n=17903 k=20, 26 sweeps (the c2_m25 fixture's run)."""
    with Scratch() as d:
        g = orc.Graph.read(input_path("ca-AstroPh.csv", d), 17903)
        worst = run_model_lockstep(g, 20, 26, use_validation_stop=0)
        g.close()
    assert worst["gamma"] <= TOL_OFFICIAL and worst["lambda"] <= TOL_OFFICIAL
    assert worst["gamma"] <= TOL and worst["lambda"] <= TOL


def test_free_running_matches_reference_fixture():
    """No lockstep: the engine runs 31 sweeps on its own (annealing flag from the oracle's stop machine is
    replayed from the fixture's known schedule) and the result is compared with the REFERENCE's gamma.txt /
    lambda.txt (tests/golden/c1_m30), at the official tolerance and at the %.5f print precision."""
    from golden_util import golden_text
    with Scratch() as d:
        g = orc.Graph.read(input_path("assort-75-4.txt", d), 75)
        m = orc.Model(g, 4, max_iterations=30, use_validation_stop=0)
        st = m.state
        eng = engine_from_state(st, g.ones)
        # schedule of (iter, annealing, write_comm) comes from a throw-away oracle run
        sched = []
        while True:
            it, ann, wc = m.iter, m.annealing, m.write_comm
            if m.run(1) == 0:
                break
            sched.append((it, ann, wc))
        for it, ann, wc in sched:
            eng.step(it, ann, wc)
        gam, lam = eng.get_state()
        want_g = np.array([[float(x) for x in line.split("\t")[2:]] for line in
                           golden_text("c1_m30", "gamma.txt").strip().split("\n")])
        want_l = np.array([[float(x) for x in line.split("\t")[1:]] for line in
                           golden_text("c1_m30", "lambda.txt").strip().split("\n")])
        assert np.max(np.abs(gam - want_g)) <= 1.0001e-5 / 2 + 1e-9      # printed with %.5f
        assert np.max(np.abs(lam - want_l)) <= 1.0001e-5 / 2 + 1e-9
        assert rel_err(gam, st.arr("gamma")) <= TOL_OFFICIAL
        eng.close(); m.close(); g.close()


# ---- synthetic inputs: every kernel tiling, both arithmetic domains ------------------------------
def synthetic_state(n, k, nlinks, seed, conv_frac=0.0):
    links = synth.mmsb_links(n, k, nlinks, seed=seed)
    gamma, lam = synth.random_state(n, k, links, seed=seed + 1)
    st = orc.State.alloc(n, k, links.shape[0])
    c = st.c
    c.alpha, c.eta0, c.eta1, c.ones = 1.0 / k, 1.0, 1.0, links.shape[0]
    st.arr("links")[:] = links
    tl = np.zeros(n)
    np.add.at(tl, links.ravel(), 2.0)
    st.arr("tl")[:] = tl
    st.arr("gamma")[:] = gamma
    st.arr("gammanext")[:] = c.alpha
    st.arr("lambda_")[:] = lam
    st.arr("lambdanext")[:] = lam
    if conv_frac:
        rng = np.random.default_rng(seed + 2)
        who = rng.random(n) < conv_frac
        st.arr("converged")[who] = rng.integers(1, k + 1, who.sum())
    st.refresh_expectations()
    return st, links


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8, 9, 16, 20, 31, 33, 64, 65, 100, 128, 150, 192, 200, 256,
                               257, 300, 500, 1000])
def test_synthetic_three_sweeps_every_tiling(k):
    n = 600 if k <= 256 else 200
    st, links = synthetic_state(n, k, 8 * n, seed=100 + k)
    eng = engine_from_state(st, links.shape[0])
    for it, ann, wc in [(0, 1, 0), (1, 1, 1), (2, 0, 1)]:
        st.step(it, ann, wc)
        eng.step(it, ann, wc)
        compare_sweep(eng, st, "k=%d iter %d" % (k, it), check_member=bool(wc))
    eng.close(); st.free()


@pytest.mark.parametrize("k", [4, 20, 64, 200, 300])
def test_converged_shortcut_and_q4_offbyone(k):
    """A third of the nodes pre-marked converged (incl. community K, whose s3 shortcut reads one past the
    row -- SURVEY.md Q4) so that branches :619-631 and :739-742 carry real weight."""
    st, links = synthetic_state(500, k, 5000, seed=7 + k, conv_frac=0.35)
    if k > 1:
        st.arr("converged")[:10] = k          # force the pc == K corner
    eng = engine_from_state(st, links.shape[0])
    for it, ann, wc in [(0, 0, 1), (1, 1, 0), (2, 0, 1)]:
        st.step(it, ann, wc)
        eng.step(it, ann, wc)
        compare_sweep(eng, st, "k=%d iter %d" % (k, it), check_member=bool(wc))
    assert st.c.cnt_shortcut > 0
    eng.close(); st.free()


@pytest.mark.parametrize("k", [20, 64, 200, 300])
def test_active_set_branch_after_1000_iterations(k):
    """iter > 1000 switches links whose endpoints both have < K/10 active communities to the active-set
    phi (:634-681).  Concentrated gamma rows make that the common case."""
    n = 400
    st, links = synthetic_state(n, k, 4000, seed=31 + k)
    rng = np.random.default_rng(k)
    gam = st.arr("gamma")
    gam[:] = 1.0 / k + 1e-3 * rng.random((n, k))
    for p in range(n):
        hot = rng.choice(k, size=rng.integers(0, max(2, k // 10 + 2)), replace=False)
        gam[p, hot] += 2.0 + 5 * rng.random(hot.size)
    st.refresh_expectations()
    orc.lib().orc_prune(st.ptr)
    st.arr("converged")[:] = 0
    eng = engine_from_state(st, links.shape[0])
    # bring the engine's active masks in line with the pruned state: one ordinary sweep from the same start
    st.step(5, 0, 0); eng.step(5, 0, 0)
    compare_sweep(eng, st, "warm")
    for it, wc in [(1001, 1), (1002, 0), (1003, 1)]:
        st.step(it, 0, wc); eng.step(it, 0, wc)
        compare_sweep(eng, st, "k=%d iter %d" % (k, it), check_member=bool(wc))
    eng.close(); st.free()


def test_isolated_nodes_hubs_and_segment_lengths():
    """Ragged input: nodes without links (tl == 0 keeps the row at alpha, :532-533), one hub touching
    every node (many segments per node), and the same answer for every segment length."""
    n, k = 1500, 20
    rng = np.random.default_rng(5)
    base = synth.mmsb_links(n - 100, k, 6000, seed=9)          # nodes n-100.. are isolated
    hub = np.stack([np.zeros(n - 101, dtype=np.uint32), np.arange(1, n - 100, dtype=np.uint32)], 1)
    links = np.unique(np.concatenate([base, hub]), axis=0).astype(np.uint32)
    gamma = 1.0 / k + rng.random((n, k))
    st = orc.State.alloc(n, k, links.shape[0])
    c = st.c
    c.alpha, c.eta0, c.eta1, c.ones = 1.0 / k, 1.0, 1.0, links.shape[0]
    st.arr("links")[:] = links
    tl = np.zeros(n); np.add.at(tl, links.ravel(), 2.0)
    st.arr("tl")[:] = tl
    st.arr("gamma")[:] = gamma
    st.arr("gammanext")[:] = c.alpha
    st.arr("lambda_")[:] = 1.0; st.arr("lambdanext")[:] = 1.0
    st.refresh_expectations()
    engines = [engine_from_state(st, links.shape[0], seg_len=s) for s in (0, 16, 37, 4096)]
    for it, ann, wc in [(0, 1, 1), (1, 0, 1)]:
        st.step(it, ann, wc)
        for e in engines:
            e.step(it, ann, wc)
            compare_sweep(e, st, "seg variant iter %d" % it, check_member=True)
    g0, _ = engines[0].get_state()
    assert np.all(g0[n - 100:] == 1.0 / k)
    for e in engines:
        e.close()
    st.free()


def test_empty_link_list_and_tiny_graphs():
    for n, links in [(3, np.zeros((0, 2), dtype=np.uint32)), (2, np.array([[0, 1]], dtype=np.uint32))]:
        k = 4
        st = orc.State.alloc(n, k, links.shape[0])
        c = st.c
        c.alpha, c.eta0, c.eta1, c.ones = 0.25, 1.0, 1.0, max(1, links.shape[0])
        st.arr("links")[:] = links
        tl = np.zeros(n); np.add.at(tl, links.ravel().astype(np.int64), 2.0)
        st.arr("tl")[:] = tl
        st.arr("gamma")[:] = 0.25 + np.arange(n * k).reshape(n, k) / 7.0
        st.arr("gammanext")[:] = 0.25
        st.arr("lambda_")[:] = 1.0; st.arr("lambdanext")[:] = 1.0
        st.refresh_expectations()
        eng = engine_from_state(st, max(1, links.shape[0]))
        st.step(0, 0, 1); eng.step(0, 0, 1)
        g, lam = eng.get_state()
        assert rel_err(g, st.arr("gamma")) <= TOL and rel_err(lam, st.arr("lambda_")) <= TOL
        eng.close(); st.free()


def test_run_to_run_determinism():
    """No atomics on floating point anywhere in the path: two handles give bit-identical state."""
    st, links = synthetic_state(3000, 100, 40000, seed=77)
    outs = []
    for _ in range(2):
        eng = engine_from_state(st, links.shape[0])
        for it in range(4):
            eng.step(it, it < 2, 1)
        outs.append(eng.get_state() + (eng.membership_bits(),))
        eng.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])
    st.free()


def test_refresh_before_s3_is_the_same_iteration():
    """svi_ls_phase_refresh / svi_ls_phase_lambda (the order sharded drivers use to hide their exchanges): the
    refresh, prune included, runs BEFORE the s3 sweep, which must still see the pre-prune `converged` flags.
    assort-75-4 is a run where nodes converge sweep after sweep, so a sweep that read the post-prune flags
    would differ; the split order must be bit-identical to svi_ls_step."""
    with Scratch() as d:
        g = orc.Graph.read(input_path("assort-75-4.txt", d), 75)
        m = orc.Model(g, 4, use_validation_stop=0)
        st = m.state
        outs, newly = [], 0
        for split in (False, True):
            eng = engine_from_state(st, g.ones)
            prev = eng.get_converged()[0]
            for it in range(28):
                ann, wc = it < 14, 1
                if split:
                    eng.phase_phi(it, wc); eng.phase_node(); eng.phase_refresh(ann); eng.phase_s3(); eng.phase_lambda(ann)
                else:
                    eng.step(it, ann, wc)
                cur = eng.get_converged()[0]
                newly += int((cur != prev).sum())
                prev = cur
            outs.append(eng.get_state() + eng.get_converged() + (eng.kvectors()["s3"],))
            eng.close()
        for a, b in zip(*outs):
            assert np.array_equal(a, b)
        assert newly > 0          # nodes did converge on the way, i.e. prune changed flags the s3 sweep reads
        m.close(); g.close()


def test_heldout_matches_literal_double_sum():
    st, links = synthetic_state(800, 50, 6000, seed=3)
    st.step(0, 1, 0)
    eng = engine_from_state(st, links.shape[0])
    rng = np.random.default_rng(0)
    p = rng.integers(0, 800, 500).astype(np.uint32)
    q = (p + 1 + rng.integers(0, 798, 500).astype(np.uint32)) % 800
    y = rng.integers(0, 2, 500).astype(np.uint8)
    got = eng.heldout(p, q, y)
    want = np.array([st.edge_likelihood(int(a), int(b), int(c)) for a, b, c in zip(p, q, y)])
    assert rel_err(got, want, floor=1e-3) <= TOL
    eng.close(); st.free()


# ---- BASELINE sizes: size-independent properties (the oracle cannot finish these in seconds) ----
@pytest.mark.parametrize("n,k,nlinks", [(100000, 100, 5000000)])
def test_properties_at_config3_size(n, k, nlinks):
    import torch
    links = synth.mmsb_links(n, k, nlinks, seed=1234, device="cuda")
    deg = np.bincount(links.ravel().astype(np.int64), minlength=n).astype(np.float64)
    rng = np.random.default_rng(1)
    gamma = (deg[:, None] / k) * (1.0 + 0.2 * rng.random((n, k))) + 1.0 / k
    lam = np.ones((k, 2))
    eng = LinkSamplingEngine(n, k, links)
    eng.set_state(gamma, lam)
    eng.step(0, 0, 1)                                  # no annealing: closed-form row sums exist
    g, lam1 = eng.get_state()
    kv = eng.kvectors()
    # every link spreads exactly one unit of phi per endpoint: sum_k sum[k] == 2 * nlinks
    assert abs(kv["sum"].sum() - 2.0 * links.shape[0]) <= 1e-9 * 2.0 * links.shape[0]
    assert rel_err(lam1[:, 0], 1.0 + kv["sum"]) <= 1e-12
    assert rel_err(lam1[:, 1], 1.0 + kv["s1"] ** 2 - kv["s2"] - kv["s3"]) <= 1e-9
    # gamma row sums: K*alpha + deg + (n - tl - 1) * deg / tl  with tl = 2 deg  (:536-539)
    has = deg > 0
    want = 1.0 + deg[has] + (n - 2 * deg[has] - 1) * 0.5
    assert rel_err(g[has].sum(1), want) <= 1e-10
    assert np.all(g[~has] == 1.0 / k)
    # s1 sums mphi over nodes: each node with links carries total mass deg/tl = 1/2
    assert abs(kv["s1"].sum() - 0.5 * has.sum()) <= 1e-9 * has.sum()
    # membership: every node with a link belongs to at least one link community
    mb = eng.membership_bits()
    assert np.all((mb != 0).any(1) == has)
    eng.close()
    torch.cuda.empty_cache()
