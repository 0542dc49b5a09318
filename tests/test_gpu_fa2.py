"""Parity of the `-rnode -stratified` device path (include/svi_fa2.h) against the FastAMM2 oracle -- needs a B200.

The oracle (oracle/oracle_fa2.c) is pinned byte-for-byte to the compiled reference (test_oracle_fa2_golden.py).
Here the engine is driven with the oracle's own minibatch sequence (the reference's mt19937 draws) and compared
iteration by iteration.  Official bar: gamma/lambda within 1e-5 relative; asserted at 1e-9 (observed ~1e-13).
"""
import numpy as np
import pytest

import oracle_py as orc
from golden_util import MANIFEST, Scratch, input_path
from philox_ref import philox4x32_10
from svinet_b200.fa2_engine import Fa2Engine
from test_oracle_fa2_golden import fa2_opts

pytestmark = pytest.mark.gpu

TOL = 1e-9
TOL_OFFICIAL = 1e-5


def rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def psi_rows(g):
    L = orc.lib()
    f = np.vectorize(L.orc_digamma)
    return f(g) - f(g.sum(axis=1, keepdims=True))


@pytest.mark.parametrize("k", [1, 2, 3, 4, 7, 16, 20, 28, 33, 64, 100, 129, 200, 256, 300, 512])
def test_phi_pair_matches_oracle(k):
    """One pair's coordinate ascent for every kernel tiling, both link and non-link form."""
    rng = np.random.default_rng(k)
    n = 6
    gamma = rng.gamma(1.0, 1.0, size=(n, k)) + 1e-3
    gamma[1] = 1.0 / k                       # a node still at its prior
    gamma[2, rng.integers(k)] += 40.0        # a peaked node
    lam = np.stack([1 + rng.gamma(2.0, 1.0, k), 1 + rng.gamma(5.0, 3.0, k)], axis=1)
    eng = Fa2Engine(n, k)
    eng.set_state(gamma, lam)
    epi = psi_rows(gamma)
    ebeta = psi_rows(lam)
    worst = 0.0
    for (p, q) in [(0, 1), (1, 2), (2, 3), (0, 5), (3, 4)]:
        for y in (0, 1):
            want1, want2, rounds = orc.fa2_phi_pair(epi[p], epi[q], ebeta[:, 0] if y else ebeta[:, 1], y)
            got1, got2, r = eng.phi_pair(p, q, y)
            assert r == rounds, (p, q, y, r, rounds)
            worst = max(worst, float(np.max(np.abs(got1 - want1))), float(np.max(np.abs(got2 - want2))))
    assert worst <= 1e-12, worst
    eng.close()


def lockstep(case, iters, eager=0):
    ent = MANIFEST[case]
    opts = fa2_opts(ent["flags"])
    with Scratch() as d:
        g = orc.Graph.read(input_path(ent["input"], d), ent["n"])
        m = orc.Fa2Model(g, ent["k"], **opts)
        eng = Fa2Engine(m.n, m.k, eager_blend=eager)
        eng.set_state(m.gamma, m.lambda_)
        worst_g = worst_l = 0.0
        types = [0, 0]
        for _ in range(iters):
            it = m.iter
            typ, start, pairs = m.plan()
            m.process()
            eng.step(it, typ, start, pairs)
            types[typ] += 1
            gam, lam = eng.get_state()
            worst_g = max(worst_g, rel_err(gam, m.gamma))
            worst_l = max(worst_l, rel_err(lam, m.lambda_))
            assert worst_g <= TOL and worst_l <= TOL, (case, it, typ, start, len(pairs), worst_g, worst_l)
        # held-out likelihood under the final state, FastAMM2::edge_likelihood
        hp = m.heldout_pairs()
        if len(hp):
            y = np.array([g.y(int(a), int(b)) for a, b in hp], dtype=np.uint8)
            got = eng.heldout(hp[:, 0], hp[:, 1], y)
            want = np.array([m.edge_likelihood(int(a), int(b), int(yy)) for (a, b), yy in zip(hp, y)])
            assert rel_err(got, want, floor=1e-3) <= TOL
        eng.close(); m.close(); g.close()
    assert types[0] > 0 and types[1] > 0       # both samplers were exercised
    return worst_g, worst_l


@pytest.mark.parametrize("eager", [0, 1], ids=["lazy", "eager"])
@pytest.mark.parametrize("case,iters", [("fa2_c1_m200", 201), ("fa2_c1_k6_seed9_m500", 300), ("fa2_lfr_k28_m300", 120),
                                        ("fa2_c2_m120", 40)])
def test_lockstep_with_reference_minibatches(case, iters, eager):
    """Both treatments of the untouched rows' decay (svi_fa2_config.eager_blend): the default scalar form and
    the reference's explicit pass over all N rows."""
    wg, wl = lockstep(case, iters, eager)
    assert wg <= TOL_OFFICIAL and wl <= TOL_OFFICIAL


def test_lazy_decay_equals_eager_blend_over_a_rebase():
    """25 000 device-drawn iterations: the scalar decay factor falls below e^-230 and the stored rows are
    re-based (k_fa2_fold) at least once; the state must still equal the eager pass's."""
    n, k = 120, 5
    links, gamma, lam, heldout, shuffled = _synthetic(n, k, 6, seed=3)
    seed = 99
    out = []
    for eager in (0, 1):
        e = Fa2Engine(n, k, eager_blend=eager)
        e.set_state(gamma, lam)
        e.set_graph(links, heldout, shuffled)
        e.run(0, 25000, seed, count=False)
        out.append(e.get_state())
        assert e.info()["kernels_per_step"] == (6 if eager == 0 else 7)
        e.close()
    (gl, ll), (ge, le) = out
    assert np.all(np.isfinite(gl)) and np.all(gl > 0)
    assert rel_err(gl, ge) <= 1e-9 and rel_err(ll, le) <= 1e-9


def test_free_running_matches_reference_fixture():
    """All 201 iterations of fa2_c1_m200 with no per-iteration resynchronisation, then compare with the
    REFERENCE's gamma.txt / lambda.txt at print precision."""
    from golden_util import golden_text
    case = "fa2_c1_m200"
    ent = MANIFEST[case]
    with Scratch() as d:
        g = orc.Graph.read(input_path(ent["input"], d), ent["n"])
        m = orc.Fa2Model(g, ent["k"], **fa2_opts(ent["flags"]))
        eng = Fa2Engine(m.n, m.k)
        eng.set_state(m.gamma, m.lambda_)
        for _ in range(201):
            it = m.iter
            typ, start, pairs = m.plan()
            m.process()                       # only to advance the oracle's RNG/plan; its state is not fed back
            eng.step(it, typ, start, pairs)
        gam, lam = eng.get_state()
        want_g = np.array([[float(x) for x in l.split("\t")[2:]] for l in golden_text(case, "gamma.txt").strip().split("\n")])
        want_l = np.array([[float(x) for x in l.split("\t")[1:]] for l in golden_text(case, "lambda.txt").strip().split("\n")])
        assert np.max(np.abs(gam - want_g)) <= 0.6e-5 and np.max(np.abs(lam - want_l)) <= 0.6e-5
        eng.close(); m.close(); g.close()


def _synthetic(n, k, deg, seed):
    rng = np.random.default_rng(seed)
    m = n * deg // 2
    a = rng.integers(0, n, size=m, dtype=np.int64)
    b = (a + 1 + rng.integers(0, n - 1, size=m, dtype=np.int64)) % n
    e = np.unique(np.stack([np.minimum(a, b), np.maximum(a, b)], axis=1), axis=0).astype(np.uint32)
    gamma = rng.gamma(1.0, 1.0, size=(n, k)) + 0.01
    lam = np.stack([1 + rng.gamma(2.0, 1.0, k), 1 + rng.gamma(5.0, 3.0, k)], axis=1)
    heldout = e[rng.choice(len(e), size=max(2, len(e) // 100), replace=False)]
    extra = np.array([[0, n - 1], [1, n - 2]], dtype=np.uint32)      # held-out non-links
    shuffled = rng.permutation(n).astype(np.uint32)
    return e, gamma, lam, np.concatenate([heldout, extra]), shuffled


def host_draw(n, m_sets, inf_eps, links, heldout, shuffled, it, seed):
    """Replay of k_fa2_draw on the host (the sampling rules of fastamm2.cc:574,936,943-960,1095-1125)."""
    r = philox4x32_10([it, 0, 0, 0], [seed & 0xffffffff, seed >> 32])
    typ = 1 if r[0] / 4294967296.0 < inf_eps else 0
    start = (r[1] * n) >> 32
    ho = {(int(a), int(b)) for a, b in heldout}
    nb = set(links[links[:, 0] == start][:, 1].tolist()) | set(links[links[:, 1] == start][:, 0].tolist())
    if typ == 0:
        out = [(min(start, a), max(start, a)) for a in sorted(nb) if (min(start, a), max(start, a)) not in ho]
        return typ, start, np.array(out, dtype=np.uint32).reshape(-1, 2), len(nb)
    setsize = int(n / m_sets)
    q0 = (((r[2] * n) >> 32) // setsize) * setsize
    out = []
    for c in range(n):
        node = int(shuffled[(q0 + c) % n])
        if node == start or node in nb or (min(start, node), max(start, node)) in ho:
            continue
        out.append((min(start, node), max(start, node)))
        if len(out) == setsize:
            break
    return typ, start, np.array(out, dtype=np.uint32).reshape(-1, 2), len(out)


@pytest.mark.parametrize("n,k,deg", [(300, 12, 8), (5000, 40, 30)])
def test_device_draw_and_run(n, k, deg):
    """svi_fa2_draw reproduces the documented Philox sampling rules; svi_fa2_run == draw + step."""
    links, gamma, lam, heldout, shuffled = _synthetic(n, k, deg, seed=n)
    seed = 0x1234_5678_9abc_def0
    a = Fa2Engine(n, k)
    b = Fa2Engine(n, k)
    for e in (a, b):
        e.set_state(gamma, lam)
        e.set_graph(links, heldout, shuffled)
    sampled = 0
    seen = [0, 0]
    for it in range(12):
        typ, start, pairs = a.draw(it, seed)
        ht, hs, hpairs, inc = host_draw(n, 10, 0.5, links, heldout, shuffled, it, seed)
        assert (typ, start) == (ht, hs) and np.array_equal(pairs, hpairs), (it, typ, start)
        seen[typ] += 1
        sampled += inc
        a.step(it, typ, start, pairs)
    got = b.run(0, 12, seed)
    assert got == sampled and min(seen) > 0
    ga, la = a.get_state()
    gb, lb = b.get_state()
    assert rel_err(gb, ga) <= 1e-12 and rel_err(lb, la) <= 1e-12     # device pow vs libm pow in rho
    # run-to-run determinism of the device path
    c = Fa2Engine(n, k)
    c.set_state(gamma, lam)
    c.set_graph(links, heldout, shuffled)
    c.run(0, 12, seed, count=False)
    gc, lc = c.get_state()
    assert np.array_equal(gc, gb) and np.array_equal(lc, lb)
    for e in (a, b, c):
        e.close()


def test_untouched_rows_decay_and_isolated_start():
    """A start node without pairs: every row decays towards alpha, lambda towards eta (fastamm2.cc:614-620)."""
    n, k = 40, 5
    rng = np.random.default_rng(0)
    gamma = rng.gamma(2.0, 1.0, size=(n, k))
    lam = 1 + rng.gamma(2.0, 1.0, size=(k, 2))
    eng = Fa2Engine(n, k)
    eng.set_state(gamma, lam)
    eng.step(0, 0, 7, np.zeros((0, 2), dtype=np.uint32))
    g1, l1 = eng.get_state()
    rho = (1025.0 + 0.0) ** -0.5
    rho_t = (1025.0 + 1.0) ** -0.9
    assert rel_err(g1, (1 - rho) * gamma + rho * (1.0 / k)) <= 1e-14
    assert rel_err(l1, (1 - rho_t) * lam + rho_t * 1.0) <= 1e-14
    eng.close()


def test_step_rejects_bad_minibatch():
    from svinet_b200.engine import SviError
    eng = Fa2Engine(10, 4)
    eng.set_state(np.ones((10, 4)), np.ones((4, 2)))
    with pytest.raises(SviError):
        eng.step(0, 0, 3, np.array([[1, 2]], dtype=np.uint32))      # does not contain the start node
    with pytest.raises(SviError):
        eng.step(0, 2, 3, np.array([[3, 4]], dtype=np.uint32))      # bad type
    eng.close()
