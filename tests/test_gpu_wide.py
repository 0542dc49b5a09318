"""K > 1024 on the device: the block-per-row kernels of svinet_b200/csrc/svi_ls_wide.cuh through the C ABI against the
oracle -- needs a B200.  (The reference's limit is K <= 65 535: communities are uint16_t, src/env.hh:37.)

The same kernel source is checked against the oracle on host threads, and its barriers under ThreadSanitizer, in
tests/test_wide_emulated.py; here the device build, its dispatch in svi_ls.cu (grids, partial-sum slots, the CUDA graph
of svi_ls_step) and the sharded path run.  Cases mirror tests/test_gpu_parity.py.  No torch import on the way in: the
file also serves as a quick stand-alone check (`pytest tests/test_gpu_wide.py`).
"""
import ctypes as C

import numpy as np
import pytest

import oracle_py as orc
from parity_util import TOL, compare_sweep, engine_from_state, rel_err
from svinet_b200 import engine
from test_wide_emulated import make_state, random_links

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [1025, 1536, 2053, 4100])
def test_three_sweeps_every_wide_width(k):
    n = 150
    rng = np.random.default_rng(k)
    links = random_links(n, 6 * n, rng, hub=True)
    st = make_state(n, k, links, seed=k + 1)
    eng = engine_from_state(st, links.shape[0])
    info = eng.info()
    assert info["lanes"] == 256 and 2 * info["lanes"] * info["vec"] >= info["ld"]
    for it, ann, wc in [(0, 1, 0), (1, 1, 1), (2, 0, 1)]:
        st.step(it, ann, wc)
        eng.step(it, ann, wc)
        compare_sweep(eng, st, "k=%d iter %d" % (k, it), check_member=bool(wc))
    eng.close(); st.free()


def test_converged_shortcut_q4_and_isolated_nodes():
    n, k = 120, 1100
    rng = np.random.default_rng(3)
    links = random_links(n - 10, 700, rng)                   # nodes n-10.. are isolated (tl == 0, :532-533)
    st = make_state(n, k, links, seed=11, conv_frac=0.35)
    st.arr("converged")[:6] = k                              # the pc == K corner
    eng = engine_from_state(st, links.shape[0], seg_len=7)
    for it, ann, wc in [(0, 1, 1), (1, 1, 0), (2, 0, 1)]:
        st.step(it, ann, wc)
        eng.step(it, ann, wc)
        assert np.isfinite(st.arr("gamma")).all()
        compare_sweep(eng, st, "iter %d" % it, check_member=bool(wc))
    assert st.c.cnt_shortcut > 0 and st.c.cnt_dense > 0
    eng.close(); st.free()


def test_active_set_branch_and_newly_converged_nodes():
    n, k = 120, 1040
    rng = np.random.default_rng(5)
    links = random_links(n, 600, rng)
    st = make_state(n, k, links, seed=17)
    gam = st.arr("gamma")
    gam[:] = 1.0 / k + 1e-3 * rng.random((n, k))
    for p in range(n):
        hot = rng.choice(k, size=int(rng.integers(1, 8)) if p % 3 else 150, replace=False)
        gam[p, hot] += 2.0 + 5 * rng.random(hot.size)
    st.refresh_expectations()
    orc.lib().orc_prune(st.ptr)
    st.arr("converged")[:] = 0
    eng = engine_from_state(st, links.shape[0])
    st.step(5, 0, 0); eng.step(5, 0, 0)
    compare_sweep(eng, st, "warm")
    sparse = 0
    for it, wc in [(1001, 1), (1002, 0), (1003, 1)]:
        st.step(it, 0, wc); eng.step(it, 0, wc)
        compare_sweep(eng, st, "iter %d" % it, check_member=bool(wc))
        sparse += st.c.cnt_sparse
    assert sparse > 0
    eng.close(); st.free()


def test_empty_link_list_single_link_and_determinism():
    for n, links in [(3, np.zeros((0, 2), dtype=np.uint32)), (2, np.array([[0, 1]], dtype=np.uint32))]:
        k = 1030
        st = make_state(n, k, links, seed=1)
        st.c.ones = max(1, links.shape[0])
        st.arr("gamma")[:] = 1.0 / k + np.arange(n * k).reshape(n, k) / 7.0
        st.refresh_expectations()
        eng = engine_from_state(st, max(1, links.shape[0]))
        st.step(0, 0, 1); eng.step(0, 0, 1)
        compare_sweep(eng, st, "n=%d" % n, check_member=True)
        eng.close(); st.free()
    # no floating-point atomics: two handles end bit-identical
    n, k = 100, 1200
    links = random_links(n, 500, np.random.default_rng(4), hub=True)
    st = make_state(n, k, links, seed=9)
    outs = []
    for _ in range(2):
        eng = engine_from_state(st, links.shape[0])
        for it in range(3):
            eng.step(it, it < 2, 1)
        outs.append(eng.get_state() + (eng.membership_bits(),))
        eng.close()
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    st.free()


def test_heldout_matches_the_literal_double_sum():
    n, k = 30, 1027
    rng = np.random.default_rng(8)
    links = random_links(n, 90, rng)
    st = make_state(n, k, links, seed=21)
    eng = engine_from_state(st, links.shape[0])
    p = rng.integers(0, n, 40).astype(np.uint32)
    q = ((p + 1 + rng.integers(0, n - 1, 40)) % n).astype(np.uint32)
    y = rng.integers(0, 2, 40).astype(np.uint8)
    want = np.array([st.edge_likelihood(int(a), int(b), int(yy)) for a, b, yy in zip(p, q, y)])
    assert rel_err(eng.heldout(p, q, y), want, floor=1e-3) <= TOL
    q[7] = n
    with pytest.raises(engine.SviError, match="pair 7 out of range"):
        eng.heldout(p, q, y)
    eng.close(); st.free()


def test_the_reference_limit_k_65535_and_beyond():
    """K = 65 535 (the largest community id a uint16_t holds) runs; K = 65 536 is refused."""
    n, k = 6, 65535
    links = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [0, 4], [1, 4]], dtype=np.uint32)
    st = make_state(n, k, links, seed=2)
    eng = engine_from_state(st, links.shape[0])
    # (no annealing sweep: with 65 535 communities and six links most phi underflow to exactly 0, whole column sums are
    # 0 and the reference's rescale ones/sum[k] (:541-542) is inf -- in the oracle too)
    for it, ann, wc in [(0, 0, 1), (1, 0, 1)]:
        st.step(it, ann, wc)
        eng.step(it, ann, wc)
        assert np.isfinite(st.arr("gamma")).all()
        compare_sweep(eng, st, "k=65535 iter %d" % it, check_member=True)
    eng.close(); st.free()
    L = engine.load_library()
    cfg = engine.SviConfig(n=n, k=65536, nlinks=0, alpha=1.0 / 65536, eta0=1, eta1=1, ones=0, device=-1, seg_len=0,
                           node_begin=0, node_end=n)
    h = C.c_void_p()
    assert L.svi_ls_create(C.byref(cfg), None, None, C.byref(h)) == -4


def loglik_long_double(st, hp, hq, hy, epsilon=1e-30):
    """LinkSampling::edge_likelihood (src/linksampling.hh:259-292) in extended precision: the non-link double sum
    sum_zp sum_zq pi_p[zp] pi_q[zq] (1 - rate(zp, zq)) with rate = beta_z on the diagonal, epsilon off it"""
    G, lam = st.arr("gamma").astype(np.longdouble), st.arr("lambda_").astype(np.longdouble)
    rate, eps = lam[:, 0] / (lam[:, 0] + lam[:, 1]), np.longdouble(epsilon)
    out = []
    for a, b, y in zip(hp, hq, hy):
        pa, pb = G[a] / G[a].sum(), G[b] / G[b].sum()
        s = (pa * pb * rate).sum() if y else (pa * (pb * (1 - rate) + (pb.sum() - pb) * (1 - eps))).sum()
        out.append(float(np.log(max(s, np.longdouble(1e-30)))))
    return np.array(out)


def sharded_vs_oracle(n, k, links, gamma, conv, bounds, chunks, sched, seg_len, share):
    """tests/test_gpu_mg.py::run_sharded_vs_oracle in small (no torch on the way in): all shards in this process on one
    device, svi_ls_mg_step sweep by sweep against the oracle; returns the oracle's shortcut count of the last sweep"""
    from svinet_b200.engine import LinkSamplingEngine
    st = make_state(n, k, links, seed=1)
    st.arr("gamma")[:] = gamma
    st.arr("lambda_")[:] = 1.0
    st.arr("converged")[:] = conv
    st.refresh_expectations()
    world = len(bounds) - 1
    engines = [LinkSamplingEngine(n, k, links, node_range=(int(bounds[r]), int(bounds[r + 1])), seg_len=seg_len)
               for r in range(world)]
    LinkSamplingEngine.attach_local(engines, np.asarray(bounds, dtype=np.uint32), chunks=chunks)
    rng = np.random.default_rng(7)
    hp = rng.integers(0, n, 24).astype(np.uint32)
    hq = ((hp + 1 + rng.integers(0, n - 1, 24)) % n).astype(np.uint32)
    hy = rng.integers(0, 2, 24).astype(np.uint8)
    for e in engines:
        if share:
            e.mg_share_gamma(True)
        e.set_state(st.arr("gamma"), st.arr("lambda_"))
        e.set_converged(st.arr("converged"))
    for it, ann, wc in sched:
        st.step(it, ann, wc)
        for e in engines:
            e.mg_step(it, ann, wc)
        for e in engines:
            e.sync()
        # The oracle's literal non-link form adds K^2 = 1.2e6 terms one by one in FP64: once gamma has concentrated that
        # sum is only good to ~1e-8 of log(s) (first hardware run: 1.06e-8 on sweep 3, the device agreeing with extended
        # precision to 1e-13).  So: the oracle at the official tolerance, the same formula in long double at 1e-9.
        want_ll = np.array([st.edge_likelihood(int(a), int(b), int(c)) for a, b, c in zip(hp, hq, hy)])
        exact_ll = loglik_long_double(st, hp, hq, hy)
        for e in engines:   # rows of other shards: replicated (share) or peer loads from the owner's arena
            got = e.heldout(hp, hq, hy)
            assert rel_err(got, exact_ll, floor=1e-3) <= TOL
            assert rel_err(got, want_ll, floor=1e-3) <= 1e-5
        if not share:
            for e in engines:
                e.mg_publish_gamma()
        mem = np.zeros((n, k), dtype=np.uint8)
        for r, e in enumerate(engines):
            tag = "world=%d shard %d iter %d" % (world, r, it)
            g, lam = e.get_state()
            assert rel_err(g, st.arr("gamma")) <= TOL and rel_err(lam, st.arr("lambda_")) <= TOL, tag
            kv = e.kvectors()
            for name in ("sum", "s1", "s2", "s3"):
                want = st.arr(name)
                assert rel_err(kv[name], want, floor=max(1e-12, 1e-6 * float(np.max(np.abs(want))))) <= TOL, (tag, name)
            cv, act = e.get_converged()
            nb, ne = int(bounds[r]), int(bounds[r + 1])
            assert np.array_equal(cv, st.arr("converged")) and np.array_equal(act[nb:ne], st.arr("active_comms")[nb:ne]), tag
            if wc:
                bits = e.membership_rows(nb, ne - nb)
                cols = np.arange(k)
                mem[nb:ne] = (bits[:, cols // 32] >> (cols % 32).astype(np.uint32)) & 1
        if wc:
            assert np.array_equal(mem, st.arr("member")), "membership, iter %d" % it
    shortcut = st.c.cnt_shortcut
    for e in engines:
        e.close()
    st.free()
    return shortcut


def test_two_and_three_shards_on_one_gpu():
    """svi_ls_mg_step with the wide tile: block-restricted launches, chunked node passes into the partial-sum slots,
    the membership merge, peer-row held-out loads -- against the oracle."""
    n, k = 160, 1100
    rng = np.random.default_rng(12)
    links = random_links(n, 10 * n, rng, hub=True)
    st = make_state(n, k, links, seed=5)
    g0 = st.arr("gamma").copy()
    st.free()
    conv = np.zeros(n, dtype=np.uint32)
    who = rng.random(n) < 0.25
    conv[who] = rng.integers(1, k + 1, who.sum())
    sched = [(0, 1, 0), (1, 1, 1), (2, 0, 1), (3, 0, 0)]
    assert sharded_vs_oracle(n, k, links, g0, conv, [0, 70, n], 3, sched, seg_len=16, share=False) > 0
    assert sharded_vs_oracle(n, k, links, g0, conv, [0, 50, 111, n], 1, sched, seg_len=16, share=True) > 0


def test_cli_output_directory_at_k_1100():
    """The drop-in CLI end to end with 1100 communities: `svinet -link-sampling -k 1100 -max-iterations 8` against the
    files the oracle's writers produce for the same run (the oracle's writers are pinned to the reference's bytes on the
    golden fixtures, tests/test_oracle_golden.py): numeric files within one unit of the last printed digit,
    communities.txt byte for byte."""
    import os
    import subprocess
    from golden_util import Scratch, compare_numeric_text, input_path
    from svinet_b200 import build as svbuild
    cli = svbuild.build_cli()
    k = 1100
    with Scratch() as d:
        inp = input_path("assort-75-4.txt", d)
        if not os.path.exists(os.path.join(d, "assort-75-4.txt")):
            os.symlink(inp, os.path.join(d, "assort-75-4.txt"))
        p = subprocess.run([cli, "-file", "assort-75-4.txt", "-n", "75", "-k", str(k), "-link-sampling", "-max-iterations", "8"],
                           cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=600)
        assert p.returncode == 0, p.stderr.decode()
        out = os.path.join(d, "n75-k%d-mmsb-linksampling" % k)
        g = orc.Graph.read(inp, 75)
        m = orc.Model(g, k, max_iterations=8)
        m.run()
        want = os.path.join(d, "want")
        m.write_outputs(want)
        m.close(); g.close()
        flips = {}
        for fname in ("gamma.txt", "lambda.txt", "groups.txt", "validation.txt", "max.txt"):
            flips[fname] = compare_numeric_text(open(os.path.join(out, fname)).read(), open(os.path.join(want, fname)).read(),
                                                skip_cols=(1,) if fname in ("validation.txt", "max.txt") else ())
        for fname in ("communities.txt", "validation-edges.txt"):
            assert open(os.path.join(out, fname)).read() == open(os.path.join(want, fname)).read(), fname
        nf, noff = flips["gamma.txt"]
        assert noff <= max(2, nf // 1000), flips
