"""The K > 1024 kernels (svinet_b200/csrc/svi_ls_wide.cuh) against the oracle WITHOUT a GPU.

The kernels are written with nothing but threadIdx/blockIdx, __syncthreads and block-shared arrays, so the very same
source compiles as host code (tests/cc/cuda_shim/) and runs with one host thread per CUDA thread
(tests/cc/wide_emul.cc, which also restates the host-side launch order of one iteration).  These tests drive that
build exactly like tests/test_gpu_parity.py drives the device: lockstep with the oracle, gamma / lambda / K-vectors to
1e-9 relative, converged / active_comms / link-community membership exactly -- and once more under ThreadSanitizer,
which checks the barrier placement (every cross-thread access to a block-shared or global location must be ordered
by a __syncthreads).  The device build of the same kernels is covered by tests/test_gpu_parity.py (K > 1024 cases).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_py as orc
import wide_emul_py as we
from parity_util import compare_sweep, rel_err

FAST_T = 32      # emulated block size of the broad cases (a barrier of 256 host threads on a few cores is slow)


def random_links(n, nlinks, rng, hub=False):
    pairs = set()
    while len(pairs) < nlinks:
        a, b = int(rng.integers(0, n)), int(rng.integers(0, n))
        if a != b:
            pairs.add((min(a, b), max(a, b)))
    if hub:
        pairs |= {(0, v) for v in range(1, n - 3)}          # node 0 touches almost everybody: several segments
    return np.array(sorted(pairs), dtype=np.uint32)


def make_state(n, k, links, seed, conv_frac=0.0, isolated=0):
    rng = np.random.default_rng(seed)
    st = orc.State.alloc(n, k, links.shape[0])
    c = st.c
    c.alpha, c.eta0, c.eta1, c.ones = 1.0 / k, 1.0, 1.0, links.shape[0]
    st.arr("links")[:] = links
    tl = np.zeros(n)
    np.add.at(tl, links.ravel(), 2.0)
    st.arr("tl")[:] = tl
    gamma = np.zeros((n, k))
    for p, q in links:                                       # the shape init_gamma2 gives (:374-401)
        phi = rng.random(k)
        phi /= phi.sum()
        gamma[p] += phi
        gamma[q] += phi
    gamma[gamma.sum(1) == 0] = 1.0 / k
    st.arr("gamma")[:] = gamma
    st.arr("gammanext")[:] = c.alpha
    st.arr("lambda_")[:] = 1.0 + rng.random((k, 2))
    st.arr("lambdanext")[:] = np.array([c.eta0, c.eta1])
    if conv_frac:
        who = rng.random(n) < conv_frac
        st.arr("converged")[who] = rng.integers(1, k + 1, who.sum())
    st.refresh_expectations()
    return st


def engine_for(st, links, threads, **kw):
    c = st.c
    eng = we.WideEmulEngine(c.n, c.k, links, st.arr("tl"), alpha=c.alpha, eta0=c.eta0, eta1=c.eta1, ones=links.shape[0],
                            threads=threads, **kw)
    eng.set_state(st.arr("gamma"), st.arr("lambda_"))
    eng.set_converged(st.arr("converged"))
    return eng


@pytest.mark.parametrize("k,threads", [(1025, FAST_T), (1500, FAST_T), (2100, FAST_T), (1030, 256)])
def test_three_sweeps_against_oracle(k, threads):
    n = 36 if threads == FAST_T else 20
    rng = np.random.default_rng(k)
    links = random_links(n, 4 * n, rng, hub=True)            # 2 isolated-from-the-hub nodes, seg_len 16 -> the hub has 3+ segments
    st = make_state(n, k, links, seed=k + 1)
    eng = engine_for(st, links, threads, seg_len=16)
    info = eng.info()
    assert info["nseg"] > n and info["blocks_node"] > 1 and info["blocks_s3"] > 1
    for it, ann, wc in [(0, 1, 0), (1, 1, 1), (2, 0, 1)]:
        st.step(it, ann, wc)
        eng.step(it, ann, wc)
        compare_sweep(eng, st, "k=%d iter %d" % (k, it), check_member=bool(wc))
    eng.close(); st.free()


def test_converged_shortcut_q4_and_isolated_nodes():
    """A third of the nodes pre-marked converged, some on community K (the s3 shortcut then reads one past the row,
    SURVEY.md Q4), plus nodes without links (tl == 0: the row stays at alpha, :532-533)."""
    n, k = 40, 1100
    rng = np.random.default_rng(3)
    links = random_links(n - 4, 150, rng)                    # nodes n-4.. are isolated
    st = make_state(n, k, links, seed=11, conv_frac=0.35)
    st.arr("converged")[:4] = k
    eng = engine_for(st, links, FAST_T, seg_len=7)
    for it, ann, wc in [(0, 1, 1), (1, 1, 0), (2, 0, 1)]:
        st.step(it, ann, wc)
        eng.step(it, ann, wc)
        assert np.isfinite(st.arr("gamma")).all()             # (a column sum of 0 under annealing would make both sides inf)
        compare_sweep(eng, st, "iter %d" % it, check_member=bool(wc))
    assert st.c.cnt_shortcut > 0 and st.c.cnt_dense > 0
    eng.close(); st.free()


def test_active_set_branch_and_newly_converged_nodes():
    """iter > 1000: links whose endpoints both have < K/10 active communities take the active-set phi (:634-681);
    rows with exactly one active community converge on the way (prune, :456-491)."""
    n, k = 40, 1040
    rng = np.random.default_rng(5)
    links = random_links(n, 160, rng)
    st = make_state(n, k, links, seed=17)
    gam = st.arr("gamma")
    gam[:] = 1.0 / k + 1e-3 * rng.random((n, k))
    for p in range(n):
        hot = rng.choice(k, size=int(rng.integers(1, 8)) if p % 3 else 150, replace=False)   # every third node: > K/10
        gam[p, hot] += 2.0 + 5 * rng.random(hot.size)
    st.refresh_expectations()
    orc.lib().orc_prune(st.ptr)
    st.arr("converged")[:] = 0
    eng = engine_for(st, links, FAST_T, seg_len=16)
    st.step(5, 0, 0); eng.step(5, 0, 0)                      # brings the engine's active masks in line with the state
    compare_sweep(eng, st, "warm")
    sparse = 0
    for it, wc in [(1001, 1), (1002, 0), (1003, 1)]:
        st.step(it, 0, wc); eng.step(it, 0, wc)
        compare_sweep(eng, st, "iter %d" % it, check_member=bool(wc))
        sparse += st.c.cnt_sparse
    assert sparse > 0
    eng.close(); st.free()


def test_empty_link_list_and_a_single_link():
    for n, links in [(3, np.zeros((0, 2), dtype=np.uint32)), (2, np.array([[0, 1]], dtype=np.uint32))]:
        k = 1030
        st = make_state(n, k, links, seed=1)
        st.c.ones = max(1, links.shape[0])
        st.arr("gamma")[:] = 1.0 / k + np.arange(n * k).reshape(n, k) / 7.0
        st.refresh_expectations()
        eng = engine_for(st, links, FAST_T)
        st.step(0, 0, 1); eng.step(0, 0, 1)
        compare_sweep(eng, st, "n=%d" % n, check_member=True)
        eng.close(); st.free()


def test_heldout_matches_the_literal_double_sum():
    n, k = 12, 1027
    rng = np.random.default_rng(8)
    links = random_links(n, 30, rng)
    st = make_state(n, k, links, seed=21)
    eng = engine_for(st, links, FAST_T)
    p = np.array([0, 1, 2, 3, 4, 5, 6], dtype=np.uint32)
    q = np.array([7, 8, 9, 10, 11, 0, 1], dtype=np.uint32)
    y = np.array([1, 0, 1, 0, 1, 0, 0], dtype=np.uint8)
    got, bad = eng.heldout(p, q, y, blocks=3)
    want = np.array([st.edge_likelihood(int(a), int(b), int(yy)) for a, b, yy in zip(p, q, y)])
    assert bad == 0 and rel_err(got, want, floor=1e-3) <= 1e-9
    q[4] = n                                                  # an out-of-range pair is reported, not read
    got, bad = eng.heldout(p, q, y, blocks=2)
    assert bad == 5 and np.isnan(got[4]) and rel_err(np.delete(got, 4), np.delete(want, 4), floor=1e-3) <= 1e-9
    eng.close(); st.free()


def test_barrier_placement_under_thread_sanitizer():
    """The same sweeps in a ThreadSanitizer build: a missing __syncthreads shows up as a data race."""
    tc = we.tsan_toolchain()
    if tc is None:
        pytest.skip("no g++ with libtsan here")
    lib = we.build(threads=16, tsan=True)
    code = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
import oracle_py as orc, wide_emul_py as we, test_wide_emulated as t
from parity_util import compare_sweep
we.build = lambda threads=256, tsan=False: %r          # the ThreadSanitizer build stands in for the plain one
n, k = 14, 1025
rng = np.random.default_rng(2)
links = t.random_links(n, 30, rng)
st = t.make_state(n, k, links, seed=4, conv_frac=0.2)
eng = t.engine_for(st, links, 16, seg_len=5)
for it, ann, wc in [(0, 1, 1), (1001, 0, 1)]:
    st.step(it, ann, wc); eng.step(it, ann, wc)
    compare_sweep(eng, st, "tsan iter %%d" %% it, check_member=True)
p = np.array([0, 1], dtype=np.uint32); q = np.array([2, 3], dtype=np.uint32); y = np.array([1, 0], dtype=np.uint8)
eng.heldout(p, q, y, blocks=2)
print("tsan run complete")
""" % (os.path.dirname(os.path.abspath(__file__)), lib)
    env = dict(os.environ, LD_PRELOAD=tc[1], TSAN_OPTIONS="exitcode=66 report_signal_unsafe=0")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert "tsan run complete" in out.stdout, out.stderr[-4000:]
    assert "WARNING: ThreadSanitizer: data race" not in out.stderr, out.stderr[-6000:]
    assert out.returncode == 0, out.stderr[-4000:]
