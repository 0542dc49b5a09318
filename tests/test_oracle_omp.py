"""The all-cores leg of the CPU baseline (oracle/oracle_ls_omp.c) equals the serial oracle up to summation order."""
import numpy as np

import oracle_py as orc
from svinet_b200 import synth


def _state(n, k, links, seed):
    gamma, lam = synth.random_state(n, k, links, seed=seed)
    st = orc.State.alloc(n, k, links.shape[0])
    c = st.c
    c.alpha, c.eta0, c.eta1, c.ones = 1.0 / k, 1.0, 1.0, links.shape[0]
    st.arr("links")[:] = links
    tl = np.zeros(n); np.add.at(tl, links.ravel().astype(np.int64), 2.0)
    st.arr("tl")[:] = tl
    st.arr("gamma")[:] = gamma; st.arr("gammanext")[:] = c.alpha
    st.arr("lambda_")[:] = lam; st.arr("lambdanext")[:] = lam
    rng = np.random.default_rng(seed)
    who = rng.random(n) < 0.25
    st.arr("converged")[who] = rng.integers(1, k + 1, who.sum())
    st.refresh_expectations()
    return st


def test_omp_sweep_equals_serial_sweep():
    n, k = 500, 30
    links = synth.mmsb_links(n, k, 6000, seed=2)
    a, b = _state(n, k, links, 5), _state(n, k, links, 5)
    for it, ann, wc in [(0, 1, 0), (1, 1, 1), (2, 0, 1)]:
        a.step(it, ann, wc)
        b.step_omp(ann, wc, threads=4)
        for name in ("gamma", "lambda_", "mphi", "Elogpi"):
            x, y = a.arr(name), b.arr(name)
            assert np.max(np.abs(x - y) / np.maximum(np.abs(x), 1e-300)) < 1e-9, (it, name)
        for name in ("sum", "s1", "s2", "s3"):
            x, y = a.arr(name), b.arr(name)
            assert np.max(np.abs(x - y)) <= 1e-9 * max(1.0, np.max(np.abs(x))), (it, name)
        assert np.array_equal(a.arr("converged"), b.arr("converged"))
        assert np.array_equal(a.arr("active_comms"), b.arr("active_comms"))
        if wc:
            assert np.array_equal(a.arr("member"), b.arr("member"))
    assert b.c.cnt_shortcut > 0 and b.c.cnt_shortcut == a.c.cnt_shortcut
    a.free(); b.free()
