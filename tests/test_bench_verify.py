"""bench.py's --verify recomputation (reference formulation in numpy + scipy digamma) checked against the
oracle on CPU: the verifier must accept the oracle's own sweep and reject a corrupted one."""
import numpy as np

import oracle_py as orc
import bench
from svinet_b200 import synth


def _state(n, k, links, conv_frac, seed):
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
    who = rng.random(n) < conv_frac
    st.arr("converged")[who] = rng.integers(1, k + 1, who.sum())
    st.refresh_expectations()
    return st


def test_verifier_accepts_the_oracle_and_rejects_a_wrong_row():
    n, k = 400, 24
    links = synth.mmsb_links(n, k, 6000, seed=8)
    for corrupt in (False, True):
        st = _state(n, k, links, 0.3, seed=21)
        st.step(0, 1, 0)                                     # a state with real structure

        def step_once():
            st.step(1, 0, 1)
            if corrupt:                                      # what a stale ring slot would do: one neighbour row off
                st.arr("gamma")[:] = st.arr("gamma") * (1.0 + 1e-6 * (np.arange(n)[:, None] % 7 == 0))

        out = bench.verify_sampled_rows(step_once, lambda: (st.arr("gamma").copy(), st.arr("lambda_").copy()),
                                        lambda: st.arr("converged").copy(), n, k, links, 1.0 / k, sample=128)
        assert out["rows"] == 128 and out["shortcut_half_edges"] > 0
        assert out["ok"] == (not corrupt), out
        st.free()
