"""The K > 512 kernels of `-rnode -stratified` (svinet_b200/csrc/svi_fa2_wide.cuh) against the FastAMM2 oracle WITHOUT a
GPU: the kernel source compiled as host code and run with one host thread per CUDA thread (tests/cc/fa2_wide_emul.cc),
driven like tests/test_gpu_fa2.py drives the device -- the pair fixed point against PhiCompute::update_phis_until_conv,
lockstep with the reference's own minibatch sequence (lazy and eager decay), the held-out likelihood -- and once more
under ThreadSanitizer (barrier placement).  The device build of the same kernels: tests/test_gpu_fa2.py (K > 512 cases).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_py as orc
import wide_emul_py as we
from golden_util import MANIFEST, Scratch, input_path
from test_oracle_fa2_golden import fa2_opts

TOL = 1e-9
FAST_T = 32


def rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def psi_rows(g):
    f = np.vectorize(orc.lib().orc_digamma)
    return f(g) - f(g.sum(axis=1, keepdims=True))


def pair_state(n, k, seed):
    rng = np.random.default_rng(seed)
    gamma = rng.gamma(1.0, 1.0, size=(n, k)) + 1e-3
    gamma[1] = 1.0 / k                       # a node still at its prior
    gamma[2, rng.integers(k)] += 40.0        # a peaked node
    lam = np.stack([1 + rng.gamma(2.0, 1.0, k), 1 + rng.gamma(5.0, 3.0, k)], axis=1)
    return gamma, lam


@pytest.mark.parametrize("k,threads", [(513, FAST_T), (1030, FAST_T), (2100, FAST_T), (600, 256)])
def test_phi_pair_matches_oracle(k, threads):
    gamma, lam = pair_state(6, k, k)
    eng = we.Fa2WideEmulEngine(6, k, threads=threads)
    eng.set_state(gamma, lam)
    epi, ebeta = psi_rows(gamma), psi_rows(lam)
    worst = 0.0
    for (p, q) in [(0, 1), (1, 2), (2, 3), (0, 5), (3, 4)]:
        for y in (0, 1):
            want1, want2, rounds = orc.fa2_phi_pair(epi[p], epi[q], ebeta[:, 0] if y else ebeta[:, 1], y)
            got1, got2, r = eng.phi_pair(p, q, y)
            assert r == rounds, (p, q, y, r, rounds)
            worst = max(worst, float(np.max(np.abs(got1 - want1))), float(np.max(np.abs(got2 - want2))))
    assert worst <= 1e-12, worst
    eng.close()


def lockstep(engine_factory, case, k, iters):
    """the reference's minibatch sequence on the fixture's graph, with `k` communities"""
    ent = MANIFEST[case]
    with Scratch() as d:
        g = orc.Graph.read(input_path(ent["input"], d), ent["n"])
        m = orc.Fa2Model(g, k, **fa2_opts(ent["flags"]))
        eng = engine_factory(m.n, m.k)
        eng.set_state(m.gamma, m.lambda_)
        types = [0, 0]
        for _ in range(iters):
            it = m.iter
            typ, start, pairs = m.plan()
            m.process()
            eng.step(it, typ, start, pairs)
            types[typ] += 1
            gam, lam = eng.get_state()
            wg, wl = rel_err(gam, m.gamma), rel_err(lam, m.lambda_)
            assert wg <= TOL and wl <= TOL, (case, it, typ, start, len(pairs), wg, wl)
        hp = m.heldout_pairs()
        y = np.array([g.y(int(a), int(b)) for a, b in hp], dtype=np.uint8)
        got = eng.heldout(hp[:, 0], hp[:, 1], y)
        want = np.array([m.edge_likelihood(int(a), int(b), int(yy)) for (a, b), yy in zip(hp, y)])
        assert len(hp) and rel_err(got, want, floor=1e-3) <= TOL
        eng.close(); m.close(); g.close()
    assert types[0] > 0 and types[1] > 0       # both samplers were exercised


@pytest.mark.parametrize("eager", [0, 1], ids=["lazy", "eager"])
def test_lockstep_with_reference_minibatches(eager):
    lockstep(lambda n, k: we.Fa2WideEmulEngine(n, k, eager_blend=eager, pair_blocks=3, threads=FAST_T), "fa2_c1_m200", 520, 24)


def test_lockstep_lazy_rows_across_a_rebase_and_full_size_blocks():
    """256-thread blocks; the re-basing fold of the lazy rows (k_fa2_fold) forced every few iterations"""
    seen = {}

    class Recording(we.Fa2WideEmulEngine):
        def close(self):
            if self.h:
                seen["folds"] = self.folds()
            super().close()

    lockstep(lambda n, k: Recording(n, k, pair_blocks=2, fold_below=-0.1, threads=256), "fa2_c1_m200", 515, 12)
    assert seen["folds"] >= 2


def test_barrier_placement_under_thread_sanitizer():
    tc = we.tsan_toolchain()
    if tc is None:
        pytest.skip("no g++ with libtsan here")
    lib = we.build(threads=16, tsan=True, src=we.FA2_SRC)
    code = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
import wide_emul_py as we, test_fa2_wide_emulated as t
we.build = lambda threads=256, tsan=False, src=None: %r          # the ThreadSanitizer build stands in for the plain one
for eager in (0, 1):
    t.lockstep(lambda n, k: we.Fa2WideEmulEngine(n, k, eager_blend=eager, pair_blocks=2, threads=16), "fa2_c1_m200", 514, 6)
gamma, lam = t.pair_state(6, 520, 1)
e = we.Fa2WideEmulEngine(6, 520, threads=16)
e.set_state(gamma, lam)
e.phi_pair(0, 1, 1)
print("tsan run complete")
""" % (os.path.dirname(os.path.abspath(__file__)), lib)
    env = dict(os.environ, LD_PRELOAD=tc[1], TSAN_OPTIONS="exitcode=66 report_signal_unsafe=0")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert "tsan run complete" in out.stdout, out.stderr[-4000:]
    assert "WARNING: ThreadSanitizer: data race" not in out.stderr, out.stderr[-6000:]
    assert out.returncode == 0, out.stderr[-4000:]
