"""K > 512 of `-rnode -stratified` on the device: the block-per-pair kernels of svinet_b200/csrc/svi_fa2_wide.cuh
through the C ABI (include/svi_fa2.h) against the FastAMM2 oracle -- needs a B200.

The kernel source is checked against the oracle on host threads, and its barriers under ThreadSanitizer, in
tests/test_fa2_wide_emulated.py; here the device build and its dispatch in svi_fa2.cu run.  Cases mirror
tests/test_gpu_fa2.py.  NOT YET RUN ON HARDWARE when it was written (the round's GPU budget was spent): this file sorts
last among the GPU tests on purpose.
"""
import numpy as np
import pytest

import oracle_py as orc
from svinet_b200.fa2_engine import Fa2Engine
from test_fa2_wide_emulated import lockstep, pair_state, psi_rows, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [513, 1030, 2100, 4100])
def test_phi_pair_matches_oracle(k):
    gamma, lam = pair_state(6, k, k)
    eng = Fa2Engine(6, k)
    info = eng.info()
    assert info["lanes"] == 256 and 2 * info["lanes"] * info["vec"] >= info["ld"]
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


@pytest.mark.parametrize("eager", [0, 1], ids=["lazy", "eager"])
@pytest.mark.parametrize("case,k,iters", [("fa2_c1_m200", 520, 60), ("fa2_lfr_k28_m300", 1100, 30)])
def test_lockstep_with_reference_minibatches(case, k, iters, eager):
    """the reference's own minibatch sequence on the fixture's graph, with K communities: gamma / lambda after every
    iteration, the held-out likelihood at the end; both treatments of the untouched rows' decay"""
    lockstep(lambda n, kk: Fa2Engine(n, kk, eager_blend=eager), case, k, iters)


def test_device_draws_lazy_equals_eager_and_is_deterministic():
    """svi_fa2_run (Philox minibatches drawn on the device) with the wide tile: the scalar decay equals the explicit
    pass, and two runs are bit-identical (no floating-point atomics)"""
    from test_gpu_fa2 import _synthetic
    n, k = 150, 600
    links, gamma, lam, heldout, shuffled = _synthetic(n, k, 6, seed=3)
    out = []
    for eager in (0, 1, 0):
        e = Fa2Engine(n, k, eager_blend=eager)
        e.set_state(gamma, lam)
        e.set_graph(links, heldout, shuffled)
        e.run(0, 300, 99, count=False)
        out.append(e.get_state())
        e.close()
    (gl, ll), (ge, le), (g2, l2) = out
    assert np.all(np.isfinite(gl)) and np.all(gl > 0)
    assert rel_err(gl, ge) <= 1e-9 and rel_err(ll, le) <= 1e-9
    assert np.array_equal(gl, g2) and np.array_equal(ll, l2)


def test_cli_output_directory_at_k_700():
    """`svinet -rnode -stratified -k 700 -max-iterations 120 -rfreq 40` against the files the FastAMM2 oracle's writers
    produce for the same run (pinned to the reference's bytes on the golden fixtures, tests/test_oracle_fa2_golden.py)."""
    import os
    import subprocess
    from golden_util import Scratch, compare_numeric_text, input_path
    from svinet_b200 import build as svbuild
    cli = svbuild.build_cli()
    k = 700
    with Scratch() as d:
        inp = input_path("assort-75-4.txt", d)
        if not os.path.exists(os.path.join(d, "assort-75-4.txt")):
            os.symlink(inp, os.path.join(d, "assort-75-4.txt"))
        p = subprocess.run([cli, "-file", "assort-75-4.txt", "-n", "75", "-k", str(k), "-rnode", "-stratified",
                            "-max-iterations", "120", "-rfreq", "40"], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                           timeout=600)
        assert p.returncode == 0, p.stderr.decode()
        out = os.path.join(d, "n75-k%d-mmsb-Srnode" % k)
        g = orc.Graph.read(inp, 75)
        m = orc.Fa2Model(g, k, max_iterations=120, reportfreq=40)
        m.run()
        want = os.path.join(d, "want")
        m.write_outputs(want)
        m.close(); g.close()
        flips = {}
        for fname in ("gamma.txt", "lambda.txt", "groups.txt", "heldout.txt"):
            flips[fname] = compare_numeric_text(open(os.path.join(out, fname)).read(), open(os.path.join(want, fname)).read(),
                                                skip_cols=(1,) if fname == "heldout.txt" else ())
        for fname in ("communities.txt", "communities_size.txt", "summary.txt", "heldout-pairs.txt"):
            assert open(os.path.join(out, fname)).read() == open(os.path.join(want, fname)).read(), fname
        nf, noff = flips["gamma.txt"]
        assert noff <= max(2, nf // 1000), flips
