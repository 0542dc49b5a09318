"""Pin the FastAMM2 oracle (oracle/oracle_fa2.c, `-rnode -stratified`) to the UNMODIFIED reference.

tests/golden/fa2_*/ were written by oracle/_ref/svinet_ref via oracle/make_golden.py.  The C restatement
re-runs each case from the same input and flags; every kept file must come out byte-identical
(heldout.txt modulo its wall-clock column, which the fixture zeroes)."""
import os

import pytest

import oracle_py as orc
from golden_util import MANIFEST, Scratch, compare_numeric_text, golden_text, input_path

FA2 = [c for c in MANIFEST if MANIFEST[c].get("mode") == "-rnode -stratified"]
FILES = ("gamma.txt", "lambda.txt", "heldout.txt", "heldout-pairs.txt", "groups.txt", "communities.txt",
         "communities_size.txt", "summary.txt")


def fa2_opts(flags):
    o, i = {}, 0
    while i < len(flags):
        f = flags[i]
        if f == "-max-iterations":
            o["max_iterations"] = int(flags[i + 1]); i += 1
        elif f == "-rfreq":
            o["reportfreq"] = int(flags[i + 1]); i += 1
        elif f == "-seed":
            o["seed"] = float(flags[i + 1]); i += 1
        else:
            raise KeyError(f)
        i += 1
    return o


@pytest.mark.parametrize("case", FA2)
def test_fa2_oracle_matches_reference(case):
    ent = MANIFEST[case]
    with Scratch() as d:
        g = orc.Graph.read(input_path(ent["input"], d), ent["n"])
        m = orc.Fa2Model(g, ent["k"], **fa2_opts(ent["flags"]))
        steps = m.run()
        assert m.stopped and steps == fa2_opts(ent["flags"])["max_iterations"] + 1   # fastamm2.cc:546, sic
        out = os.path.join(d, "out")
        m.write_outputs(out)
        for fname in FILES:
            want = golden_text(case, fname)
            assert want is not None, fname
            got = open(os.path.join(out, fname)).read()
            if fname in ("gamma.txt", "lambda.txt", "heldout.txt", "groups.txt"):
                nf, noff = compare_numeric_text(got, want, skip_cols=(1,) if fname == "heldout.txt" else ())
                assert noff == 0, (fname, nf, noff)          # byte-identical in this container
            else:
                assert got == want, "%s/%s differs" % (case, fname)
        m.close()
        g.close()


def test_fa2_phi_pair_properties():
    """update_phis_until_conv: both outputs are distributions; identical rows + symmetric start stay symmetric."""
    import numpy as np
    rng = np.random.default_rng(3)
    k = 9
    ep = np.log(rng.dirichlet(np.ones(k)))
    ef = -rng.random(k) * 3
    for y in (0, 1):
        p1, p2, rounds = orc.fa2_phi_pair(ep, ep, ef, y)
        assert 2 <= rounds <= 50
        assert abs(p1.sum() - 1) < 1e-12 and abs(p2.sum() - 1) < 1e-12
        assert np.allclose(p1, p2, rtol=0, atol=1e-15)
