"""Pin the oracle (oracle/oracle_ls.c) to the UNMODIFIED reference.

The fixtures under tests/golden/<case>/ were written by oracle/_ref/svinet_ref (the reference's own
sources compiled with the GSL stand-in) via oracle/make_golden.py.  Here the C restatement re-runs
each case from the same input file and flags, and its text outputs must match the reference's:
identical integers / structure, decimals within +-1 unit of the last printed digit (SURVEY.md
Appendix F explains why byte-equality is too strict for a different FP64 summation order).
"""
import os

import pytest

import oracle_py as orc
from golden_util import MANIFEST, Scratch, compare_numeric_text, golden_text, input_path


def _flags_to_opts(flags, scratch=None):
    o = {}
    i = 0
    while i < len(flags):
        f = flags[i]
        if f == "-max-iterations":
            o["max_iterations"] = int(flags[i + 1]); i += 1
        elif f == "-no-stop":
            o["use_validation_stop"] = 0
        elif f == "-seed":
            o["seed"] = float(flags[i + 1]); i += 1
        elif f == "-accuracy":
            o["accuracy"] = 1
        elif f == "-init-communities":              # linksampling.cc:113-116
            o["init_communities"] = input_path(flags[i + 1], scratch).encode(); i += 1
        elif f == "-eta-type":                      # network.cc:233-250
            o["eta0"], o["eta1"] = {"uniform": (1.0, 1.0), "sparse": (0.97, 6.33), "dense": (4700.59, 0.77)}[flags[i + 1]]
            i += 1
        else:
            raise KeyError(f)
        i += 1
    return o


LS = [c for c in MANIFEST if MANIFEST[c].get("mode", "-link-sampling") == "-link-sampling"]
FAST = [c for c in LS if not c.startswith("c2_")]
SLOW = [c for c in LS if c.startswith("c2_")]


def _run_case(case):
    ent = MANIFEST[case]
    with Scratch() as d:
        g = orc.Graph.read(input_path(ent["input"], d), ent["n"])
        m = orc.Model(g, ent["k"], **_flags_to_opts(ent["flags"], d))
        m.run()
        assert m.stopped
        out = os.path.join(d, "out")
        m.write_outputs(out)
        report = {}
        for fname in ("lambda.txt", "gamma.txt", "groups.txt", "communities.txt", "validation.txt", "max.txt",
                      "validation-edges.txt"):
            want = golden_text(case, fname)
            if want is None:
                continue
            got = open(os.path.join(out, fname)).read()
            if fname in ("communities.txt", "validation-edges.txt"):
                assert got == want, "%s/%s differs" % (case, fname)
                report[fname] = (0, 0)
            else:
                report[fname] = compare_numeric_text(got, want)
        m.close()
        g.close()
    return report


@pytest.mark.parametrize("case", FAST)
def test_oracle_matches_reference_small(case):
    rep = _run_case(case)
    assert "gamma.txt" in rep and "lambda.txt" in rep
    # the tiny cases come out with at most a handful of last-digit flips
    nf, noff = rep["gamma.txt"]
    assert noff <= max(2, nf // 10000), rep


@pytest.mark.slow
@pytest.mark.parametrize("case", SLOW)
def test_oracle_matches_reference_astroph(case):
    rep = _run_case(case)
    if "gamma.txt" not in rep:        # c2_natural keeps the small files only: the stop iteration is the point
        assert golden_text(case, "max.txt").split("\t")[0] == "30" and rep["max.txt"][1] == 0 and rep["lambda.txt"][1] <= 1
        return
    nf, noff = rep["gamma.txt"]
    assert nf == 17903 * 22
    assert noff <= nf // 10000, rep


def test_natural_stop_iteration_matches_reference():
    # c1_natural ran with the validation stop enabled: the restated stop machine must end on the
    # same iteration the reference did (max.txt first column; SURVEY.md section 6: iteration 31).
    want = golden_text("c1_natural", "max.txt").split("\t")
    assert want[0] == "31" and want[-1].strip() == "1"
