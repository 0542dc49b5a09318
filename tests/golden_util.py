"""Helpers shared by the golden-vector tests (test infrastructure)."""
import gzip
from decimal import Decimal
import json
import os
import shutil
import tempfile

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "MANIFEST.json")))


def input_path(name, scratch):
    """Materialise tests/golden/inputs/<name>[.gz] as a plain file and return its path."""
    plain = os.path.join(GOLD, "inputs", name)
    if os.path.exists(plain):
        return plain
    out = os.path.join(scratch, name)
    if not os.path.exists(out):
        with gzip.open(plain + ".gz", "rb") as g, open(out, "wb") as f:
            shutil.copyfileobj(g, f)
    return out


def golden_text(case, fname):
    p = os.path.join(GOLD, case, fname)
    if os.path.exists(p):
        return open(p).read()
    if os.path.exists(p + ".gz"):
        return gzip.open(p + ".gz", "rt").read()
    return None


def compare_numeric_text(got, want, ulp_last_digit=1, skip_cols=()):
    """Field-by-field comparison of two tab/space separated numeric texts.

    Integer fields must match exactly; decimal fields may differ by `ulp_last_digit` units in
    the last PRINTED digit (a different FP64 summation order legitimately
    flips the final rounded digit).  Returns (n_fields, n_off_by_last_digit); raises on mismatch.
    """
    gl, wl = got.split("\n"), want.split("\n")
    assert len(gl) == len(wl), "line count %d != %d" % (len(gl), len(wl))
    nf = noff = 0
    for ln, (a, b) in enumerate(zip(gl, wl)):
        fa, fb = a.split(), b.split()
        assert len(fa) == len(fb), "line %d: field count %d != %d" % (ln, len(fa), len(fb))
        for col, (x, y) in enumerate(zip(fa, fb)):
            nf += 1
            if col in skip_cols or x == y:
                continue
            if "." in y and "nan" not in y and "inf" not in y:
                decimals = len(y.split(".")[1])
                # exact decimal arithmetic on the printed digits (binary floats cannot represent 1e-5 steps)
                diff = abs(Decimal(x) - Decimal(y)).scaleb(decimals)
                assert diff <= ulp_last_digit, "line %d col %d: %s vs %s" % (ln, col, x, y)
                noff += 1
            else:
                raise AssertionError("line %d col %d: %s vs %s" % (ln, col, x, y))
    return nf, noff


class Scratch:
    def __enter__(self):
        self.d = tempfile.mkdtemp(prefix="svinet_t_")
        return self.d

    def __exit__(self, *a):
        shutil.rmtree(self.d, ignore_errors=True)
