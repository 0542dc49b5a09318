"""Comparison helpers shared by the parity tests (GPU through the C ABI, and the host-thread emulation of the K > 1024
kernels): no torch, no GPU needed to import.

Bar (BASELINE.json north_star): gamma/lambda within 1e-5 relative after a fixed seed and a fixed iteration count.  The
FP64 device path actually lands near 1e-12; the tests assert 1e-9 so that a regression in arithmetic (a wrong branch, a
missed term) cannot hide inside the official tolerance.  Integer outputs (converged, active_comms, link-community
membership) must match exactly.
"""
import numpy as np

from svinet_b200.engine import LinkSamplingEngine

TOL = 1e-9
TOL_OFFICIAL = 1e-5


def rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def engine_from_state(st, ones, **kw):
    c = st.c
    eng = LinkSamplingEngine(c.n, c.k, st.arr("links"), tl=st.arr("tl"), alpha=c.alpha, eta0=c.eta0,
                             eta1=c.eta1, ones=ones, **kw)
    eng.set_state(st.arr("gamma"), st.arr("lambda_"))
    eng.set_converged(st.arr("converged"))
    return eng


def compare_sweep(eng, st, tag, tol=TOL, check_member=False):
    g, lam = eng.get_state()
    kv = eng.kvectors()
    conv, act = eng.get_converged()
    errs = {"gamma": rel_err(g, st.arr("gamma")), "lambda": rel_err(lam, st.arr("lambda_"))}
    for name in ("sum", "s1", "s2", "s3"):
        # column sums feed lambda as eta + (...): what matters is the error relative to the vector's scale
        # (a community holding ~1e-11 of the mass suffers (alpha + x) - alpha cancellation in BOTH codes)
        want = st.arr(name)
        errs[name] = rel_err(kv[name], want, floor=max(1e-12, 1e-6 * float(np.max(np.abs(want), initial=0.0))))
    for name, e in errs.items():
        assert e <= tol, "%s: %s rel err %.3e" % (tag, name, e)
    assert np.array_equal(conv, st.arr("converged")), "%s: converged differs" % tag
    assert np.array_equal(act, st.arr("active_comms")), "%s: active_comms differs" % tag
    if check_member:
        assert np.array_equal(eng.membership(), st.arr("member")), "%s: membership differs" % tag
    return errs
