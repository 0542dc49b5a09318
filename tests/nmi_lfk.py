"""Normalised mutual information of two COVERS (overlapping communities) after Lancichinetti, Fortunato & Kertesz,
New J. Phys. 11 (2009) 033015, appendix B -- the measure the reference obtains from the external `mutual` binary
(src/linksampling.cc:843-851, not shipped with it).  Test infrastructure."""
import numpy as np


def _h(p):
    p = np.asarray(p, dtype=np.float64)
    out = np.zeros_like(p)
    pos = p > 0
    out[pos] = -p[pos] * np.log2(p[pos])
    return out


def _cond(a, b, n):
    """(1/|A|) sum_k H(A_k | B) / H(A_k) for boolean membership matrices a [n, ca], b [n, cb]"""
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    p11 = a.T @ b / n
    pa, pb = a.sum(0) / n, b.sum(0) / n
    p10 = pa[:, None] - p11
    p01 = pb[None, :] - p11
    p00 = 1.0 - p11 - p10 - p01
    hj = _h(p11) + _h(p10) + _h(p01) + _h(p00)
    hb = _h(pb) + _h(1 - pb)
    ha = _h(pa) + _h(1 - pa)
    cond = hj - hb[None, :]                                    # H(A_k | B_l)
    ok = _h(p11) + _h(p00) > _h(p01) + _h(p10)                 # eq. (B.14): excludes complements
    cond = np.where(ok, cond, np.inf)
    best = cond.min(1) if cond.shape[1] else np.full(a.shape[1], np.inf)
    best = np.where(np.isfinite(best), best, ha)
    use = ha > 0
    return float(np.mean(best[use] / ha[use])) if use.any() else 0.0


def cover_from_lines(lines, index):
    """lines: iterable of whitespace separated node ids, one community per line"""
    comms = [[index[int(t)] for t in ln.split()] for ln in lines if ln.strip()]
    m = np.zeros((len(index), len(comms)), dtype=bool)
    for c, nodes in enumerate(comms):
        m[nodes, c] = True
    return m


def nmi_lfk(a, b):
    n = a.shape[0]
    return 1.0 - 0.5 * (_cond(a, b, n) + _cond(b, a, n))
