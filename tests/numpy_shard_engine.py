"""numpy stand-in for one shard's engine -- TEST INFRASTRUCTURE for the gloo (CPU) tests of
svinet_b200/sharded.py.  It follows the DEVICE formulation (pull form over the shard's half-edges,
factorised exp, deferred annealing rescale), phase by phase, so that the collective choreography of
ShardedLinkSampling can be exercised with world_size > 1 where no GPU exists.  The dense and shortcut
branches are covered; the iter > 1000 active-set branch is not (the CUDA tests cover it)."""
import numpy as np
import torch
from scipy.special import digamma


class NumpyShardEngine:
    def __init__(self, n, k, links, node_range, device, stream, tl=None, ones=None, alpha=None, eta0=1.0, eta1=1.0):
        self.n, self.k = n, k
        self.nb, self.ne = node_range
        self.alpha = 1.0 / k if alpha is None else alpha
        self.eta0, self.eta1 = eta0, eta1
        self.ones = float(ones)
        links = np.asarray(links, dtype=np.int64).reshape(-1, 2)
        self.tl = np.asarray(tl, dtype=np.float64)
        # half-edges whose source is local; "owned" ones (source is the link's smaller endpoint) feed s3
        src = np.concatenate([links[:, 0], links[:, 1]])
        dst = np.concatenate([links[:, 1], links[:, 0]])
        own = np.concatenate([np.ones(len(links), bool), np.zeros(len(links), bool)])
        keep = (src >= self.nb) & (src < self.ne)
        self.src, self.dst, self.own = src[keep], dst[keep], own[keep]
        z = lambda *s: np.zeros(s)
        self.b, self.mphi, self.gamma, self.gacc = z(n, k), z(n, k), z(n, k), z(n, k)
        self.kvec = z(4, k)
        self.lam = z(k, 2)
        self.eb, self.scale = z(k), np.ones(k)
        self.conv = np.zeros(n, dtype=np.int32)
        self.active = np.zeros(n, dtype=np.int32)
        self.words = (k + 31) // 32
        self.abits = np.zeros((n, self.words), dtype=np.int32)
        self.mbits = np.zeros((n, self.words), dtype=np.int32)
        self.acc = z(n, k)
        self.conv_snap = None

    def buffer(self, name):
        arr = {"exppi": self.b, "mphi": self.mphi, "gamma": self.gamma, "kvec": self.kvec, "converged": self.conv,
               "active": self.active, "active_bits": self.abits, "member_bits": self.mbits}[name]
        return torch.from_numpy(arr)

    def _refresh_rows(self, rows):
        g = self.gamma[rows]
        e = digamma(g) - digamma(g.sum(1, keepdims=True))
        self.b[rows] = np.exp(e - e.max(1, keepdims=True))

    def _refresh_lambda(self):
        e0 = digamma(self.lam[:, 0]) - digamma(self.lam.sum(1))
        self.eb[:] = np.exp(e0 - e0.max())

    def set_state(self, gamma, lam):
        self.gamma[:] = gamma
        self.lam[:] = lam
        self._refresh_rows(slice(0, self.n))
        self._refresh_lambda()

    def get_state(self):
        return self.gamma.copy(), self.lam.copy()

    def membership(self):
        cols = np.arange(self.k)
        return ((self.mbits[:, cols // 32] >> (cols % 32)) & 1).astype(np.uint8)

    def phase_phi(self, it, write_comm):
        if write_comm:
            self.mbits[:] = 0
        self.acc[self.nb:self.ne] = 0
        pc, qc = self.conv[self.src], self.conv[self.dst]
        short = (pc != 0) != (qc != 0)
        f = ~short
        w = self.b[self.src[f]] * self.eb[None, :] * self.b[self.dst[f]]
        phi = w / w.sum(1, keepdims=True)
        np.add.at(self.acc, self.src[f], phi)
        c = np.where(pc[short] != 0, pc[short], qc[short]) - 1
        np.add.at(self.acc, (self.src[short], c), 1.0)
        if write_comm:
            kmax = phi.argmax(1)
            np.bitwise_or.at(self.mbits, (self.src[f], kmax // 32), (1 << (kmax % 32)).astype(np.int32))

    def phase_node(self):
        rows = np.arange(self.nb, self.ne)
        acc, tl = self.acc[rows], self.tl[rows]
        has = tl != 0
        g = self.alpha + acc
        m = np.zeros_like(acc)
        m[has] = (g[has] - self.alpha) / tl[has, None]
        self.kvec[0] = acc[has].sum(0)
        self.kvec[1] = m[has].sum(0)
        self.kvec[2] = (m[has] ** 2).sum(0)
        g[has] += (self.n - tl[has, None] - 1) * m[has]
        g[~has] = self.alpha
        self.gacc[rows] = g
        self.mphi[rows[has]] = m[has]

    def phase_s3(self):
        s3 = np.zeros(self.k)
        o = self.own
        p, q = self.src[o], self.dst[o]
        conv = self.conv_snap if self.conv_snap is not None else self.conv
        pc, qc = conv[p], conv[q]
        full = ~((pc != 0) != (qc != 0))
        s3 += (self.mphi[p[full]] * self.mphi[q[full]]).sum(0)
        a = (pc != 0) & (qc == 0)
        col = pc[a]
        np.add.at(s3, col - 1, np.where(col < self.k, self.mphi[q[a], np.minimum(col, self.k - 1)], 0.0))
        bb = (pc == 0) & (qc != 0)
        col = qc[bb]
        np.add.at(s3, col - 1, np.where(col < self.k, self.mphi[p[bb], np.minimum(col, self.k - 1)], 0.0))
        self.kvec[3] = s3

    def phase_finish(self, annealing):
        self.phase_lambda(annealing)
        self._refresh_own_rows(annealing)

    def phase_lambda(self, annealing):
        s, s1, s2, s3 = self.kvec
        self.lam[:, 0] = self.eta0 + s
        self.lam[:, 1] = self.eta1 + (s1 * s1 - s2 - s3)
        self._refresh_lambda()
        self.conv_snap = None

    def phase_refresh(self, annealing):
        self.conv_snap = self.conv.copy()
        self._refresh_own_rows(annealing)

    def _refresh_own_rows(self, annealing):
        self.scale[:] = self.ones / self.kvec[0] if annealing else 1.0
        rows = np.arange(self.nb, self.ne)
        g = self.gacc[rows].copy()
        has = self.tl[rows] != 0
        g[has] *= self.scale[None, :]
        self.gamma[rows] = g
        self._refresh_rows(rows)
        act = (g - self.alpha >= 1)
        cnt = act.sum(1)
        self.active[rows] = cnt
        one = cnt == 1
        self.conv[rows[one]] = act[one].argmax(1) + 1
