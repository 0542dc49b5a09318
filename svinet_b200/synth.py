"""Sparse synthetic assortative-MMSB edge lists (BASELINE.json configs 3-5).

The reference's own generator (src/mmsbgen.cc:44-71, src/mmsbgen.hh:144-208) draws every one of
the N^2/2 pairs: pi_i ~ Dirichlet(alpha), z_p->q ~ Mult(pi_p), z_q->p ~ Mult(pi_q),
y ~ Bernoulli(beta_k) iff z_p->q == z_q->p == k.  That is O(N^2) and unusable at N = 1e6, so the
same draw rule is sampled sparsely: a link of community k appears between p and q with probability
proportional to pi_pk * pi_qk * beta_k, hence
    k ~ Categorical( beta_k * (sum_i pi_ik)^2 ),  p, q ~ Categorical( pi_.k )  independently,
self-pairs and duplicates dropped.  Memberships are sparse (1-3 communities per node, Dirichlet(1)
weights), the regime alpha = 0.05 (src/main.cc:277) produces.  Node ids are randomly permuted so
that no locality is handed to the kernels for free.

torch is used only as an array library (CPU or CUDA); this is workload generation, not the path.
"""
import numpy as np
import torch


def mmsb_links(n, k, target_links, seed=1234, device="cpu", background=0.02, oversample=1.08):
    """Return (links[E,2] uint32 numpy with p<q, unique; membership list for diagnostics)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)

    def rand(*shape):
        return torch.rand(*shape, generator=g, device=dev, dtype=torch.float64)

    # sparse memberships: m_i in {1,2,3}
    u = rand(n)
    m = 1 + (u > 0.7).long() + (u > 0.9).long()
    slots = torch.arange(3, device=dev).unsqueeze(0) < m.unsqueeze(1)            # [n,3]
    comm = torch.randint(0, k, (n, 3), generator=g, device=dev)
    wts = -torch.log(rand(n, 3).clamp_min(1e-300)) * slots                       # Dirichlet(1) via exponentials
    wts = wts / wts.sum(1, keepdim=True)
    node = torch.arange(n, device=dev).unsqueeze(1).expand(n, 3)
    node, comm, wts = node[slots], comm[slots], wts[slots]
    order = torch.argsort(comm, stable=True)
    node, comm, wts = node[order], comm[order], wts[order]
    csize = torch.bincount(comm, minlength=k)
    cstart = torch.cumsum(csize, 0) - csize
    cmass = torch.zeros(k, device=dev, dtype=torch.float64).index_add_(0, comm, wts)
    beta = 0.5 + 0.5 * rand(k)                                                    # community strengths
    # within-community cumulative weights (global cumsum minus the community's offset)
    cw = torch.cumsum(wts, 0)
    coff = torch.cat([torch.zeros(1, device=dev, dtype=torch.float64), cw])[cstart]

    perm = torch.randperm(n, generator=g, device=dev)
    keys = torch.empty(0, dtype=torch.int64, device=dev)
    want = int(target_links)
    need = want
    rounds = 0
    while need > 0 and rounds < 12:
        rounds += 1
        draw = int(need * oversample) + 1024
        nb = int(draw * background)
        kk = torch.multinomial(beta * cmass * cmass, draw - nb, replacement=True, generator=g)

        def pick(kk):
            t = coff[kk] + rand(kk.shape[0]) * cmass[kk]
            idx = torch.searchsorted(cw, t).clamp_(max=cw.shape[0] - 1)
            idx = torch.minimum(torch.maximum(idx, cstart[kk]), cstart[kk] + csize[kk] - 1)
            return node[idx]

        p = torch.cat([pick(kk), torch.randint(0, n, (nb,), generator=g, device=dev)])
        q = torch.cat([pick(kk), torch.randint(0, n, (nb,), generator=g, device=dev)])
        p, q = perm[p], perm[q]
        lo, hi = torch.minimum(p, q), torch.maximum(p, q)
        ok = lo != hi
        new = lo[ok] * n + hi[ok]
        keys = torch.unique(torch.cat([keys, new]))
        need = want - keys.shape[0]
    if keys.shape[0] > want:  # drop a random surplus so the count is exact
        sel = torch.randperm(keys.shape[0], generator=g, device=dev)[:want]
        keys = keys[torch.sort(sel).values]
    links = torch.stack([keys // n, keys % n], 1).to(torch.int32).cpu().numpy().view(np.uint32)
    return np.ascontiguousarray(links)


def random_state(n, k, links, seed=0):
    """gamma as the reference's init_gamma2 would shape it (src/linksampling.cc:374-401: every link adds a
    normalised uniform K-vector to both endpoints), lambda = eta = (1,1).  Vectorised; NOT the reference's
    RNG stream -- for synthetic workloads only."""
    rng = np.random.default_rng(seed)
    gamma = np.zeros((n, k))
    step = max(1, (1 << 24) // max(k, 1))
    for s in range(0, links.shape[0], step):
        blk = links[s:s + step]
        phi = rng.random((blk.shape[0], k))
        phi /= phi.sum(1, keepdims=True)
        np.add.at(gamma, blk[:, 0], phi)
        np.add.at(gamma, blk[:, 1], phi)
    gamma[gamma.sum(1) == 0] = 1.0 / k      # isolated nodes: keep digamma's argument positive
    lam = np.ones((k, 2))
    return gamma, lam
