"""N > 1 path on CPU: svinet_b200.sharded.ShardedLinkSampling driven over gloo with world_size 2 (and 3),
a numpy stand-in executing the phases.  Checks the exchange plan (what is all-reduced / all-gathered and
when) by requiring the sharded result to equal the oracle's, sweep after sweep, incl. annealing rescale,
converged shortcuts and the link-community tally."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, k, links, gamma0, sched, conv0, out, overlap=True, ones=None):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from numpy_shard_engine import NumpyShardEngine
    from svinet_b200.sharded import ShardedLinkSampling
    sh = ShardedLinkSampling(n, k, links, rank=rank, world=world, engine_factory=NumpyShardEngine, overlap=overlap,
                             ones=ones)
    sh.set_state(gamma0, np.ones((k, 2)))
    sh.eng.conv[:] = conv0
    res = []
    for it, ann, wc in sched:
        sh.step(it, ann, wc)
        g, lam = sh.gather_state()
        res.append((g, lam, sh.eng.conv.copy(), sh.gather_membership() if wc else None))
    if rank == 0:
        torch.save({"res": res, "bounds": sh.bounds}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,overlap", [(2, True), (3, True), (2, False)])
def test_sharded_equals_oracle_over_gloo(world, overlap, tmp_path):
    sys.path.insert(0, HERE)
    import oracle_py as orc
    from svinet_b200 import synth
    n, k = 300, 12
    links = synth.mmsb_links(n, k, 2400, seed=3)
    gamma0, lam0 = synth.random_state(n, k, links, seed=4)
    rng = np.random.default_rng(0)
    conv0 = np.zeros(n, dtype=np.int32)
    who = rng.random(n) < 0.2
    conv0[who] = rng.integers(1, k + 1, who.sum())
    sched = [(0, 1, 0), (1, 1, 1), (2, 0, 1), (3, 0, 0)]

    st = orc.State.alloc(n, k, links.shape[0])
    c = st.c
    # `ones` (numerator of the annealing rescale, linksampling.cc:541-542) counts ALL links of the network, held-out ones
    # included: larger than the training-link count whenever a validation set is held out
    ones = links.shape[0] + (37 if world == 3 else 0)
    c.alpha, c.eta0, c.eta1, c.ones = 1.0 / k, 1.0, 1.0, ones
    st.arr("links")[:] = links
    tl = np.zeros(n); np.add.at(tl, links.ravel().astype(np.int64), 2.0)
    st.arr("tl")[:] = tl
    st.arr("gamma")[:] = gamma0; st.arr("gammanext")[:] = c.alpha
    st.arr("lambda_")[:] = 1.0; st.arr("lambdanext")[:] = 1.0
    st.arr("converged")[:] = conv0
    st.refresh_expectations()

    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), n, k, links, gamma0, sched, conv0, out, overlap, ones), nprocs=world, join=True)
    got = torch.load(out, weights_only=False)
    bounds = got["bounds"]
    assert bounds[0] == 0 and bounds[-1] == n and np.all(np.diff(bounds) > 0)
    for (it, ann, wc), (g, lam, conv, mem) in zip(sched, got["res"]):
        st.step(it, ann, wc)
        assert np.max(np.abs(g - st.arr("gamma")) / np.abs(st.arr("gamma"))) < 1e-9, it
        assert np.max(np.abs(lam - st.arr("lambda_")) / np.abs(st.arr("lambda_"))) < 1e-9, it
        assert np.array_equal(conv.astype(np.uint32), st.arr("converged")), it
        if wc:
            assert np.array_equal(mem, st.arr("member")), it
    st.free()


def test_plan_shards_is_edge_balanced():
    from svinet_b200.sharded import plan_shards
    from svinet_b200 import synth
    n = 5000
    links = synth.mmsb_links(n, 20, 60000, seed=1)
    # make the first nodes much heavier
    hub = np.stack([np.zeros(2000, dtype=np.uint32), np.arange(1, 2001, dtype=np.uint32)], 1)
    links = np.unique(np.concatenate([links, hub]), axis=0)
    deg = np.bincount(links.ravel().astype(np.int64), minlength=n)
    for world in (2, 4, 8):
        b = plan_shards(n, links, world)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1
        loads = np.array([deg[b[r]:b[r + 1]].sum() for r in range(world)])
        assert loads.max() <= loads.mean() + deg.max() + 1          # within one node's degree of the ideal
