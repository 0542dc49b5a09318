"""The sharded CUDA path (svi_ls_mg_step: node-block shards exchanging rows over peer memory, include/svi_ls.h)
against the oracle ON HARDWARE -- needs one B200.

All shards live in this process and on the same device (svi_ls_peer_attach_local): the kernels restricted to a node
block, the chunked phi/node pipeline, the refresh-before-s3 order with the double-buffered converged flags, the row
pushes, the flag protocol and the slot all-reduces are exactly the code the multi-GPU runs execute; only the copies
stay inside one GPU.  (bench.py under torchrun runs the same calls with one process per GPU and CUDA IPC; its
state checksum must equal the single-GPU one, see bench.py --checksum.)
"""
import numpy as np
import pytest

import oracle_py as orc
from svinet_b200 import synth
from svinet_b200.engine import LinkSamplingEngine
from svinet_b200.sharded import plan_shards
from test_gpu_parity import TOL, rel_err
from test_gpu_parity_ring import oracle_state

pytestmark = pytest.mark.gpu


def run_sharded_vs_oracle(n, k, links, gamma, conv, world, chunks, sched, seg_len=0, share=True):
    st = oracle_state(n, k, links, gamma, np.ones((k, 2)), conv)
    bounds = plan_shards(n, links, world)
    engines = [LinkSamplingEngine(n, k, links, node_range=(int(bounds[r]), int(bounds[r + 1])), seg_len=seg_len)
               for r in range(world)]
    LinkSamplingEngine.attach_local(engines, bounds, chunks=chunks)
    rng = np.random.default_rng(7)
    hp = rng.integers(0, n, 300).astype(np.uint32)
    hq = ((hp + 1 + rng.integers(0, n - 1, 300)) % n).astype(np.uint32)
    hy = rng.integers(0, 2, 300).astype(np.uint8)
    for e in engines:
        if share:
            e.mg_share_gamma(True)
        e.set_state(st.arr("gamma"), st.arr("lambda_"))
        e.set_converged(st.arr("converged"))
    for it, ann, wc in sched:
        st.step(it, ann, wc)
        for e in engines:
            e.mg_step(it, ann, wc)
        for e in engines:
            e.sync()
        # held-out likelihood of arbitrary pairs on every shard: rows of other shards are replicated (share) or read
        # from the owner's arena (peer loads)
        want_ll = np.array([st.edge_likelihood(int(a), int(b), int(c)) for a, b, c in zip(hp, hq, hy)])
        for e in engines:
            assert rel_err(e.heldout(hp, hq, hy), want_ll, floor=1e-3) <= TOL
        if not share:
            for e in engines:
                e.mg_publish_gamma()
        mem = np.zeros((n, k), dtype=np.uint8)
        for r, e in enumerate(engines):
            tag = "world=%d shard %d iter %d" % (world, r, it)
            g, lam = e.get_state()
            assert rel_err(g, st.arr("gamma")) <= TOL, tag
            assert rel_err(lam, st.arr("lambda_")) <= TOL, tag
            kv = e.kvectors()
            for name in ("sum", "s1", "s2", "s3"):
                want = st.arr(name)
                assert rel_err(kv[name], want, floor=max(1e-12, 1e-6 * float(np.max(np.abs(want))))) <= TOL, (tag, name)
            cv, act = e.get_converged()
            assert np.array_equal(cv, st.arr("converged")), tag
            nb, ne = int(bounds[r]), int(bounds[r + 1])
            assert np.array_equal(act[nb:ne], st.arr("active_comms")[nb:ne]), tag
            if wc:
                bits = e.membership_rows(nb, ne - nb)
                cols = np.arange(k)
                mem[nb:ne] = (bits[:, cols // 32] >> (cols % 32).astype(np.uint32)) & 1
        if wc:
            assert np.array_equal(mem, st.arr("member")), "membership, iter %d" % it
        # every shard ends the iteration with bit-identical replicated state
        g0, l0 = engines[0].get_state()
        for e in engines[1:]:
            g, lam = e.get_state()
            assert np.array_equal(g, g0) and np.array_equal(lam, l0)
    shortcut = st.c.cnt_shortcut
    for e in engines:
        e.close()
    st.free()
    return shortcut


@pytest.mark.parametrize("world,k,chunks,share", [(2, 12, 1, True), (2, 200, 4, False), (3, 100, 3, False), (4, 64, 2, True)])
def test_shards_on_one_gpu_match_oracle(world, k, chunks, share):
    n = 900
    links = synth.mmsb_links(n, k, 30 * n, seed=50 + k)
    gamma, _ = synth.random_state(n, k, links, seed=k)
    rng = np.random.default_rng(world)
    conv = np.zeros(n, dtype=np.uint32)
    who = rng.random(n) < 0.25
    conv[who] = rng.integers(1, k + 1, who.sum())
    sched = [(0, 1, 0), (1, 1, 1), (2, 0, 1), (3, 0, 0), (4, 1, 1)]
    assert run_sharded_vs_oracle(n, k, links, gamma, conv, world, chunks, sched, seg_len=64, share=share) > 0


def test_shards_where_nodes_converge_on_the_way():
    """assort-75-4 converges node after node: the flags pruned by one shard must reach the others' next sweep, and
    the s3 sweep of the SAME iteration must still see the pre-prune flags (src/linksampling.cc:731-761)."""
    from golden_util import Scratch, input_path
    with Scratch() as d:
        g = orc.Graph.read(input_path("assort-75-4.txt", d), 75)
        m = orc.Model(g, 4, use_validation_stop=0)
        st = m.state
        links, gamma = st.arr("links").copy(), st.arr("gamma").copy()
        m.close(); g.close()
    sched = [(it, it < 14, 1) for it in range(28)]
    run_sharded_vs_oracle(75, 4, links, gamma, np.zeros(75, dtype=np.uint32), 3, 2, sched, share=False)


@pytest.mark.parametrize("overlap", [False, True])
def test_phase_level_shards_with_caller_side_exchange(overlap):
    """The phase-level entry points a driver with its own collectives uses (svinet_b200/sharded.py, exchange="nccl"):
    two shard handles on one GPU, the exchanges done here with plain tensor copies on the buffers of
    svi_ls_device_buffer -- the same choreography as ShardedLinkSampling.step / _step_overlapped, incl. the
    double-buffered `converged` pointer -- against the oracle."""
    import torch
    from svinet_b200.sharded import cuda_view
    n, k = 700, 40
    links = synth.mmsb_links(n, k, 25 * n, seed=77)
    gamma, _ = synth.random_state(n, k, links, seed=5)
    rng = np.random.default_rng(1)
    conv = np.zeros(n, dtype=np.uint32)
    who = rng.random(n) < 0.25
    conv[who] = rng.integers(1, k + 1, who.sum())
    st = oracle_state(n, k, links, gamma, np.ones((k, 2)), conv)
    bounds = plan_shards(n, links, 2)
    tl = 2.0 * np.bincount(links.ravel().astype(np.int64), minlength=n).astype(np.float64)
    engines = [LinkSamplingEngine(n, k, links, tl=tl, node_range=(int(bounds[r]), int(bounds[r + 1])), seg_len=64)
               for r in range(2)]
    dev = torch.device("cuda", torch.cuda.current_device())
    ld = engines[0].info()["ld"]
    words = (k + 31) // 32

    def view(e, name):
        ptr, _ = e.device_buffer(name)
        shape, dt = {"exppi": ((n, ld), torch.float64), "mphi": ((n, ld), torch.float64), "gamma": ((n, ld), torch.float64),
                     "kvec": ((4, ld), torch.float64), "converged": ((n,), torch.int32), "active": ((n,), torch.int32),
                     "active_bits": ((n, words), torch.int32), "member_bits": ((n, words), torch.int32)}[name]
        return cuda_view(ptr, shape, dt, dev)

    def allgather(name):                      # every shard's block of `name` into the other handle
        for r, e in enumerate(engines):
            e.sync()
        for r in range(2):
            nb, ne = int(bounds[r]), int(bounds[r + 1])
            view(engines[1 - r], name)[nb:ne] = view(engines[r], name)[nb:ne]
        torch.cuda.synchronize()

    def allreduce(rows):
        for e in engines:
            e.sync()
        tot = view(engines[0], "kvec")[rows].clone() + view(engines[1], "kvec")[rows]
        for e in engines:
            view(e, "kvec")[rows] = tot
        torch.cuda.synchronize()

    for e in engines:
        e.set_state(st.arr("gamma"), st.arr("lambda_"))
        e.set_converged(st.arr("converged"))
    for it, ann, wc in [(0, 1, 0), (1, 1, 1), (2, 0, 1), (3, 1, 1)]:
        st.step(it, ann, wc)
        for e in engines:
            e.phase_phi(it, wc); e.phase_node()
        allreduce(slice(0, 3))
        allgather("mphi")
        if overlap:
            for e in engines:
                e.phase_refresh(ann)
            allgather("exppi"); allgather("converged")
            for e in engines:
                e.phase_s3()
            allreduce(slice(3, 4))
            for e in engines:
                e.phase_lambda(ann)
        else:
            for e in engines:
                e.phase_s3()
            allreduce(slice(3, 4))
            for e in engines:
                e.phase_finish(ann)
            allgather("exppi"); allgather("converged")
        allgather("gamma")
        mem = np.zeros((n, k), dtype=np.uint8)
        for r, e in enumerate(engines):
            g, lam = e.get_state()
            assert rel_err(g, st.arr("gamma")) <= TOL and rel_err(lam, st.arr("lambda_")) <= TOL, (it, r)
            assert np.array_equal(e.get_converged()[0], st.arr("converged")), (it, r)
            nb, ne = int(bounds[r]), int(bounds[r + 1])
            mem[nb:ne] = e.membership()[nb:ne]
        if wc:
            assert np.array_equal(mem, st.arr("member")), it
    for e in engines:
        e.close()
    st.free()
