"""Node-block sharding of the link-sampling iteration across the GPUs of one box.

One process per GPU (torchrun), torch.distributed for the plumbing.  Rank r owns a contiguous block of
nodes chosen so that every rank sweeps about the same number of half-edges; it holds the CSR of the
half-edges whose SOURCE is in its block (pull form: a link is processed by both endpoint owners, each
updating only its own row, so there is no cross-GPU scatter), the full N x K factor / mean-indicator
matrices (neighbour rows can live anywhere), and refreshes only its own rows.  Per iteration the path
has three real exchange steps, done with NCCL over NVLink/NVSwitch:

    phase_phi ; phase_node
        all-reduce   sum, s1, s2            (3 K-vectors; `sum` feeds the annealing rescale, :541-542)
        all-gather   mphi rows              (N x K doubles in total)      | beside  phase_refresh
    phase_s3                                                              | beside  all-gather exp(Elogpi) rows,
        all-reduce   s3                     (1 K-vector)                  |         converged[] (+ active masks)
    phase_lambda

(`overlap=False` keeps the plain order phase_s3 ; phase_finish ; all-gathers.)

The row all-gathers are the only bulk traffic: 2 x N*K*8 bytes per iteration (3.2 GB at n=1M, k=200)
against ~480 GB of local HBM traffic divided by the number of GPUs.

Two exchange mechanisms:
  * exchange="peer" (default on CUDA): the whole iteration incl. its exchanges runs inside the C library
    (svi_ls_mg_step, include/svi_ls.h): rows are pushed into the peers' arenas over NVLink by the copy engines beside
    the sweeps (the mphi rows of a finished chunk beside the next chunk's phi sweep, the exp(Elogpi) rows beside the
    s3 sweep), announced by epoch flags; K-vectors are reduced through per-source slots in a fixed order.
    torch.distributed only carries the CUDA IPC handles at start-up and the timing barriers.
  * exchange="nccl": the choreography below over torch.distributed collectives -- the round-1 path, kept as the
    library baseline to compare against and as what the CPU (gloo) tests drive.

The collective choreography is independent of what executes the phases: `engine_factory` lets the CPU
tests (gloo, world_size 2) drive it with a numpy stand-in; the product path builds the CUDA engine and
refuses to run without a GPU.
"""
import numpy as np
import torch
import torch.distributed as dist


def plan_shards(n, links, world):
    """Edge-balanced contiguous node blocks: boundaries[r] .. boundaries[r+1] holds ~1/world of the
    half-edges (every link contributes one half-edge to each endpoint)."""
    links = np.asarray(links).reshape(-1, 2)
    deg = np.bincount(links.ravel().astype(np.int64), minlength=n).astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(deg)])
    total = csum[-1]
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        b = int(np.searchsorted(csum, target, side="left"))
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


class _DevArray:
    """Zero-copy torch view of a device buffer owned by the C library (__cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr,
                                         "version": 2, "strides": None}


def cuda_view(ptr, shape, dtype, device):
    typestr = {torch.float64: "<f8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


class CudaShardEngine:
    """The product engine for one shard: libsvi_ls.so through svinet_b200.engine (CUDA only)."""

    def __init__(self, n, k, links, node_range, device, stream, **kw):
        if not torch.cuda.is_available():
            raise RuntimeError("svinet_b200.sharded: no CUDA device; the product path has no CPU fallback")
        from .engine import LinkSamplingEngine
        self.eng = LinkSamplingEngine(n, k, links, device=device, node_range=node_range, stream=stream, **kw)
        self.n, self.k = n, k
        self.device = torch.device("cuda", device)
        self.ld = self.eng.info()["ld"]
        self.words = (k + 31) // 32

    def buffer(self, name):
        ptr, ld = self.eng.device_buffer(name)
        if name in ("exppi", "mphi", "gamma"):
            return cuda_view(ptr, (self.n, self.ld), torch.float64, self.device)
        if name == "kvec":
            return cuda_view(ptr, (4, self.ld), torch.float64, self.device)
        if name in ("converged", "active"):
            return cuda_view(ptr, (self.n,), torch.int32, self.device)
        if name in ("active_bits", "member_bits"):
            return cuda_view(ptr, (self.n, self.words), torch.int32, self.device)
        raise KeyError(name)

    def __getattr__(self, item):          # phase_phi, phase_node, phase_s3, phase_finish, set_state, ...
        return getattr(self.eng, item)


class ShardedLinkSampling:
    def __init__(self, n, k, links, rank, world, device=0, stream=None, engine_factory=None, group=None,
                 overlap=True, exchange=None, ones=None, chunks=0, share_gamma=False, **kw):
        self.n, self.k, self.rank, self.world, self.group = n, k, rank, world, group
        if exchange is None:
            exchange = "peer" if (engine_factory is None and torch.cuda.is_available()) else "nccl"
        self.exchange = exchange
        if exchange == "peer":
            self._init_peer(n, k, links, rank, world, device, stream, ones, chunks, share_gamma, kw)
            return
        # overlap=True: the iteration order of _step_overlapped; over NCCL the bulk row exchanges additionally get
        # their own communicator and a side stream so that they run beside the kernels
        self.overlap = overlap
        self.side = self.bulk_group = None
        if overlap and world > 1 and dist.get_backend(group) == "nccl":
            ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
            self.bulk_group = dist.new_group(ranks=ranks, backend="nccl")
            self.side = torch.cuda.Stream(device=device)
        links = np.ascontiguousarray(links, dtype=np.uint32).reshape(-1, 2)
        self.bounds = plan_shards(n, links, world)
        nb, ne = int(self.bounds[rank]), int(self.bounds[rank + 1])
        # only the links incident to this block are needed to build the shard's CSR; tl (2 x degree) is a
        # whole-graph quantity and is passed explicitly
        tl = 2.0 * np.bincount(links.ravel().astype(np.int64), minlength=n).astype(np.float64)
        mine = ((links[:, 0] >= nb) & (links[:, 0] < ne)) | ((links[:, 1] >= nb) & (links[:, 1] < ne))
        factory = engine_factory or CudaShardEngine
        # `ones` = the reference's _network.ones(): ALL links incl. held-out ones (numerator of the annealing rescale,
        # src/linksampling.cc:541-542); defaults to the training-link count
        self.eng = factory(n, k, links[mine], (nb, ne), device, stream, tl=tl,
                           ones=(links.shape[0] if ones is None else ones), **kw)
        self.nlinks = links.shape[0]
        self.local_half_edges = int(tl[nb:ne].sum() // 2)
        self._buf = {name: self.eng.buffer(name) for name in
                     ("exppi", "mphi", "gamma", "kvec", "active", "active_bits", "member_bits")}

    def _init_peer(self, n, k, links, rank, world, device, stream, ones, chunks, share_gamma, kw):
        """Product path: one CUDA engine per rank, arenas exchanged as CUDA IPC handles, svi_ls_mg_step."""
        if not torch.cuda.is_available():
            raise RuntimeError("svinet_b200.sharded: no CUDA device; the product path has no CPU fallback")
        from .engine import LinkSamplingEngine
        links = np.ascontiguousarray(links, dtype=np.uint32).reshape(-1, 2)
        self.bounds = plan_shards(n, links, world)
        nb, ne = int(self.bounds[rank]), int(self.bounds[rank + 1])
        # the WHOLE link list goes to the library: the device-side CSR build keeps the half-edges of this block
        self.eng = LinkSamplingEngine(n, k, links, device=device, node_range=(nb, ne), stream=stream,
                                      ones=(links.shape[0] if ones is None else ones), **kw)
        self.nlinks = links.shape[0]
        self.local_half_edges = int(self.eng.info()["half_edges_phi"])
        blob = torch.from_numpy(self.eng.peer_blob())
        if world > 1:
            dev = torch.device("cuda", device)
            mine = blob.to(dev)
            allb = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allb, mine, group=self.group)
            blobs = torch.stack(allb).cpu().numpy()
        else:
            blobs = blob.numpy()[None]
        self.eng.peer_attach(world, rank, self.bounds, blobs, chunks=chunks)
        if share_gamma:
            self.eng.mg_share_gamma(True)
        self.share_gamma = share_gamma

    # ---- collectives on the engine's own buffers ----
    def _allreduce(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _allgather_rows(self, name, bulk=False):
        """Every rank publishes its own row block of the (replicated-layout) buffer `name`, in place.

        The blocks are uneven (edge-balanced), so this is an all-gather-v.  Over NCCL it is ONE group of
        point-to-point transfers (every rank sends its block to every peer and receives theirs), so all
        NVLink ports work at once; a sequence of `world` broadcasts (the gloo path of the CPU tests)
        serialises them and measured 2x slower at 8 GPUs."""
        # (`converged` is double-buffered inside the library: ask for the current pointer every time)
        buf = self._buf[name] if name in self._buf else self.eng.buffer(name)
        group = self.bulk_group if (bulk and self.bulk_group is not None) else self.group
        blocks = [buf[int(self.bounds[r]):int(self.bounds[r + 1])] for r in range(self.world)]
        if dist.get_backend(self.group) == "nccl" and self.world > 1:
            mine = blocks[self.rank]
            ops = []
            for d in range(1, self.world):
                dst, src = (self.rank + d) % self.world, (self.rank - d) % self.world
                if mine.shape[0]:
                    ops.append(dist.P2POp(dist.isend, mine, self._global(dst), group=group))
                if blocks[src].shape[0]:
                    ops.append(dist.P2POp(dist.irecv, blocks[src], self._global(src), group=group))
            for w in (dist.batch_isend_irecv(ops) if ops else []):
                w.wait()
            return
        for r in range(self.world):
            if blocks[r].shape[0]:
                dist.broadcast(blocks[r], src=self._global(r), group=group)

    def _global(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def set_state(self, gamma, lam):
        self.eng.set_state(gamma, lam)     # derives the factors of ALL rows, no exchange needed

    def step(self, it, annealing, write_comm, events=None, stream=None):
        if self.exchange == "peer":
            return self.eng.mg_step(it, annealing, write_comm)
        if self.overlap:
            return self._step_overlapped(it, annealing, write_comm, events, stream)

        def mark(i):
            if events is not None:
                events[i].record(stream) if stream is not None else events[i].record()
        kv = self._buf["kvec"]
        self.eng.phase_phi(it, write_comm)
        mark(1)
        self.eng.phase_node()
        self._allreduce(kv[0:3])
        self._allgather_rows("mphi")
        mark(2)
        self.eng.phase_s3()
        self._allreduce(kv[3:4])
        mark(3)
        self.eng.phase_finish(annealing)
        self._allgather_rows("exppi")
        self._allgather_rows("converged")
        if it >= 1000:                     # the active-set branch reads neighbours' masks (:634)
            self._allgather_rows("active")
            self._allgather_rows("active_bits")

    def _step_overlapped(self, it, annealing, write_comm, events=None, stream=None):
        """Same iteration with the refresh moved in front of the s3 sweep (svi_ls_phase_refresh /
        svi_ls_phase_lambda), so that the bulk exchanges run beside compute on a side stream and their own
        communicator: the mphi gather beside the refresh, the exp(Elogpi) / converged gather beside the s3
        sweep.  Only the mphi gather stays exposed."""
        def mark(i):
            if events is not None:
                events[i].record(stream) if stream is not None else events[i].record()
        cuda = self.side is not None
        main = torch.cuda.current_stream() if cuda else None
        kv = self._buf["kvec"]

        def on_side(after, fn):
            """run fn() on the side stream once `after` (an event of the main stream) has happened"""
            if not cuda:
                fn()
                return None
            with torch.cuda.stream(self.side):
                self.side.wait_event(after)
                fn()
                done = torch.cuda.Event()
                done.record(self.side)
            return done

        def event():
            if not cuda:
                return None
            e = torch.cuda.Event()
            e.record(main)
            return e

        self.eng.phase_phi(it, write_comm)
        mark(1)
        self.eng.phase_node()
        self._allreduce(kv[0:3])
        got_mphi = on_side(event(), lambda: self._allgather_rows("mphi", bulk=True))
        self.eng.phase_refresh(annealing)

        def publish():
            self._allgather_rows("exppi", bulk=True)
            self._allgather_rows("converged", bulk=True)
            if it >= 1000:
                self._allgather_rows("active", bulk=True)
                self._allgather_rows("active_bits", bulk=True)
        refreshed = event()
        if cuda:
            main.wait_event(got_mphi)
        mark(2)
        self.eng.phase_s3()
        published = on_side(refreshed, publish)
        self._allreduce(kv[3:4])
        mark(3)
        self.eng.phase_lambda(annealing)
        if cuda:
            main.wait_event(published)

    def gather_state(self):
        """Full gamma [n,k] and lambda [k,2] on every rank (for save_model / parity checks)."""
        if self.exchange == "peer":
            if not self.share_gamma:
                self.eng.mg_publish_gamma()        # collective: every rank pushes its rows to every peer
            return self.eng.get_state()
        self._allgather_rows("gamma")
        return self.eng.get_state()

    def gather_membership(self):
        self._allgather_rows("member_bits")
        return self.eng.membership()

    def phase_ms(self, ev):
        import numpy as _np
        return [float(_np.mean([e[i].elapsed_time(e[i + 1]) for e in ev])) for i in range(4)]

    def e2e(self, step_fn, it0, steps, nlinks, unit, heldout=None):
        """The step driven the way a reference-facing caller drives it, with HOST buffers every iteration: the
        sharded iteration, a gamma row all-gather, the held-out likelihood of this rank's slice of the
        validation pairs (host pair lists in, host log-likelihoods out) and the link-community membership bits
        (device -> host).  Bytes are per rank."""
        import time
        hp, hq, hy = heldout if heldout is not None else (np.zeros(0, np.uint32),) * 2 + (np.zeros(0, np.uint8),)
        if self.exchange == "peer":
            # every pair is evaluated by a shard that owns one of its endpoints (which one: the parity rule of the s3
            # ownership, so that the pairs spread evenly): only the other row is a peer load
            key = np.where(((hp ^ hq) & 1) == 1, hp, hq)
            sl = (np.searchsorted(self.bounds, key, side="right") - 1) == self.rank
        else:
            sl = slice(self.rank, None, self.world)
        hp, hq, hy = np.ascontiguousarray(hp[sl]), np.ascontiguousarray(hq[sl]), np.ascontiguousarray(hy[sl])
        it = it0
        ll = bits = np.zeros(0)
        for timed in (False, True):
            dist.barrier(group=self.group)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(1 if not timed else steps):
                step_fn(it)
                it += 1
                if self.exchange == "peer":
                    # held-out pairs: rows of other shards are peer loads inside svi_ls_heldout; only this rank's
                    # block of the membership words is read back
                    ll = self.eng.heldout(hp, hq, hy)
                    nb, ne = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
                    bits = self.eng.membership_rows(nb, ne - nb)
                else:
                    self._allgather_rows("gamma")
                    ll = self.eng.heldout(hp, hq, hy)
                    bits = self.eng.membership_bits()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return {"value": nlinks * steps / float(t.item()), "unit": unit, "steps": steps,
                "h2d_bytes_per_step": int(hp.nbytes + hq.nbytes + hy.nbytes),
                "d2h_bytes_per_step": int(ll.nbytes + bits.nbytes),
                "what": ("per rank and iteration: svi_ls_mg_step (peer-memory exchanges) + "
                         "svi_ls_heldout on the pairs one of whose endpoints the rank owns (host in/out; the other row is a "
                         "peer load) + svi_ls_get_membership_rows of the rank's own block (host bits)") if self.exchange == "peer" else
                        ("per rank and iteration: sharded step (NCCL exchanges) + gamma row all-gather + "
                         "svi_ls_heldout on 1/world of the pairs (host in/out) + svi_ls_get_membership (host bits)")}
