#!/usr/bin/env python3
"""Development probe (torchrun): per-phase times of svi_ls_mg_step for several exchange configurations on one graph.

    torchrun --nproc-per-node N tools/probe_mg.py c4 push=sm,chunks=3,ratio=0.7,blocks=32 push=ce,chunks=4 ...
"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import WORKLOADS, fast_state
from svinet_b200 import synth, sharded
from svinet_b200.sharded import ShardedLinkSampling

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
wl = sys.argv[1]
n, k, target = WORKLOADS[wl]
links = synth.mmsb_links(n, k, target, seed=1234, device=str(dev))
torch.cuda.empty_cache()
g0, l0 = fast_state(n, k, links)
bounds = sharded.plan_shards(n, links, world)
sharded.plan_shards = lambda *a, **kw: bounds          # computed once
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
for spec in sys.argv[2:]:
    cfg = dict(kv.split("=") for kv in spec.split(","))
    os.environ["SVI_LS_MG_PUSH"] = cfg.get("push", "sm")
    os.environ["SVI_LS_MG_CHUNK_RATIO"] = cfg.get("ratio", "0.7")
    os.environ["SVI_LS_MG_PUSH_BLOCKS"] = cfg.get("blocks", "32")
    r = ShardedLinkSampling(n, k, links, rank=rank, world=world, device=lr, stream=stream.cuda_stream,
                            chunks=int(cfg.get("chunks", 4)))
    if cfg.get("gamma", "0") == "1":
        r.eng.mg_share_gamma(True)
    r.set_state(g0, l0)
    it = 0
    for _ in range(3):
        r.step(it, True, it > 0); it += 1
    r.eng.sync(); dist.barrier(); torch.cuda.synchronize()
    r.eng.mg_timing(True)
    steps = int(cfg.get("steps", 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        r.step(it, True, it > 0); it += 1
    e1.record(stream)
    r.eng.sync(); dist.barrier(); torch.cuda.synchronize()
    ms, _ = r.eng.mg_timing(False, read=True)
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    tmin = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    allms = [None] * world
    dist.all_gather_object(allms, ms)
    if rank == 0:
        worst = {key: max(m[key] for m in allms) for key in ms}
        print(json.dumps({"spec": spec, "world": world, "ms_per_step_max": float(t), "ms_per_step_min": float(tmin),
                          "rank0": {a: round(b, 3) for a, b in ms.items()}, "max_over_ranks": {a: round(b, 3) for a, b in worst.items()}}), flush=True)
    dist.barrier()
    r.eng.close()
    del r
    torch.cuda.empty_cache()
dist.barrier()
dist.destroy_process_group()
