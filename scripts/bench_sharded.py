#!/usr/bin/env python
"""Strong scaling of ONE similarity call over the GPUs of a box: configs[1] (cosine item-item, 200k target
rows) with the target rows cut into work-balanced ranges, with and without the all-gather of the output slab.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_sharded.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import similaripy_b200 as sim
from similaripy_b200 import _engine, sharded

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
dev = torch.device("cuda", local)
ip, ix, dv = bench.gen_urm_device(1_000_000, 200_000, 1e-3, 2, dev)  # every rank generates the same URM (replicated operands)
urm = sim.bm25(sim.DeviceMatrix(_engine.DeviceCSR(1_000_000, 200_000, ip, ix, dv, sorted_rows=True), False), inplace=True)


def timed(gather, steps=3, warmup=2):
    for _ in range(warmup):
        with sharded.shard_rows(gather=gather):
            r = sim.cosine(urm.T, k=100, verbose=False, on_device=True, device=local)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        with sharded.shard_rows(gather=gather):
            r = sim.cosine(urm.T, k=100, verbose=False, on_device=True, device=local)
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t), r


ms_local, r_local = timed(False)
ms_gather, r_full = timed(True)
nnz_local = torch.tensor([r_local.nnz], device=dev); dist.all_reduce(nnz_local)
if rank == 0:
    print(json.dumps({"workload": "configs[1] cosine item-item k=100, 200k target rows, ONE call sharded by work", "n_gpus": world,
                      "ms_per_call_local_rows_only": round(ms_local, 2), "rows_per_s_local": round(200_000 / ms_local * 1e3, 1),
                      "ms_per_call_with_all_gather": round(ms_gather, 2), "rows_per_s_gathered": round(200_000 / ms_gather * 1e3, 1),
                      "all_gather_ms": round(ms_gather - ms_local, 2), "gathered_slab_bytes": 200_000 * 100 * 8 + 200_000 * 4,
                      "out_nnz_total": int(nnz_local.item()), "out_nnz_gathered": int(r_full.nnz)}), flush=True)
dist.destroy_process_group()
