#!/usr/bin/env python
"""binary=True on the configs[1]-shaped URM (Jaccard item-item, k=100): hot-kernel time with the counting form of the panel
(native integer shared-memory adds) against the float adds.  usage: python scripts/bench_binary.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import similaripy_b200 as sim
from similaripy_b200 import _engine

dev = torch.device("cuda", 0)
ip, ix, dv = bench.gen_urm_device(1_000_000, 200_000, 1e-3, 2, dev)
urm = sim.DeviceMatrix(_engine.DeviceCSR(1_000_000, 200_000, ip, ix, dv, sorted_rows=True), False)
out = {}
for label, tuning in (("float adds", dict(unit_values=False)), ("integer adds", None)):
    job = _engine.prepare_job(urm.T, None, l1=1.0, t1=1.0, t2=1.0, k=100, binary=True, verbose=False, device=0, tuning=tuning)
    ms = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); job.run(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out[label] = (float(np.median(ms[1:])), job.out_vals.clone(), job.out_counts.clone(), int(job.args.engine), int(job.args.group))
    print(json.dumps({"jaccard binary=True, URM 1M x 200k d=1e-3, k=100": label, "kernel_ms": round(out[label][0], 3),
                      "engine": out[label][3], "drain_warps": out[label][4]}), flush=True)
a, b = out["float adds"], out["integer adds"]
same = bool(torch.equal(a[2], b[2])) and bool(torch.equal(torch.sort(a[1].view(-1, 100), dim=1).values, torch.sort(b[1].view(-1, 100), dim=1).values))
print(json.dumps({"values identical": same, "speed-up": round(a[0] / b[0], 3)}))
