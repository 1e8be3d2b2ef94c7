#!/usr/bin/env python
"""Per-phase cycle shares of the hot kernel on the cfg2 bench workload (needs a library built with
-DSPY_PHASE_TIMING=1, selected through SIMILARIPY_B200_LIB).  usage: python scripts/phase_timing.py [scale] [tuning k=v,...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import similaripy_b200 as sim
from similaripy_b200 import _engine

cfg5 = len(sys.argv) > 1 and sys.argv[1] == "cfg5"
scale = 1.0 if cfg5 else (float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
tuning = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[2].split(","))} if len(sys.argv) > 2 else None
dev = torch.device("cuda", 0)
n_users, n_items = int(1_000_000 * scale), int(200_000 * scale)
density = 1e-3 if scale == 1.0 else min(0.5, 1e-3 / scale ** 0.5)
ip, ix, dv = bench.gen_urm_device(n_users, n_items, density, 2, dev)
urm = sim.DeviceMatrix(_engine.DeviceCSR(n_users, n_items, ip, ix, dv, sorted_rows=True), False)
urm = sim.bm25(urm, inplace=True) if not cfg5 else None
if cfg5:
    import numpy as np
    urm = None
    ip, ix, dv = bench.gen_urm_device(5_000_000, 200_000, 1e-3, 5, dev)
    u5 = sim.DeviceMatrix(_engine.DeviceCSR(5_000_000, 200_000, ip, ix, dv, sorted_rows=True), False)
    ip, ix, dv = bench.gen_urm_device(200_000, 200_000, 5e-4, 55, dev)
    st = sim.DeviceMatrix(_engine.DeviceCSR(200_000, 200_000, ip, ix, dv, sorted_rows=True), False)
    rows = np.sort(np.random.default_rng(5).choice(5_000_000, size=200_000, replace=False)).astype(np.int32)
    job = _engine.prepare_job(u5, st, k=100, target_rows=rows, filter_cols=u5, verbose=False, device=0, tuning=tuning)
else:
    job = _engine.prepare_job(urm.T, None, l2=1.0, c1=0.5, c2=0.5, k=100, verbose=False, device=0, tuning=tuning)
for it in range(2):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); job.run(); ev1.record(); torch.cuda.synchronize()
    ph = job.scratch[128:256].view(torch.int64).cpu().tolist()
names = ["stage", "accumulate (wall, to barrier)", "accumulate (mean warp busy)", "drain barrier wait, panel 0", "evaluate", "tighten",
         "final select + write", "row fetch / other", "drain barrier wait, panels >= 1", "#drain passes panel 0 (x CTAs)", "#drain passes panels >= 1",
         "drain scan (thread 0), panel 0", "drain scan (thread 0), panels >= 1", "post-accumulate (next-row loads issue)"]
wall = [ph[i] for i in (0, 1, 3, 4, 5, 6, 7, 8, 11, 12, 13)]
tot = sum(wall)
print(f"kernel {ev0.elapsed_time(ev1):.1f} ms; plan panels={job.args.n_panels} W={job.args.panel_width} threads={job.args.threads} group={job.args.group}")
for i, (n, v) in enumerate(zip(names, ph)):
    if i in (9, 10):
        print(f"  {n:40s} {v} passes = {v / (200000 * scale):.2f} per row")
    else:
        print(f"  {n:40s} {100.0 * v / tot:6.2f} %")
