#!/usr/bin/env python
"""Where the end-to-end call on host buffers spends its time (configs[1], pinned scipy CSC in, scipy CSR out): the phases of
_engine.prepare_job / s_plus wrapped with a device synchronisation on both sides (so phases that normally overlap are
serialised: the sum is an upper bound of the real call, printed beside it).  usage: python scripts/e2e_breakdown.py"""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import similaripy_b200 as sim
from similaripy_b200 import _engine

dev = torch.device("cuda", 0)
ip, ix, dv = bench.gen_urm_device(1_000_000, 200_000, 1e-3, 2, dev)
urm = sim.bm25(sim.DeviceMatrix(_engine.DeviceCSR(1_000_000, 200_000, ip, ix, dv, sorted_rows=True), False), inplace=True)
h = [bench.pinned_numpy(t) for t in (urm.stored.indptr, urm.stored.indices, urm.stored.data)]
m = bench.host_csr_views(h[0], h[1], h[2], (1_000_000, 200_000))
del urm, ip, ix, dv
torch.cuda.empty_cache()
acc = collections.OrderedDict()


def timed(owner, name, label=None):
    fn = getattr(owner, name)
    def wrapper(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); acc[label or name] = acc.get(label or name, 0.0) + (time.perf_counter() - t0) * 1e3
        return out
    setattr(owner, name, wrapper)


call = lambda: sim.cosine(m.T, None, k=100, verbose=False, format_output="csr")
for _ in range(2):
    r = call()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    r = call()
torch.cuda.synchronize(); plain = (time.perf_counter() - t0) / 3 * 1e3
for owner, name in ((_engine, "upload_stored"), (_engine, "transpose_csr"), (_engine.KnnJob, "build_vectors"), (_engine.KnnJob, "build_selectors"),
                    (_engine.KnnJob, "plan"), (_engine.KnnJob, "run"), (_engine.KnnJob, "assemble_device"), (_engine.KnnJob, "to_host")):
    timed(owner, name)
timed(_engine.Ctx, "h2d", "  (h2d inside upload_stored)")
timed(_engine, "filter_csr", "  (filter_csr inside upload_stored)")
n = 3
for _ in range(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = call()
    torch.cuda.synchronize(); acc["TOTAL (serialised)"] = acc.get("TOTAL (serialised)", 0.0) + (time.perf_counter() - t0) * 1e3
print(f"plain call: {plain:.1f} ms per call")
for k, v in acc.items():
    print(f"{k:40s} {v / n:8.2f} ms")
