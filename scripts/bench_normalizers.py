#!/usr/bin/env python
"""Throughput of the in-place CSR normalizers on the cfg2 URM (1M x 200k, 2e8 nnz, float32 / int32) resident in HBM,
next to the reference's serial Cython loops on the host (oracle/_ref) on a 1/10 sample of the rows.
Algorithmic bytes: nnz * (2 * sizeof(value) + sizeof(index)) per pass over the matrix."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scipy.sparse as sp
import torch
import bench
import similaripy_b200 as sim
from similaripy_b200 import _engine

dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
ip, ix, dv = bench.gen_urm_device(1_000_000, 200_000, 1e-3, 2, dev)
nnz = ix.numel()
cases = [("normalize l1", lambda m: sim.normalize(m, norm="l1", inplace=True), 1, False),
         ("normalize l2", lambda m: sim.normalize(m, norm="l2", inplace=True), 1, False),
         ("normalize max", lambda m: sim.normalize(m, norm="max", inplace=True), 1, False),
         ("tfidf", lambda m: sim.tfidf(m, inplace=True), 2, True),
         ("bm25", lambda m: sim.bm25(m, inplace=True), 2, True),
         ("bm25plus", lambda m: sim.bm25plus(m, inplace=True), 2, True)]
try:
    from oracle import ref_api
    have_ref = ref_api.available()
except Exception:
    have_ref = False
host = None
if have_ref:
    n_s = 100_000
    e = int(ip[n_s])
    host = sp.csr_array((dv[:e].cpu().numpy(), ix[:e].cpu().numpy(), ip[:n_s + 1].cpu().numpy()), shape=(n_s, 200_000))
for name, fn, passes, uses_idx in cases:
    ms = []
    for _ in range(4):
        m = sim.DeviceMatrix(_engine.DeviceCSR(1_000_000, 200_000, ip, ix, dv.clone(), sorted_rows=True), False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(m); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms[1:]))
    alg = nnz * ((2 * 4) * passes + (4 * passes if uses_idx else 0)) + (8 * 1_000_000)
    line = {"op": name, "nnz": nnz, "ms": round(t, 3), "gnnz_per_s": round(nnz / t / 1e6, 2), "algorithmic_gbs": round(alg / t / 1e6, 1),
            "frac_of_measured_hbm_peak": round(alg / t / 1e6 / peak, 3)}
    if host is not None:
        key = name.split()[-1] if name.startswith("normalize") else name
        h = host.copy()
        t0 = time.perf_counter()
        if name.startswith("normalize"):
            ref_api.normalize(h, norm=key, inplace=True)
        else:
            getattr(ref_api, key)(h, inplace=True)
        dt = time.perf_counter() - t0
        line["reference_cpu_gnnz_per_s"] = round(h.nnz / dt / 1e9, 4)
        line["reference_sample"] = f"{h.shape[0]} rows, {h.nnz} nnz, serial Cython loop (oracle/_ref)"
    print(json.dumps(line), flush=True)
