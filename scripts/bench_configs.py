#!/usr/bin/env python
"""Hot-kernel throughput on the other BASELINE.json configurations (bench.py measures configs[1]): full-size
synthetic operands generated on the device, a uniform sample of the target rows (rows are independent),
CUDA-event time of the hot kernel, algorithmic bytes -> fraction of the measured HBM peak.
usage: python scripts/bench_configs.py [cfg3 cfg4 cfg5 ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import similaripy_b200 as sim
from similaripy_b200 import _engine

dev = torch.device("cuda", 0)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def gen(n_rows, n_cols, density, seed):
    ip, ix, dv = bench.gen_urm_device(n_rows, n_cols, density, seed, dev)
    return sim.DeviceMatrix(_engine.DeviceCSR(n_rows, n_cols, ip, ix, dv, sorted_rows=True), False)


def measure(name, job, n_rows_total, reps=3):
    A, B = job.A, job.B
    b_len = (B.indptr[1:] - B.indptr[:-1]).to(torch.int64)
    t = job.targets.long()
    a_len = (A.indptr[1:] - A.indptr[:-1]).to(torch.int64)
    cum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(b_len[A.indices.long()], 0)])
    products = int((cum[A.indptr[t + 1].long()] - cum[A.indptr[t].long()]).sum())
    nnz_a = int(a_len[t].sum())
    ms = []
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); job.run(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    k_ms = float(np.median(ms[1:]))
    out_nnz = int(job.out_counts.sum())
    alg = 8 * job.n_targets + 16 * nnz_a + 8 * products + 8 * out_nnz
    line = {"config": name, "target_rows_sampled": job.n_targets, "target_rows_total": n_rows_total, "k": job.k,
            "products": products, "kernel_ms": round(k_ms, 3), "rows_per_s": round(job.n_targets / k_ms * 1e3, 1),
            "gproducts_per_s": round(products / k_ms / 1e6, 1), "out_nnz_per_s": round(out_nnz / k_ms * 1e3, 1),
            "achieved_gbs": round(alg / k_ms / 1e6, 1), "frac_of_measured_hbm_peak": round(alg / k_ms / 1e6 / peak, 4),
            "full_job_estimate_s": round(n_rows_total / (job.n_targets / k_ms * 1e3), 1),
            "plan": {"n_panels": int(job.args.n_panels), "panel_width": int(job.args.panel_width), "threads": int(job.args.threads),
                     "group": int(job.args.group)}}
    print(json.dumps(line), flush=True)


def sample(n, m, seed):
    return np.sort(np.random.default_rng(seed).choice(n, size=m, replace=False)).astype(np.int32)


which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
tuning = {k: int(v) for k, v in (kv.split("=") for kv in os.environ["SPY_TUNING"].split(","))} if os.environ.get("SPY_TUNING") else None
if "cfg3" in which:  # s_plus(X, k=200, shrink=10), X 500k x 500k d=2e-3
    x = gen(500_000, 500_000, 2e-3, 3)
    job = _engine.prepare_job(x, None, k=200, target_rows=sample(500_000, 20_000, 3), verbose=False, device=0, tuning=tuning,
                              l1=0.5, l2=0.5, t1=1.0, t2=1.0, c1=0.5, c2=0.5, stabilized_shrink=10.0)
    measure("configs[2]: s_plus k=200 shrink=10, 500k x 500k d=2e-3", job, 500_000)
    del x, job; torch.cuda.empty_cache()
if "cfg4" in which:  # rp3beta(URM.T, alpha=1, beta=0.6, k=100), URM 2M x 500k d=5e-4
    urm = gen(2_000_000, 500_000, 5e-4, 4)
    pop = _engine.axis_sum(urm, 0)
    job = _engine.prepare_job(sim.normalize(urm.T, norm="l1", axis=1), sim.normalize(urm, norm="l1", axis=1), k=100,
                              target_rows=sample(500_000, 60_000, 4), verbose=False, device=0, tuning=tuning, weight_depop_matrix2=pop, p2=0.6, l3=1.0)
    measure("configs[3]: rp3beta beta=0.6 k=100 item-item, URM 2M x 500k d=5e-4 (per GPU; rows shard over 8)", job, 500_000)
    del urm, pop, job; torch.cuda.empty_cache()
if "cfg5" in which:  # dot_product(URM, S.T, k=100, filter_cols=URM), URM 5M x 200k d=1e-3, S ~100 neighbours per item
    urm = gen(5_000_000, 200_000, 1e-3, 5)
    s_t = gen(200_000, 200_000, 5e-4, 55)
    job = _engine.prepare_job(urm, s_t, k=100, target_rows=sample(5_000_000, 500_000, 5), filter_cols=urm, verbose=False, device=0, tuning=tuning)
    measure("configs[4]: dot_product URM x S.T filter_cols=URM k=100, URM 5M x 200k d=1e-3 (per GPU; rows shard over 8)", job, 5_000_000)
