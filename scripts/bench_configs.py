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


def to_scipy(dm):
    import scipy.sparse as sp
    st = dm.stored
    arrs = (st.data.cpu().numpy(), st.indices.cpu().numpy(), st.indptr.cpu().numpy())
    return (sp.csc_array if dm.transposed else sp.csr_array)(arrs, shape=dm.shape)


def reference_rows_per_s(name, call, rows, n_small=8):
    """The unmodified reference (oracle/_ref) on the host cores: wall(n) = T_fixed + n * t_row, T_fixed from a call with
    `n_small` target rows (its O(nnz) pre-processing does not depend on the number of target rows)."""
    try:
        from oracle import ref_api
        if not ref_api.available():
            return None
        t0 = time.perf_counter(); call(ref_api, rows[:n_small]); t_fixed = time.perf_counter() - t0
        t0 = time.perf_counter(); call(ref_api, rows); t_all = time.perf_counter() - t0
        t_row = max(t_all - t_fixed, 1e-9) / (len(rows) - n_small)
        out = {"config": name, "impl": "reference (oracle/_ref, OpenMP)", "host_threads": ref_api.num_threads(), "target_rows_timed": len(rows),
               "t_fixed_s": round(t_fixed, 2), "t_row_ms": round(t_row * 1e3, 4), "rows_per_s_excluding_fixed": round(1.0 / t_row, 1)}
        print(json.dumps(out), flush=True)
    except Exception as exc:
        print(json.dumps({"config": name, "impl": "reference", "error": repr(exc)[:300]}), flush=True)


WITH_REF = os.environ.get("SPY_WITH_REFERENCE") == "1"
import time


def sample(n, m, seed):
    return np.sort(np.random.default_rng(seed).choice(n, size=m, replace=False)).astype(np.int32)


which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
tuning = {k: int(v) for k, v in (kv.split("=") for kv in os.environ["SPY_TUNING"].split(","))} if os.environ.get("SPY_TUNING") else None
if "cfg3" in which:  # s_plus(X, k=200, shrink=10), X 500k x 500k d=2e-3
    x = gen(500_000, 500_000, 2e-3, 3)
    job = _engine.prepare_job(x, None, k=200, target_rows=sample(500_000, 20_000, 3), verbose=False, device=0, tuning=tuning,
                              l1=0.5, l2=0.5, t1=1.0, t2=1.0, c1=0.5, c2=0.5, stabilized_shrink=10.0)
    measure("configs[2]: s_plus k=200 shrink=10, 500k x 500k d=2e-3", job, 500_000)
    if WITH_REF:
        xh = to_scipy(x)
        reference_rows_per_s("configs[2]", lambda api, r: api.similarity("s_plus", xh, k=200, shrink=10.0, target_rows=r, format_output="csr",
                                                                        verbose=False), sample(500_000, 1_500, 33))
        del xh
    del x, job; torch.cuda.empty_cache()
if "cfg4" in which:  # rp3beta(URM.T, alpha=1, beta=0.6, k=100), URM 2M x 500k d=5e-4
    urm = gen(2_000_000, 500_000, 5e-4, 4)
    pop = _engine.axis_sum(urm, 0)
    job = _engine.prepare_job(sim.normalize(urm.T, norm="l1", axis=1), sim.normalize(urm, norm="l1", axis=1), k=100,
                              target_rows=sample(500_000, 60_000, 4), verbose=False, device=0, tuning=tuning, weight_depop_matrix2=pop, p2=0.6, l3=1.0)
    measure("configs[3]: rp3beta beta=0.6 k=100 item-item, URM 2M x 500k d=5e-4 (per GPU; rows shard over 8)", job, 500_000)
    if WITH_REF:
        uh = to_scipy(urm)
        uht = uh.T.tocsr()
        reference_rows_per_s("configs[3]", lambda api, r: api.similarity("rp3beta", uht, alpha=1.0, beta=0.6, k=100, target_rows=r,
                                                                        format_output="csr", verbose=False), sample(500_000, 40_000, 44))
        del uh, uht
    del urm, pop, job; torch.cuda.empty_cache()
if "cfg5" in which:  # dot_product(URM, S.T, k=100, filter_cols=URM), URM 5M x 200k d=1e-3, S ~100 neighbours per item
    urm = gen(5_000_000, 200_000, 1e-3, 5)
    s_t = gen(200_000, 200_000, 5e-4, 55)
    job = _engine.prepare_job(urm, s_t, k=100, target_rows=sample(5_000_000, 500_000, 5), filter_cols=urm, verbose=False, device=0, tuning=tuning)
    measure("configs[4]: dot_product URM x S.T filter_cols=URM k=100, URM 5M x 200k d=1e-3 (per GPU; rows shard over 8)", job, 5_000_000)
    if WITH_REF:
        uh, sh = to_scipy(urm), to_scipy(s_t)
        reference_rows_per_s("configs[4]", lambda api, r: api.similarity("dot_product", uh, sh, k=100, target_rows=r, filter_cols=uh,
                                                                        format_output="csr", verbose=False), sample(5_000_000, 40_000, 55))
