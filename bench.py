#!/usr/bin/env python
"""bench.py -- similarity rows/sec of the sparse-KNN hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale F]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (N=1): BASELINE.json configs[1] -- item-item ``cosine(URM.T, k=100)`` on a BM25-normalised
synthetic URM of 1M users x 200k items, density 1e-3 (per-row Binomial nnz, unique uniform
sorted columns, float32 values, seed 2).  One "step" = one complete similarity call over all 200k target
rows: transpose, norm vectors, panel split points, the fused expand/accumulate/similarity/top-k kernel and
the CSR assembly of the result.

  value  rows/s with the inputs already resident in HBM (DeviceMatrix in, DeviceMatrix out);
  e2e    rows/s through the public drop-in call ``similaripy_b200.cosine(scipy_matrix, ...)`` with HOST
         buffers (pinned): the H2D copy of the CSR and the D2H copy of the result are inside the timed region;
  roofline  algorithmic bytes of the hot kernel / its CUDA-event duration against MEASURED_PEAKS.json hbm_gbs;
  cpu_baseline  the compiled, unmodified reference (oracle/_ref, OpenMP, all host cores) on a bounded sample
         of the same target rows, extrapolated to the full job (see ``sample``).

N>1 (weak scaling): target rows shard across ranks with no data-path collective; B (=URM) is replicated;
rank r owns its own 200k-row shard A_r (rank 0's is URM.T itself, rank r>0's is a fresh draw of the same
shape, i.e. 200k further items scored against the same catalogue).  value = all ranks' rows / max-over-ranks time.

``--impl reference`` times the reference's own CPU implementation (same metric / config) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import scipy.sparse as sp

METRIC = "similarity rows/sec (cosine item-item, k=100)"
UNIT = "rows/s"
K_NEIGHBOURS = 100
CFG2 = dict(n_users=1_000_000, n_items=200_000, density=1e-3, seed=2)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# synthetic data (generated on the GPU: scipy.sparse.random is far too slow at 2e8 nnz)
# ------------------------------------------------------------------------------------------------
def gen_urm_device(n_rows, n_cols, density, seed, device):
    """CSR (indptr i32, indices i32, data f32) on the device: per-row nnz ~ Binomial(n_cols, density),
    columns uniform without replacement (duplicates removed), ascending; values in (0, 1]."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cnt = torch.binomial(torch.full((n_rows,), float(n_cols), device=device),
                         torch.full((n_rows,), float(density), device=device), generator=g).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(n_rows, device=device, dtype=torch.int64), cnt)
    cols = torch.randint(0, n_cols, (rows.numel(),), device=device, dtype=torch.int64, generator=g)
    key = torch.unique(rows * n_cols + cols, sorted=True)
    del rows, cols
    r = torch.div(key, n_cols, rounding_mode="floor")
    indices = (key - r * n_cols).to(torch.int32)
    del key
    indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.bincount(r, minlength=n_rows), 0)
    del r
    data = (1.0 - torch.rand(indices.numel(), device=device, dtype=torch.float32, generator=g))
    return indptr.to(torch.int32), indices, data


def pinned_numpy(t):
    """Device tensor -> numpy array backed by pinned host memory (kept alive by the array's base)."""
    import torch
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t)
    torch.cuda.synchronize()
    return h.numpy()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception as exc:  # nvidia-smi missing: report it, do not fail the bench
            self.proc = None
            self.error = str(exc)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


# ------------------------------------------------------------------------------------------------
# the reference on the host cores (oracle/_ref = the compiled, unmodified reference)
# ------------------------------------------------------------------------------------------------
def reference_callable():
    """(fn, kind): fn(matrix1, target_rows) -> csr result through the reference's own s_plus driver."""
    from oracle import ref_api
    if ref_api.available():
        from oracle import ref_api as impl
        kind, threads = "reference", impl.num_threads()
    else:
        from oracle import oracle as impl
        kind, threads = "port", impl.max_threads()

    def fn(matrix1, target_rows):
        return impl.similarity("cosine", matrix1, None, k=K_NEIGHBOURS, target_rows=target_rows,
                               format_output="csr", verbose=False, num_threads=0)
    return fn, kind, threads


class ReferenceTimer:
    """Full-job rows/s of the reference from bounded samples.

    wall(n) = T_fixed + n * t_row: the reference's O(nnz) per-call preprocessing (eliminate_zeros, casts,
    squared norms, output assembly) does not depend on the number of target rows, the OpenMP row loop does.
    T_fixed is measured by a call with 8 target rows, t_row from a call with `n_sample` seeded random rows;
    full-job time = T_fixed + n_rows * t_row.  matrix1 is handed over already in CSR (scipy tocsr of URM.T
    done once, outside the timing, to bound the bench; it would add to T_fixed)."""

    def __init__(self, a_csr, n_rows, seed=123):
        self.fn, self.kind, self.threads = reference_callable()
        self.a = a_csr
        self.n_rows = n_rows
        self.rng = np.random.default_rng(seed)
        self.t_fixed = None

    def _call(self, rows):
        t0 = time.perf_counter()
        res = self.fn(self.a, rows)
        return time.perf_counter() - t0, res

    def calibrate(self):
        tiny = np.sort(self.rng.choice(self.n_rows, size=min(8, self.n_rows), replace=False)).astype(np.int32)
        self.t_fixed, _ = self._call(tiny)
        probe_n = min(256, self.n_rows)
        probe = np.sort(self.rng.choice(self.n_rows, size=probe_n, replace=False)).astype(np.int32)
        t, _ = self._call(probe)
        self.t_row_est = max(t - self.t_fixed, 1e-6) / probe_n
        return self.t_fixed, self.t_row_est

    def sample_size(self, seconds):
        n = int(max(256, min(self.n_rows, seconds / self.t_row_est)))
        return n

    def step(self, n_sample):
        rows = np.sort(self.rng.choice(self.n_rows, size=n_sample, replace=False)).astype(np.int32)
        t, res = self._call(rows)
        t_row = max(t - self.t_fixed, 1e-9) / n_sample
        full = self.t_fixed + self.n_rows * t_row
        return dict(wall=t, t_row=t_row, full_job_s=full, rows_per_s=self.n_rows / full, out_nnz=int(res.nnz))


# ------------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def host_csr_views(indptr, indices, data, shape):
    """scipy csr_array over pinned host buffers without copying."""
    m = sp.csr_array((data, indices, indptr), shape=shape, copy=False)
    assert np.shares_memory(m.data, data) and np.shares_memory(m.indices, indices)
    m.has_sorted_indices = True
    return m


def run_ours(args):
    import torch
    import similaripy_b200 as sim
    from similaripy_b200 import _engine, _lib

    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    n_users = max(64, int(CFG2["n_users"] * args.scale))
    n_items = max(64, int(CFG2["n_items"] * args.scale))
    density = CFG2["density"] if args.scale == 1.0 else min(0.5, CFG2["density"] / args.scale ** 0.5)
    t_setup = time.perf_counter()

    # ---- inputs: URM (B, replicated on every rank) and this rank's target shard A_r -----------------------
    indptr, indices, data = gen_urm_device(n_users, n_items, density, CFG2["seed"], dev)
    urm = sim.DeviceMatrix(_engine.DeviceCSR(n_users, n_items, indptr, indices, data, sorted_rows=True), False)
    urm = sim.bm25(urm, inplace=True)  # BM25-normalised, as configs[1] says; outside the timed region
    if rank == 0:
        m1 = urm.T  # exactly configs[1]: cosine(URM.T)
        m2 = None
    else:  # weak scaling: 200k further item rows scored against the same catalogue
        ip, ix, dv = gen_urm_device(n_users, n_items, density, CFG2["seed"] + 1000 * rank, dev)
        shard = sim.DeviceMatrix(_engine.DeviceCSR(n_users, n_items, ip, ix, dv, sorted_rows=True), False)
        shard = sim.bm25(shard, inplace=True)
        m1, m2 = shard.T, urm
    torch.cuda.synchronize()
    nnz = urm.nnz
    n_targets = n_items
    log(f"[bench r{rank}] URM {n_users}x{n_items} nnz={nnz} generated+bm25 in {time.perf_counter() - t_setup:.1f}s")

    common = dict(k=K_NEIGHBOURS, verbose=False, format_output="csr", device=local)
    if os.environ.get("SPY_TUNING"):  # kernel experiments only, e.g. SPY_TUNING="threads=512,pairs=0"
        common["tuning"] = {k: int(v) for k, v in (kv.split("=") for kv in os.environ["SPY_TUNING"].split(","))}

    def step_device():
        return sim.cosine(m1, m2, on_device=True, **common)

    # ---- device-resident timing ------------------------------------------------------------------------
    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        res = step_device()
    out_nnz = res.nnz
    del res
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _engine.KERNEL_TRACE = []
    _lib.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res = step_device()
        del res
    ev1.record()
    barrier()
    launches = _lib.launch_count()
    trace, _engine.KERNEL_TRACE = _engine.KERNEL_TRACE, None
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    kern_ms = float(np.mean([t["start"].elapsed_time(t["end"]) for t in trace]))
    plan = {k: trace[0][k] for k in ("n_panels", "panel_width", "threads", "group")}
    if world > 1:
        t = torch.tensor([elapsed_ms, kern_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        elapsed_ms, kern_ms_max = float(t[0]), float(t[1])
    ms_per_step = elapsed_ms / args.steps
    value = world * n_targets / (ms_per_step / 1e3)

    # ---- algorithmic bytes of the hot kernel for this rank's launch ------------------------------
    A, B = (_engine.transpose_csr(_engine.Ctx(local), m1.stored), urm.stored)  # A = CSR of the target shard
    b_len = (B.indptr[1:] - B.indptr[:-1]).to(torch.int64)
    products = int(b_len[A.indices.long()].sum().item())
    alg_bytes = 8 * n_targets + 16 * A.nnz + 8 * products + 8 * out_nnz
    del A, b_len
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if peaks else "fallback 6650",
                "kernel": "knn_flat_kernel", "kernel_ms": round(kern_ms, 3), "algorithmic_bytes": alg_bytes,
                "products": products, "gproducts_per_s": round(products / (kern_ms / 1e3) / 1e9, 1),
                "kernel_share_of_step": round(kern_ms / ms_per_step, 4), "plan": plan}
    traffic_file = os.path.join(ROOT, "profiles", "knn_traffic.json")
    if os.path.exists(traffic_file):  # dram bytes per launch from the committed ncu --set full capture
        try:
            tj = json.load(open(traffic_file))
            if tj.get("workload_nnz") == nnz:
                roofline["traffic"] = tj.get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---- end to end: host buffers in, host result out ---------------------------------------------------------
    e2e = None
    cpu = None
    if not args.no_e2e:
        src = m1.stored  # CSR of (shard) URM; matrix1 = its transpose, as a scipy CSC view -- no host conversion
        h = [pinned_numpy(t) for t in (src.indptr, src.indices, src.data)]
        urm_host = host_csr_views(h[0], h[1], h[2], (n_users, n_items))
        if m2 is not None:
            hb = [pinned_numpy(t) for t in (urm.stored.indptr, urm.stored.indices, urm.stored.data)]
            b_host = host_csr_views(hb[0], hb[1], hb[2], (n_users, n_items))
        else:
            b_host = None
        h2d = sum(x.nbytes for x in h) + (sum(x.nbytes for x in hb) if b_host is not None else 0)

        def step_host():
            return sim.cosine(urm_host.T, b_host, **common)

        for _ in range(max(1, min(args.warmup, 2))):
            r = step_host()
        d2h = r.data.nbytes + r.indices.nbytes + r.indptr.nbytes
        e_steps = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(e_steps):
            r = step_host()
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
        e_ms = max(ev0.elapsed_time(ev1), wall * 1e3) / e_steps
        if world > 1:
            t = torch.tensor([e_ms], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            e_ms = float(t[0])
        e2e = {"value": round(world * n_targets / (e_ms / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": round(e_ms, 2), "steps": e_steps,
               "call": "similaripy_b200.cosine(scipy csc (pinned), k=100, format_output='csr') -> scipy csr"}

        # ---- CPU baseline (rank 0, N=1 only) ------------------------------------------------------------------
        if world == 1 and not args.no_cpu:
            try:
                t0 = time.perf_counter()
                a_csr = urm_host.T.tocsr()
                t_tocsr = time.perf_counter() - t0
                rt = ReferenceTimer(a_csr, n_targets)
                t_fixed, t_row = rt.calibrate()
                n_s = rt.sample_size(args.cpu_seconds)
                st = rt.step(n_s)
                cpu = {"value": round(st["rows_per_s"], 1), "unit": UNIT, "cores": rt.threads, "kind": rt.kind,
                       "host_cpus": os.cpu_count(),
                       "sample": (f"{n_s} seeded random target rows of the same {n_targets}-row job in {st['wall']:.1f}s; "
                                  f"full job = T_fixed {t_fixed:.1f}s (call with 8 rows) + {n_targets} x {st['t_row'] * 1e3:.3f} ms/row "
                                  f"= {st['full_job_s']:.1f}s; scipy tocsr of URM.T ({t_tocsr:.1f}s) excluded")}
            except Exception as exc:  # the baseline must never take the GPU numbers down with it
                cpu = {"value": None, "unit": UNIT, "error": repr(exc)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: cosine item-item on BM25-normalised URM, k=100",
                       "n_users": n_users, "n_items": n_items, "density": density, "nnz": nnz, "k": K_NEIGHBOURS,
                       "target_rows_per_gpu": n_targets, "out_nnz_per_gpu": out_nnz,
                       "out_nnz_per_s": round(world * out_nnz / (ms_per_step / 1e3), 1),
                       "l2_policy": "inputs (1.6 GB CSR per operand) far larger than the 126 MB L2; no flush needed",
                       "parallelism": f"target rows sharded over {world} GPU(s), B replicated, no collective"},
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_users = max(64, int(CFG2["n_users"] * args.scale))
    n_items = max(64, int(CFG2["n_items"] * args.scale))
    density = CFG2["density"] if args.scale == 1.0 else min(0.5, CFG2["density"] / args.scale ** 0.5)
    urm = None
    try:  # same generator as the GPU arm when a GPU is there (bit-identical workload), numpy otherwise
        import torch
        if torch.cuda.is_available():
            import similaripy_b200 as sim
            from similaripy_b200 import _engine
            dev = torch.device("cuda", 0)
            indptr, indices, data = gen_urm_device(n_users, n_items, density, CFG2["seed"], dev)
            m = sim.DeviceMatrix(_engine.DeviceCSR(n_users, n_items, indptr, indices, data, sorted_rows=True), False)
            m = sim.bm25(m, inplace=True)
            s = m.stored
            urm = sp.csr_array((s.data.cpu().numpy(), s.indices.cpu().numpy(), s.indptr.cpu().numpy()),
                               shape=(n_users, n_items))
            del m, s, indptr, indices, data
            torch.cuda.empty_cache()
    except Exception as exc:
        log(f"[bench reference] GPU generator unavailable ({exc!r}); generating on the host")
    if urm is None:
        from oracle import oracle
        rng = np.random.default_rng(CFG2["seed"])
        cnt = rng.binomial(n_items, density, size=n_users)
        rows = np.repeat(np.arange(n_users, dtype=np.int64), cnt)
        key = np.unique(rows * n_items + rng.integers(0, n_items, size=rows.shape[0]))
        r = key // n_items
        indptr = np.zeros(n_users + 1, dtype=np.int64)
        np.cumsum(np.bincount(r, minlength=n_users), out=indptr[1:])
        urm = sp.csr_array(((1.0 - rng.random(key.shape[0], dtype=np.float32)), (key - r * n_items).astype(np.int32),
                            indptr.astype(np.int32)), shape=(n_users, n_items))
        urm = oracle.bm25(urm)
    a_csr = urm.T.tocsr()
    rt = ReferenceTimer(a_csr, n_items)
    t_fixed, t_row = rt.calibrate()
    total_steps = args.steps + args.warmup
    per_step = max(2.0, (args.ref_budget_s - total_steps * t_fixed) / max(total_steps, 1))
    n_s = rt.sample_size(per_step)
    log(f"[bench reference] kind={rt.kind} threads={rt.threads} T_fixed={t_fixed:.2f}s t_row~{t_row * 1e3:.3f}ms sample={n_s}")
    for _ in range(args.warmup):
        rt.step(n_s)
    stats = [rt.step(n_s) for _ in range(args.steps)]
    full = float(np.mean([s["full_job_s"] for s in stats]))
    wall = float(np.mean([s["wall"] for s in stats]))
    value = n_items / full
    sample = (f"each step = one reference call on {n_s} seeded random target rows of the {n_items}-row job "
              f"({wall:.1f}s); value extrapolates to the full job: T_fixed {rt.t_fixed:.1f}s + {n_items} rows x "
              f"{np.mean([s['t_row'] for s in stats]) * 1e3:.3f} ms/row = {full:.1f}s")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(wall * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: cosine item-item on BM25-normalised URM, k=100", "n_users": n_users,
                       "n_items": n_items, "density": density, "nnz": int(urm.nnz), "k": K_NEIGHBOURS},
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": rt.threads, "kind": rt.kind,
                             "host_cpus": os.cpu_count(), "sample": sample},
            "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (testing only; 1.0 = configs[1])")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work in the cpu_baseline sample")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="wall budget of the --impl reference run")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: timing rules ask for >= 3 warm-up steps")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
