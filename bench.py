#!/usr/bin/env python
"""bench.py -- similarity rows/sec of the sparse-KNN hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scale F]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: BASELINE.json configs[1] -- item-item ``cosine(URM.T, k=100)`` on a BM25-normalised synthetic URM of
1M users x 200k items, density 1e-3 (SURVEY.md 8d generator, seed 2; the same bits in both arms, see gen_urm_*).
One "step" = ONE complete similarity call over all 200k target rows: transposition / norm vectors / kernel tables
(reused from the DeviceMatrix handle after the first call), the fused expand + accumulate + similarity + top-k
kernel and the CSR assembly of the result.

  value  rows/s with the inputs already resident in HBM (DeviceMatrix in, DeviceMatrix out);
  e2e    rows/s through the public drop-in call ``similaripy_b200.cosine(scipy_matrix, ...)`` with HOST
         buffers (pinned): the H2D copy of the CSR and the D2H copy of the result are inside the timed region;
         `e2e.pageable` is the same call on ordinary (pageable) numpy buffers, `e2e.int64_indices` on a scipy
         matrix with 64-bit index arrays;
  roofline  algorithmic bytes of the hot kernel / its CUDA-event duration against MEASURED_PEAKS.json hbm_gbs;
  cpu_baseline  the compiled, unmodified reference (oracle/_ref, OpenMP, all host cores) on a bounded sample
         of the same target rows, extrapolated to the full job (see ``sample``);
  configs   BASELINE.json configs[2..4] (s_plus 500k x 500k, rp3beta 2M x 500k, dot_product 5M x 200k with
         filter_cols) at full operand size on a seeded sample of their target rows: rows/s, out-nnz/s, roofline;
  normalizers  l1 / l2 / tfidf / bm25 / bm25plus on the configs[1] URM: GB/s and fraction of the HBM peak.

N>1 is STRONG scaling of the same call (north_star): under ``sharded.shard_rows(gather=True)`` the 200k
target rows are cut into work-balanced contiguous ranges, B is replicated, every rank runs the kernel on its range
and writes into its slice of the gather buffer, one NCCL all-gather per slab array reassembles the full result on
every rank -- all inside the timed region.  value = 200k rows / max-over-ranks time.

``--impl reference`` times the reference's own CPU implementation (same metric / config) on ALL host cores; it
imports nothing of similaripy_b200.
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
# synthetic data: the SAME bits from numpy on the host (reference arm) and from torch on the device (GPU arm)
# ------------------------------------------------------------------------------------------------
# SURVEY 8d's generator -- per-row nnz ~ Binomial(n_cols, density), uniform distinct ascending columns, float32
# values -- with a counter-based hash (splitmix64) in place of a sequential RNG stream, so that the reference arm
# (numpy, no GPU, no import of the product) and the GPU arm (torch, scipy.sparse.random is far too slow at 2e8 nnz)
# build bit-identical matrices.  The per-row counts come from np.random.default_rng(seed).binomial on the host in
# both arms (n_rows draws).
_C0, _C1, _C2 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB


def _s64(c):  # 64-bit constant as a signed python int (torch int64)
    return c - (1 << 64) if c >= (1 << 63) else c


def _mix_np(x):
    x = x + np.uint64(_C0)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(_C1)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(_C2)
    return x ^ (x >> np.uint64(31))


def _mix_t(x):  # int64 tensor; logical right shifts emulated by masking the sign extension
    x = x + _s64(_C0)
    x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * _s64(_C1)
    x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * _s64(_C2)
    return x ^ ((x >> 31) & ((1 << 33) - 1))


def row_counts(n_rows, n_cols, density, seed):
    return np.random.default_rng(seed).binomial(n_cols, density, size=n_rows).astype(np.int64)


def gen_urm_host(n_rows, n_cols, density, seed):
    """scipy csr_array (int32 indices, float32 data in (0, 1]) -- numpy only."""
    cnt = row_counts(n_rows, n_cols, density, seed)
    rows = np.repeat(np.arange(n_rows, dtype=np.uint64), cnt)
    with np.errstate(over="ignore"):
        z = _mix_np(np.arange(rows.shape[0], dtype=np.uint64) + np.uint64(seed) * np.uint64(_C2))
        key = np.unique(rows * np.uint64(n_cols) + (z >> np.uint64(11)) % np.uint64(n_cols))
        del rows, z
        r = key // np.uint64(n_cols)
        indices = (key - r * np.uint64(n_cols)).astype(np.int32)
        z = _mix_np(key ^ (np.uint64(seed) * np.uint64(_C1)))
    data = (1.0 - ((z >> np.uint64(40)) & np.uint64(0xFFFFFF)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(np.bincount(r.astype(np.int64), minlength=n_rows), out=indptr[1:])
    m = sp.csr_array((data, indices, indptr.astype(np.int32)), shape=(n_rows, n_cols))
    m.has_sorted_indices = True
    return m


def gen_urm_device(n_rows, n_cols, density, seed, device):
    """The same matrix as gen_urm_host as device tensors (indptr i32, indices i32, data f32)."""
    import torch
    cnt = torch.from_numpy(row_counts(n_rows, n_cols, density, seed)).to(device)
    rows = torch.repeat_interleave(torch.arange(n_rows, device=device, dtype=torch.int64), cnt)
    seed_a = _s64((seed * _C2) & ((1 << 64) - 1))
    seed_b = _s64((seed * _C1) & ((1 << 64) - 1))
    z = _mix_t(torch.arange(rows.numel(), device=device, dtype=torch.int64) + seed_a)
    key = torch.unique(rows * n_cols + ((z >> 11) & ((1 << 53) - 1)) % n_cols, sorted=True)
    del rows, z
    r = torch.div(key, n_cols, rounding_mode="floor")
    indices = (key - r * n_cols).to(torch.int32)
    z = _mix_t(key ^ seed_b)
    del key
    data = 1.0 - ((z >> 40) & 0xFFFFFF).to(torch.float32) * (2.0 ** -24)
    del z
    indptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.bincount(r, minlength=n_rows), 0)
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
        self.first = 0

    def wait_ready(self, timeout_s=5.0):
        """Block until nvidia-smi has delivered its first sample (it takes a few hundred ms to come up)."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.01)

    def mark(self):
        """The timed region starts now: only samples from here on count (nvidia-smi needs a few hundred ms to come up,
        so it is started before the warm-up steps)."""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
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
        lines, note = self.lines[self.first:], None
        if not lines and self.lines:  # a timed region shorter than the sampling period: the last samples of the warm-up steps
            lines, note = self.lines[-3:], "timed region shorter than the sampling period: last samples of the warm-up steps (same load)"
        for ln in lines:
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
        out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
               "samples": len(sm), "power_w_max": float(max(power))}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------------------
# the reference on the host cores (oracle/_ref = the compiled, unmodified reference)
# ------------------------------------------------------------------------------------------------
def host_threads():
    """All the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the
    reference: its thread count is passed explicitly, `#pragma omp parallel num_threads(n)`, s_plus.h:313)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def reference_callable():
    """(fn, kind, threads): fn(matrix1, target_rows) -> csr result through the reference's own s_plus driver."""
    from oracle import ref_api
    threads = host_threads()
    if ref_api.available():
        from oracle import ref_api as impl
        kind = "reference"
    else:
        from oracle import oracle as impl
        kind = "port"

    def fn(matrix1, target_rows):
        return impl.similarity("cosine", matrix1, None, k=K_NEIGHBOURS, target_rows=target_rows,
                               format_output="csr", verbose=False, num_threads=threads)
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
    """scipy csr_array over existing host buffers without copying."""
    m = sp.csr_array((data, indices, indptr), shape=shape, copy=False)
    assert np.shares_memory(m.data, data) and np.shares_memory(m.indices, indices)
    m.has_sorted_indices = True
    return m


def load_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def device_matrix(n_rows, n_cols, density, seed, dev):
    import similaripy_b200 as sim
    from similaripy_b200 import _engine
    ip, ix, dv = gen_urm_device(n_rows, n_cols, density, seed, dev)
    return sim.DeviceMatrix(_engine.DeviceCSR(n_rows, n_cols, ip, ix, dv, sorted_rows=True), False)


def job_products(job):
    """(scalar products, stored entries of A) of the job's target rows -- SURVEY 8d's P and nnz_A."""
    import torch
    A, B = job.A, job.B
    b_len = (B.indptr[1:] - B.indptr[:-1]).to(torch.int64)
    cum = torch.zeros(A.nnz + 1, dtype=torch.int64, device=A.indptr.device)
    torch.cumsum(b_len[A.indices.long()], 0, out=cum[1:])
    t = job.targets.long()
    lo, hi = A.indptr[t].long(), A.indptr[t + 1].long()
    return int((cum[hi] - cum[lo]).sum().item()), int((hi - lo).sum().item())


def kernel_line(name, job, n_rows_total, peak, reps=3, world=1):
    """Hot-kernel numbers of one prepared job (tables built, operands resident): median of `reps` launches.  world > 1: the
    job was prepared under sharded.shard_rows(gather=True) -- this rank holds its share of the sampled target rows; the time
    is the maximum over the ranks, plus ONE NCCL all-gather of the output slab (timed once, after the launches)."""
    import torch
    products, nnz_a = job_products(job)
    n_t = job.n_targets
    ms = []
    for _ in range(reps + 1):
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(job.ctx.stream); job.run(); e1.record(job.ctx.stream); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    k_ms = float(np.median(ms[1:]))
    out_nnz = int(job.out_counts.sum().item())
    gather_ms = None
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        job.gather()
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - t0) * 1e3
        mx = torch.tensor([k_ms, gather_ms], device=job.out_cols.device, dtype=torch.float64)
        sm = torch.tensor([products, nnz_a, out_nnz, n_t], device=job.out_cols.device, dtype=torch.float64)
        torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(sm, op=torch.distributed.ReduceOp.SUM)
        k_ms, gather_ms = float(mx[0]), float(mx[1])
        products, nnz_a, out_nnz, n_t = (int(x) for x in sm.tolist())
    alg = 8 * n_t + 16 * nnz_a + 8 * products + 8 * out_nnz
    step_ms = k_ms + (gather_ms or 0.0)
    line = {"config": name, "target_rows_sampled": n_t, "target_rows_total": n_rows_total, "k": job.k,
            "rows_per_s": round(n_t / step_ms * 1e3, 1), "out_nnz_per_s": round(out_nnz / step_ms * 1e3, 1),
            "kernel_ms": round(k_ms, 3), "products": products, "gproducts_per_s": round(products / k_ms / 1e6, 1),
            "roofline": {"bound": "hbm", "achieved": round(alg / k_ms / 1e6, 1), "peak": peak * world, "unit": "GB/s",
                         "frac": round(alg / k_ms / 1e6 / (peak * world), 4), "algorithmic_bytes": alg},
            "full_job_estimate_s": round(n_rows_total / (n_t / step_ms * 1e3), 2),
            "plan": {"engine": int(job.args.engine), "n_panels": int(job.args.n_panels),
                     "panel_width": int(job.args.panel_width), "group": int(job.args.group)},
            "sample": f"hot kernel on {n_t} seeded random target rows of {n_rows_total}; operands at full size",
            "cpu_baseline": None}
    if world > 1:
        line["n_gpus"] = world
        line["all_gather_ms"] = round(gather_ms, 3)
        line["sample"] += (f"; ONE call's target rows cut by work over {world} GPUs, kernel time = max over the ranks, rows/s "
                           "includes one NCCL all-gather of the output slab")
    return line


def sample_rows(n, m, seed):
    return np.sort(np.random.default_rng(seed).choice(n, size=min(m, n), replace=False)).astype(np.int32)


def other_configs(local, peak, scale, world=1):
    """BASELINE.json configs[2..4]: full-size synthetic operands, a seeded sample of the target rows; on one GPU, or (world > 1)
    the sample cut by work over the ranks like any sharded call."""
    import torch
    import similaripy_b200 as sim
    from similaripy_b200 import _engine, sharded
    import contextlib
    dev = torch.device("cuda", local)
    shard = (lambda: sharded.shard_rows(gather=True)) if world > 1 else contextlib.nullcontext
    sc = lambda n: max(64, int(n * scale))
    out = []

    def guarded(name, fn):
        try:
            out.append(fn())
        except Exception as exc:  # one configuration must not take the line down
            out.append({"config": name, "error": repr(exc)[:300]})
        torch.cuda.empty_cache()

    def cfg2():  # s_plus(X, k=200, shrink=10), public defaults l1=l2=0.5, t1=t2=1, c1=c2=0.5 (similarity.py:509-515)
        x = device_matrix(sc(500_000), sc(500_000), 2e-3, 3, dev)
        with shard():
            job = _engine.prepare_job(x, None, k=200, target_rows=sample_rows(sc(500_000), 20_000, 3), verbose=False, device=local,
                                      l1=0.5, l2=0.5, t1=1.0, t2=1.0, c1=0.5, c2=0.5, stabilized_shrink=10.0)
        return kernel_line("configs[2]: s_plus k=200 shrink=10, 500k x 500k d=2e-3", job, sc(500_000), peak, world=world)

    def cfg3():  # rp3beta(URM.T, alpha=1, beta=0.6, k=100): item-item (benchmark.py:161), similarity.py:477-503
        urm = device_matrix(sc(2_000_000), sc(500_000), 5e-4, 4, dev)
        pop = _engine.axis_sum(urm, 0)
        with shard():
            job = _engine.prepare_job(sim.normalize(urm.T, norm="l1", axis=1), sim.normalize(urm, norm="l1", axis=1), k=100,
                                      target_rows=sample_rows(sc(500_000), 60_000, 4), verbose=False, device=local,
                                      weight_depop_matrix2=pop, p2=0.6, l3=1.0)
        return kernel_line("configs[3]: rp3beta beta=0.6 k=100 item-item, URM 2M x 500k d=5e-4", job, sc(500_000), peak, world=world)

    def cfg4():  # dot_product(URM, S.T, k=100, filter_cols=URM): S with ~100 neighbours per item
        urm = device_matrix(sc(5_000_000), sc(200_000), 1e-3, 5, dev)
        s_t = device_matrix(sc(200_000), sc(200_000), 5e-4, 55, dev)
        with shard():
            job = _engine.prepare_job(urm, s_t, k=100, target_rows=sample_rows(sc(5_000_000), 500_000, 5), filter_cols=urm,
                                      verbose=False, device=local)
        return kernel_line("configs[4]: dot_product URM x S.T filter_cols=URM k=100, URM 5M x 200k d=1e-3", job, sc(5_000_000), peak, world=world)

    guarded("configs[2]", cfg2)
    guarded("configs[3]", cfg3)
    guarded("configs[4]", cfg4)
    return out


def normalizer_lines(urm, peak):
    """In-place CSR normalizers on the configs[1] URM (2e8 nnz, f32 / i32): GB/s over the algorithmic bytes and the
    fraction of the HBM peak."""
    import torch
    import similaripy_b200 as sim
    nnz = urm.nnz
    work = urm.stored
    saved = work.data.clone()
    cases = [("l1", lambda m: sim.normalize(m, norm="l1", inplace=True)),
             ("l2", lambda m: sim.normalize(m, norm="l2", inplace=True)),
             ("tfidf", lambda m: sim.tfidf(m, inplace=True)),
             ("bm25", lambda m: sim.bm25(m, inplace=True)),
             ("bm25plus", lambda m: sim.bm25plus(m, inplace=True))]
    out = []
    for name, fn in cases:
        ms = []
        for _ in range(4):
            work.data.copy_(saved)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(urm); e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = float(np.median(ms[1:]))
        # algorithmic bytes: l1 / l2 read and rewrite the values (8 B per entry); tfidf / bm25 read values + column ids in the
        # statistics pass (8 B) and read both + rewrite the values in the apply pass (12 B): 20 B per entry
        bytes_ = nnz * 20 if name in ("tfidf", "bm25", "bm25plus") else nnz * 8
        gbs = bytes_ / t / 1e6
        out.append({"normalizer": name, "ms": round(t, 3), "gnnz_per_s": round(nnz / t / 1e6, 2), "gb_per_s": round(gbs, 1),
                    "frac": round(gbs / peak, 4), "bytes_model": f"{bytes_} B per call"})
    work.data.copy_(saved)
    work.invalidate()
    return out


def run_ours(args):
    import torch
    import similaripy_b200 as sim
    from similaripy_b200 import _engine, _lib, sharded

    world, rank, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    n_users = max(64, int(CFG2["n_users"] * args.scale))
    n_items = max(64, int(CFG2["n_items"] * args.scale))
    density = CFG2["density"] if args.scale == 1.0 else min(0.5, CFG2["density"] / args.scale ** 0.5)
    peak, peak_source = load_peak()
    t_setup = time.perf_counter()

    # ---- inputs: every rank holds the same URM (B is replicated) ----------------------------------
    urm_raw = device_matrix(n_users, n_items, density, CFG2["seed"], dev)
    normalizers = None
    if rank == 0 and world == 1 and not args.no_extras:
        normalizers = normalizer_lines(urm_raw, peak)
    urm = sim.bm25(urm_raw, inplace=True)  # BM25-normalised, as configs[1] says; outside the timed region
    m1 = urm.T  # exactly configs[1]: cosine(URM.T)
    torch.cuda.synchronize()
    nnz = urm.nnz
    n_targets = n_items
    log(f"[bench r{rank}] URM {n_users}x{n_items} nnz={nnz} generated+bm25 in {time.perf_counter() - t_setup:.1f}s")

    common = dict(k=K_NEIGHBOURS, verbose=False, format_output="csr", device=local)
    if os.environ.get("SPY_TUNING"):  # kernel experiments only, e.g. SPY_TUNING="engine=1,threads=512"
        common["tuning"] = {k: int(v) for k, v in (kv.split("=") for kv in os.environ["SPY_TUNING"].split(","))}

    def sharded_call(fn):
        if world > 1:  # ONE call, target rows cut by work over the ranks, full result gathered on every rank
            with sharded.shard_rows(gather=True):
                return fn()
        return fn()

    def step_device():
        return sharded_call(lambda: sim.cosine(m1, None, on_device=True, **common))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for _ in range(args.warmup):
        res = step_device()
    out_nnz = res.nnz
    del res
    barrier()
    _engine.KERNEL_TRACE = []
    _lib.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
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
    plan = {k: trace[0][k] for k in ("engine", "n_panels", "panel_width", "threads", "group")}
    rows_local = int(trace[0]["n_targets"])
    kern_ms_max = kern_ms
    if world > 1:
        t = torch.tensor([elapsed_ms, kern_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        elapsed_ms, kern_ms_max = float(t[0]), float(t[1])
    ms_per_step = elapsed_ms / args.steps
    value = n_targets / (ms_per_step / 1e3)

    # ---- algorithmic bytes of THIS rank's launch of the hot kernel ------------------------------
    probe = sharded_call(lambda: _engine.prepare_job(m1, None, **{k: v for k, v in common.items() if k != "format_output"}))
    products, nnz_a = job_products(probe)
    share = probe.n_targets / max(n_targets, 1)
    del probe
    alg_bytes = 8 * rows_local + 16 * nnz_a + 8 * products + 8 * int(out_nnz * share)
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_source,
                "kernel": "knn_stream_kernel" if plan["engine"] == 2 else "knn_flat_kernel", "kernel_ms": round(kern_ms, 3),
                "kernel_ms_max_over_ranks": round(kern_ms_max, 3), "algorithmic_bytes": alg_bytes, "target_rows_this_rank": rows_local,
                "products": products, "gproducts_per_s": round(products / (kern_ms / 1e3) / 1e9, 1),
                "kernel_share_of_step": round(kern_ms_max / ms_per_step, 4), "plan": plan}
    traffic_file = os.path.join(ROOT, "profiles", "knn_traffic.json")
    if os.path.exists(traffic_file) and world == 1:  # dram bytes per launch from the committed ncu --set full capture
        try:
            tj = json.load(open(traffic_file))
            # ("kernel": "<kernel name> <version tag of the capture>")
            if str(tj.get("kernel", "")).split(" ")[0] == roofline["kernel"] and abs(tj.get("workload_nnz", 0) - nnz) <= 0.001 * nnz:
                roofline["traffic"] = tj.get("dram_bytes_per_launch")
                roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- end to end: host buffers in, host result out ---------------------------------------------------------
    e2e = None
    cpu = None
    if not args.no_e2e:
        src = urm.stored  # CSR of URM; matrix1 = its transpose, as a scipy CSC view -- no host conversion
        h = [pinned_numpy(t) for t in (src.indptr, src.indices, src.data)]
        urm_host = host_csr_views(h[0], h[1], h[2], (n_users, n_items))
        h2d = sum(x.nbytes for x in h)

        def timed_host(matrix, steps):
            call = lambda: sharded_call(lambda: sim.cosine(matrix.T, None, **common))
            for _ in range(max(1, args.warmup)):  # (all W warm-up steps: the first host-buffer calls also pin staging memory and, under torchrun, open the NCCL channels)
                r = call()
            barrier()
            t0 = time.perf_counter()
            ev0.record()
            for _ in range(steps):
                r = call()
            ev1.record()
            barrier()
            wall = time.perf_counter() - t0
            ms = max(ev0.elapsed_time(ev1), wall * 1e3) / steps
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                ms = float(t[0])
            return ms, r

        e_steps = max(1, min(args.steps, 5))
        e_ms, r = timed_host(urm_host, e_steps)
        d2h = r.data.nbytes + r.indices.nbytes + r.indptr.nbytes
        e2e = {"value": round(n_targets / (e_ms / 1e3), 1), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": round(e_ms, 2), "steps": e_steps,
               "call": "similaripy_b200.cosine(scipy csc (pinned), k=100, format_output='csr') -> scipy csr"
                       + (" under sharded.shard_rows(gather=True) on every rank" if world > 1 else "")}
        if world == 1 and not args.no_extras:
            # the same call on what a user normally holds: pageable numpy buffers, and 64-bit index arrays
            pg = host_csr_views(h[0].copy(), h[1].copy(), h[2].copy(), (n_users, n_items))
            ms_pg, _ = timed_host(pg, max(1, min(e_steps, 2)))
            e2e["pageable"] = {"value": round(n_targets / (ms_pg / 1e3), 1), "ms_per_step": round(ms_pg, 2)}
            i64 = host_csr_views(h[0].astype(np.int64), h[1].astype(np.int64), pg.data, (n_users, n_items))
            ms_64, _ = timed_host(i64, max(1, min(e_steps, 2)))
            e2e["int64_indices"] = {"value": round(n_targets / (ms_64 / 1e3), 1), "ms_per_step": round(ms_64, 2),
                                    "h2d_bytes_per_step": int(sum(x.nbytes for x in (i64.indptr, i64.indices, i64.data)))}
            del pg, i64

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
        del urm_host, h

    configs = None
    if not args.no_extras:  # (every rank takes part when world > 1: the sampled target rows are cut over the ranks)
        del urm, urm_raw, m1
        torch.cuda.empty_cache()
        configs = other_configs(local, peak, args.scale, world)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: cosine item-item on BM25-normalised URM, k=100",
                       "n_users": n_users, "n_items": n_items, "density": density, "nnz": nnz, "k": K_NEIGHBOURS,
                       "target_rows": n_targets, "out_nnz": out_nnz,
                       "out_nnz_per_s": round(out_nnz / (ms_per_step / 1e3), 1),
                       "l2_policy": "inputs (1.6 GB CSR per operand) far larger than the 126 MB L2; no flush needed",
                       "parallelism": (f"ONE call: target rows cut by work over {world} GPU(s), B replicated, "
                                       + ("NCCL all-gather of the output slab inside the timed region" if world > 1 else "no collective"))},
            "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "gpu_launches": int(launches),
            "clocks": clocks, "configs": configs, "normalizers": normalizers,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_reference(args):
    """The reference's own CPU implementation on ALL host cores (rank 0 only); nothing of similaripy_b200 is imported."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_users = max(64, int(CFG2["n_users"] * args.scale))
    n_items = max(64, int(CFG2["n_items"] * args.scale))
    density = CFG2["density"] if args.scale == 1.0 else min(0.5, CFG2["density"] / args.scale ** 0.5)
    from oracle import ref_api
    t0 = time.perf_counter()
    urm = gen_urm_host(n_users, n_items, density, CFG2["seed"])  # bit-identical to the GPU arm's matrix
    if ref_api.available():
        urm = ref_api.bm25(urm)
    else:
        from oracle import oracle
        urm = oracle.bm25(urm)
    a_csr = urm.T.tocsr()
    log(f"[bench reference] URM {n_users}x{n_items} nnz={urm.nnz} generated + bm25 + tocsr on the host in {time.perf_counter() - t0:.1f}s")
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
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
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
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[2..4] / normalizers / pageable-e2e lines")
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
