"""Multi-GPU path: work-balanced partition of the target rows, padded slab exchange and
assembly.  The N>1 plumbing is exercised here with world_size 2 over gloo on CPU tensors (each rank's slab
is produced by the oracle -- the checker -- because the CUDA kernel cannot run without a GPU); the GPU
test at the bottom runs the real sharded call on one device with an injected (rank, world)."""
import os
import socket
import sys

import numpy as np
import pytest
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from parity import assert_same_matrix, assert_topk_parity, random_csr  # noqa: E402
from similaripy_b200 import sharded  # noqa: E402


# ---- planning ------------------------------------------------------------------------------------
def test_balanced_bounds_covers_and_balances():
    rng = np.random.default_rng(0)
    work = rng.integers(0, 1000, size=5000)
    for parts in (1, 2, 3, 8):
        b = sharded.balanced_bounds(work, parts)
        assert b[0] == 0 and b[-1] == len(work) and len(b) == parts + 1
        assert all(b[i] <= b[i + 1] for i in range(parts))
        sums = [int(work[b[i]: b[i + 1]].sum()) for i in range(parts)]
        assert max(sums) - min(sums) <= 2 * 1000 + 2, sums  # within one row's work of even


def test_balanced_bounds_skewed_and_degenerate():
    work = np.array([10_000] + [1] * 99)
    b = sharded.balanced_bounds(work, 4)
    assert b[1] == 1  # the heavy row is a part of its own
    assert sharded.balanced_bounds([], 4) == [0, 0, 0, 0, 0]
    assert sharded.balanced_bounds([0, 0, 0, 0], 2) == [0, 2, 4]  # empty rows still spread evenly
    assert sharded.balanced_bounds([5], 3)[-1] == 1


def test_row_work_host_counts_products():
    a = random_csr(40, 30, 0.2, 1)
    b = random_csr(30, 50, 0.1, 2)
    targets = np.array([3, 0, 39, 7], dtype=np.int32)
    w = sharded.row_work_host(a.indptr, a.indices, b.indptr, targets)
    b_len = np.diff(b.indptr)
    expect = [int(b_len[a.indices[a.indptr[t]: a.indptr[t + 1]]].sum()) for t in targets]
    assert w.tolist() == expect


def test_padded_targets_layout():
    plan = sharded.ShardPlan(rank=1, world=3, bounds=[0, 2, 3, 7])
    t = np.arange(10, 17, dtype=np.int32)
    assert plan.n_max == 4 and plan.n_local == 1 and (plan.lo, plan.hi) == (2, 3)
    assert plan.padded_targets(t).tolist() == [10, 11, -1, -1, 12, -1, -1, -1, 13, 14, 15, 16]


def test_shard_context_requires_process_group():
    with sharded.shard_rows():
        with pytest.raises(RuntimeError):
            sharded.active().resolve()
    assert sharded.active() is None
    with sharded.shard_rows(rank=0, world=2):
        assert sharded.active().resolve() == (0, 2)


# ---- exchange over gloo, world_size 2 --------------------------------------------------------------
def _csr_rows_to_slab(res, targets, k):
    """Rows `targets` of a CSR result as a best-first slab (cols, vals, counts)."""
    n = len(targets)
    cols = np.zeros((n, k), dtype=np.int32)
    vals = np.zeros((n, k), dtype=np.float32)
    counts = np.zeros(n, dtype=np.int32)
    for i, t in enumerate(targets):
        s, e = res.indptr[t], res.indptr[t + 1]
        order = np.lexsort((res.indices[s:e], -res.data[s:e]))
        c = e - s
        cols[i, :c], vals[i, :c], counts[i] = res.indices[s:e][order], res.data[s:e][order], c
    return cols.ravel(), vals.ravel(), counts


def _assemble_padded(cols, vals, counts, padded_targets, k, shape):
    """Host statement of spy_slab_row_nnz_dev + scan + spy_slab_compact_dev on a padded slab."""
    rows_out, cols_out, vals_out = [], [], []
    for i, t in enumerate(padded_targets):
        if t < 0:
            continue
        c = counts[i]
        rows_out += [t] * c
        cols_out += cols[i * k: i * k + c].tolist()
        vals_out += vals[i * k: i * k + c].tolist()
    m = sp.csr_array((np.array(vals_out, dtype=np.float32), (np.array(rows_out), np.array(cols_out))), shape=shape)
    return m


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = 7
        urm = random_csr(150, 90, 0.08, 11)
        a = urm.T.tocsr()  # item-item: A = URM^T, B = URM
        targets = np.arange(a.shape[0], dtype=np.int32)[::-1].copy()  # any order; here descending
        work = sharded.row_work_host(a.indptr, a.indices, urm.indptr, targets)
        with sharded.shard_rows(gather=True):
            r, w = sharded.active().resolve()
        assert (r, w) == (rank, world)
        plan = sharded.ShardPlan(rank, world, sharded.balanced_bounds(work, world))
        local_targets = targets[plan.lo: plan.hi]
        local = oracle.similarity("cosine", a, k=k, target_rows=local_targets, format_output="csr")
        ex = sharded.SlabExchange(plan, k, torch, torch.device("cpu"))
        lc, lv, ln = ex.local()
        c, v, n = _csr_rows_to_slab(local, local_targets, k)
        lc[: c.shape[0]] = torch.from_numpy(c)
        lv[: v.shape[0]] = torch.from_numpy(v)
        ln[: n.shape[0]] = torch.from_numpy(n)
        cols, vals, counts = ex.all_gather()
        full = _assemble_padded(cols.numpy(), vals.numpy(), counts.numpy(), plan.padded_targets(targets), k,
                                (a.shape[0], urm.shape[1]))
        ref = oracle.similarity("cosine", a, k=k, format_output="csr")
        assert_same_matrix(ref, full, rtol=0, what=f"rank {rank}: gathered result")
        np.save(os.path.join(out_dir, f"bounds_{rank}.npy"), np.array(plan.bounds))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_matches_unsharded(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    b0, b1 = (np.load(tmp_path / f"bounds_{r}.npy") for r in range(world))
    assert b0.tolist() == b1.tolist()  # every rank computed the same cut, nothing was communicated for it
    assert 0 < b0[1] < b0[2]


# ---- the real sharded call on a GPU (one device plays each rank in turn) ---------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["csr", "coo"])
def test_gpu_sharded_local_parts_union_equals_full(fmt):
    import similaripy_b200 as sim
    urm = random_csr(600, 400, 0.05, 5)
    full = sim.cosine(urm.T, k=20, verbose=False, format_output="csr")
    world = 3
    acc = None
    seen = 0
    for rank in range(world):
        with sharded.shard_rows(gather=False, rank=rank, world=world):
            part = sim.cosine(urm.T, k=20, verbose=False, format_output=fmt).tocsr()
        part.eliminate_zeros()
        seen += int((np.diff(part.indptr) > 0).sum())
        acc = part if acc is None else acc + part
    assert_same_matrix(full, acc.tocsr(), rtol=1e-5, what="union of the ranks' rows")  # fp32 add order differs run to run
    assert seen == int((np.diff(full.indptr) > 0).sum())  # ranges are disjoint


@pytest.mark.gpu
def test_gpu_sharded_gather_single_process_group():
    """gather=True through a real (world_size 1) NCCL group: the in-place all-gather + padded assembly path."""
    import torch
    import torch.distributed as dist
    import similaripy_b200 as sim
    from oracle import oracle
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(_free_port()))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        urm = random_csr(500, 300, 0.05, 9)
        rows = np.arange(0, 300, 2, dtype=np.int32)
        for fmt in ("csr", "coo"):
            with sharded.shard_rows(gather=True):
                got = sim.rp3beta(urm.T, alpha=0.8, beta=0.5, k=15, target_rows=rows, verbose=False, format_output=fmt)
            ref = oracle.similarity("rp3beta", urm.T.tocsr(), alpha=0.8, beta=0.5, k=15, target_rows=rows,
                                    format_output="csr")
            assert_topk_parity(ref, got, k=15, rtol=1e-5, what=f"sharded gather {fmt}")
            if fmt == "coo":
                assert got.data.shape[0] == rows.shape[0] * 15  # the reference's COO keeps every slab entry
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_nccl_gather():
    """The real N>1 path: torchrun with one rank per GPU, NCCL all-gather over NVLink (needs >= 2 visible GPUs)."""
    import subprocess
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "sharded_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("SHARDED_OK") == world
