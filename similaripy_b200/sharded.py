"""Target rows sharded over the GPUs of one box: one process per GPU under
``torch.distributed``; every rank holds B (and A) and computes a contiguous range of the target rows
balanced by WORK (scalar products, not row count); there is no data-path collective.  Only when the caller
wants the full matrix on every rank are the per-rank output slabs exchanged: ONE NCCL all-gather per slab
array over NVLink, in place -- the hot kernel writes its rows straight into the rank's slice of the gather
buffer, so nothing is copied before or after the collective.

The reference has a single parallel strategy, ``#pragma omp for schedule(dynamic)`` over the target rows
(similaripy/cython_code/s_plus.h:313-338, disjoint output ranges s_plus.h:443-450); this module is that
loop split across devices.

    import torch.distributed as dist, similaripy_b200 as sim
    dist.init_process_group("nccl")                       # torchrun, one rank per GPU
    with sim.sharded.shard_rows(gather=True):             # every similarity call inside is sharded
        S = sim.cosine(urm.T, k=100, format_output="csr")  # full matrix on every rank
    with sim.sharded.shard_rows(gather=False):
        S_local = sim.cosine(urm.T, k=100)                # this rank's rows only (others empty)

The planning / exchange helpers below are backend-agnostic (they are exercised with gloo on CPU tensors in
tests/test_sharded.py); the compute always runs through the CUDA library.
"""
from __future__ import annotations

import contextlib
import contextvars
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_ACTIVE: contextvars.ContextVar = contextvars.ContextVar("similaripy_b200_shard", default=None)


@dataclass
class ShardSpec:
    """How the current similarity calls are sharded."""
    gather: bool = True
    group: object = None          # torch.distributed process group (None: the default group)
    rank: Optional[int] = None    # override (tests); default: dist.get_rank(group)
    world: Optional[int] = None

    def resolve(self) -> Tuple[int, int]:
        if self.rank is not None and self.world is not None:
            return int(self.rank), int(self.world)
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("shard_rows needs an initialised torch.distributed process group "
                               "(launch one process per GPU with torchrun)")
        return dist.get_rank(self.group), dist.get_world_size(self.group)


@contextlib.contextmanager
def shard_rows(gather: bool = True, group=None, rank: Optional[int] = None, world: Optional[int] = None):
    """Context manager: similarity calls inside run on this rank's share of the target rows."""
    token = _ACTIVE.set(ShardSpec(gather=gather, group=group, rank=rank, world=world))
    try:
        yield
    finally:
        _ACTIVE.reset(token)


def active() -> Optional[ShardSpec]:
    return _ACTIVE.get()


def run(fn, *args, gather: bool = True, group=None, **kwargs):
    """Functional form: ``sharded.run(sim.cosine, urm.T, k=100, gather=True)``."""
    with shard_rows(gather=gather, group=group):
        return fn(*args, **kwargs)


# ------------------------------------------------------------------------------------------------
# planning: contiguous ranges of the target list with equal work
# ------------------------------------------------------------------------------------------------
def balanced_bounds(work: Sequence[int], n_parts: int) -> List[int]:
    """Cut positions ``b[0]=0 <= b[1] <= ... <= b[n_parts]=len(work)`` such that part p = [b[p], b[p+1]) and
    the parts' summed work is as even as a contiguous split allows: b[p] is the first position whose
    preceding cumulative work reaches p/n_parts of the total.  Rows with zero work still cost a slab row,
    so one unit is added to every row (this also spreads an all-empty target list evenly)."""
    w = np.asarray(work, dtype=np.int64) + 1
    n = int(w.shape[0])
    cum = np.concatenate(([0], np.cumsum(w)))
    total = int(cum[-1])
    bounds = [0]
    for p in range(1, n_parts):
        goal = (total * p + n_parts - 1) // n_parts
        b = int(np.searchsorted(cum, goal, side="left"))
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return bounds


def row_work_host(a_indptr, a_indices, b_indptr, targets) -> np.ndarray:
    """Host statement of spy_knn_row_work_dev (tests, planning from host CSR)."""
    b_len = np.diff(np.asarray(b_indptr, dtype=np.int64))
    per_entry = b_len[np.asarray(a_indices, dtype=np.int64)]
    cum = np.concatenate(([0], np.cumsum(per_entry)))
    a_indptr = np.asarray(a_indptr, dtype=np.int64)
    t = np.asarray(targets, dtype=np.int64)
    return cum[a_indptr[t + 1]] - cum[a_indptr[t]]


@dataclass
class ShardPlan:
    rank: int
    world: int
    bounds: List[int]   # positions in the target list

    @property
    def lo(self) -> int:
        return self.bounds[self.rank]

    @property
    def hi(self) -> int:
        return self.bounds[self.rank + 1]

    @property
    def n_local(self) -> int:
        return self.hi - self.lo

    @property
    def n_max(self) -> int:
        """Rows of the padded per-rank slab (all-gather needs equal contributions)."""
        return max(1, max(self.bounds[p + 1] - self.bounds[p] for p in range(self.world)))

    def padded_targets(self, targets: np.ndarray) -> np.ndarray:
        """Target row id of every row of the gathered slab [world * n_max]; -1 marks padding."""
        out = np.full(self.world * self.n_max, -1, dtype=np.int32)
        for p in range(self.world):
            lo, hi = self.bounds[p], self.bounds[p + 1]
            out[p * self.n_max: p * self.n_max + (hi - lo)] = targets[lo:hi]
        return out


# ------------------------------------------------------------------------------------------------
# exchange: in-place all-gather of the padded slabs
# ------------------------------------------------------------------------------------------------
class SlabExchange:
    """Gather buffers for (cols, values, counts) with this rank's slice exposed for the kernel to write."""

    def __init__(self, plan: ShardPlan, k: int, torch, device):
        self.plan, self.k, self.torch = plan, int(k), torch
        n = plan.world * plan.n_max
        self.cols = torch.empty(n * self.k, dtype=torch.int32, device=device)
        self.vals = torch.empty(n * self.k, dtype=torch.float32, device=device)
        self.counts = torch.zeros(n, dtype=torch.int32, device=device)  # padding rows: 0 entries

    def local(self):
        """(cols, vals, counts) views of this rank's slice: what the kernel's out_* pointers are set to."""
        p, m, k = self.plan.rank, self.plan.n_max, self.k
        return (self.cols[p * m * k: (p + 1) * m * k], self.vals[p * m * k: (p + 1) * m * k],
                self.counts[p * m: (p + 1) * m])

    def all_gather(self, group=None):
        """One collective per array; input == the rank's slice of the output (NCCL in-place all-gather)."""
        import torch.distributed as dist
        lc, lv, ln = self.local()
        for full, mine in ((self.cols, lc), (self.vals, lv), (self.counts, ln)):
            if dist.get_backend(group) == "gloo":  # gloo has no all_gather_into_tensor on every build: list form
                parts = list(full.chunk(self.plan.world))
                dist.all_gather(parts, mine.clone(), group=group)
            else:
                dist.all_gather_into_tensor(full, mine, group=group)
        return self.cols, self.vals, self.counts
