"""Host driver of the B200 similarity path: the Python mirror of the reference's Cython glue
``similaripy/cython_code/s_plus.pyx:95-433`` (validate -> CSR -> norm vectors -> selectors ->
kernel -> output matrix), with every O(nnz) step running as a CUDA kernel behind the C ABI
(``include/similaripy_b200.h``).  PyTorch is used only for device memory, streams and copies.

There is no CPU fallback: without the CUDA library or without a GPU the calls raise.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import scipy.sparse as sp

from . import _lib
from . import sharded as _sharded

# Reuse prepared operands across calls on the same DeviceMatrix (SIMILARIPY_B200_CACHE=0 switches it off)
CACHE_OPERANDS = os.environ.get("SIMILARIPY_B200_CACHE", "1") != "0"

# Pageable host arrays at least this large are uploaded through spy_h2d_staged (0 disables: torch's own staging)
STAGED_H2D_MIN_BYTES = int(os.environ.get("SIMILARIPY_B200_STAGED_H2D", str(8 << 20))) or (1 << 62)

# Defaults merged under every call's `tuning` (tests switch kernel generations with it): e.g. {"engine_prefer": "stream"}
DEFAULT_TUNING: dict = {}

# When set to a list, every hot-kernel launch appends {start, end (CUDA events), plan...} to it (bench.py).
KERNEL_TRACE = None

MODE_NONE, MODE_ARRAY, MODE_MATRIX = _lib.SEL_NONE, _lib.SEL_ARRAY, _lib.SEL_MATRIX
INT32_MAX = np.iinfo(np.int32).max
# A stored matrix with more entries than this keeps a 64-bit indptr (WideCSR) and is computed in int32-indexed blocks of
# at most this many entries (the reference narrows to int32, s_plus.pyx:241-244, and overflows).  Tests
# lower it to drive small matrices through the block path.
WIDE_NNZ_LIMIT = INT32_MAX


def _torch():
    import torch
    return torch


def resolve_device(device=None):
    """Device the call runs on: explicit argument, else $SIMILARIPY_B200_DEVICE, else the current CUDA device."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("similaripy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        env = os.environ.get("SIMILARIPY_B200_DEVICE")
        device = int(env) if env not in (None, "") else torch.cuda.current_device()
    if isinstance(device, int):
        return torch.device("cuda", device)
    return torch.device(device)


def preserve_device(fn):
    """Public entry points run on the device they were asked for and leave the caller's current CUDA device as it was
    (the C library launches on the current device, so a call on another device has to switch)."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        torch = _torch()
        prev = torch.cuda.current_device() if torch.cuda.is_available() else None
        try:
            return fn(*args, **kwargs)
        finally:
            if prev is not None and torch.cuda.current_device() != prev:
                torch.cuda.set_device(prev)
    return wrapper


def _ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Ctx:
    """Device + stream + the loaded C library for one call."""

    def __init__(self, device=None):
        torch = _torch()
        self.torch = torch
        self.lib = _lib.load()
        self.device = resolve_device(device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        torch.cuda.set_device(self.index)
        self.stream = torch.cuda.current_stream(self.device)
        self.sptr = C.c_void_p(self.stream.cuda_stream)

    # -- memory helpers ---------------------------------------------------------------------
    def empty(self, n, dtype):
        return self.torch.empty(int(n), dtype=dtype, device=self.device)

    def zeros(self, n, dtype):
        return self.torch.zeros(int(n), dtype=dtype, device=self.device)

    def h2d(self, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        if not arr.flags.writeable:
            arr = arr.copy()
        # pinned host memory (is_pinned) makes this a true async DMA; large pageable arrays -- what a scipy matrix normally
        # lives in -- go through the library's two pinned staging buffers filled by several host threads
        t = self.torch.from_numpy(arr)
        if arr.nbytes >= STAGED_H2D_MIN_BYTES and not t.is_pinned():
            out = self.torch.empty(t.shape, dtype=t.dtype, device=self.device)
            _lib.check(self.lib.spy_h2d_staged(out.data_ptr(), arr.ctypes.data, arr.nbytes, self.index, self.sptr))
            return out
        return t.to(self.device, non_blocking=True)

    def d2h(self, t) -> np.ndarray:
        host = self.torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        return host

    def sync(self):
        self.stream.synchronize()

    # -- handles shared between calls (and streams) ---------------------------------------------------
    def produced(self, csr):
        """`csr` was written by kernels queued on this call's stream: remember where."""
        ev = self.torch.cuda.Event()
        ev.record(self.stream)
        csr.ready = ev
        return csr

    def consume(self, csr):
        """This call is about to read `csr` on its own stream: order it behind the stream that produced it."""
        ev = getattr(csr, "ready", None)
        if ev is not None:
            self.stream.wait_event(ev)
        return csr

    # -- small kernels ----------------------------------------------------------------------
    def scan_i32(self, counts):
        n = counts.numel()
        out = self.empty(n + 1, self.torch.int32)
        tmp = self.empty(self.lib.spy_scan_tmp_bytes(n), self.torch.uint8)
        _lib.check(self.lib.spy_exclusive_scan_i32_dev(n, _ptr(counts), _ptr(out), _ptr(tmp), self.sptr))
        return out

    def scan_i64(self, counts):
        n = counts.numel()
        out = self.empty(n + 1, self.torch.int64)
        tmp = self.empty(self.lib.spy_scan_tmp_bytes(n), self.torch.uint8)
        _lib.check(self.lib.spy_exclusive_scan_i64_dev(n, _ptr(counts), _ptr(out), _ptr(tmp), self.sptr))
        return out


@dataclass
class DeviceCSR:
    """A CSR matrix resident in HBM: int32 indptr/indices, float32 data."""
    n_rows: int
    n_cols: int
    indptr: object
    indices: object
    data: object
    sorted_rows: bool = False
    scatter_order: bool = False  # rows in the arbitrary arrival order of an unsorted GPU transposition (not the caller's order)
    # Operands prepared once and reused by later calls on the same handle (s_plus.pyx:205-269 redoes all of this on
    # every call): transposes, the row-sorted copy, squared norms / sums, the kernel's stream layouts and split
    # tables.  Anything that overwrites `data` in place must call invalidate().
    cache: dict = field(default_factory=dict, repr=False, compare=False)
    # CUDA event recorded on the stream that produced (or last overwrote) the arrays; a call that consumes the handle on
    # another stream waits for it first (Ctx.consume).  None: produced by a call that synchronised before it returned.
    ready: object = field(default=None, repr=False, compare=False)

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())

    def invalidate(self) -> None:
        self.cache.clear()

    def cached(self, key, build, partner=None):
        """cache[key], built on first use.  `partner`: another object the value was derived from (e.g. the B
        operand of a table stored on A); the entry only counts while it is the very same object."""
        if not CACHE_OPERANDS:
            return build()
        hit = self.cache.get(key)
        if hit is not None and hit[0] is partner:
            return hit[1]
        value = build()
        self.cache[key] = (partner, value)
        return value


@dataclass
class WideCSR(DeviceCSR):
    """A CSR matrix with more than WIDE_NNZ_LIMIT stored entries: int64 indptr, int32 indices, float32 data.  The int32
    kernels never see it whole: `wide_blocks` cuts it into DeviceCSR blocks that keep its shape (see there)."""

    def host_indptr(self) -> np.ndarray:
        return self.cached("indptr_host", lambda: self.indptr.cpu().numpy())


def _greedy_cuts(prefix: np.ndarray, limit: int, what: str) -> list:
    """Cut positions 0 = c[0] < c[1] < ... = n such that prefix[c[i+1]] - prefix[c[i]] <= limit (prefix: int64 exclusive
    prefix sums with n + 1 entries)."""
    n = prefix.shape[0] - 1
    cuts = [0]
    while cuts[-1] < n:
        lo = cuts[-1]
        hi = int(np.searchsorted(prefix, prefix[lo] + limit, side="right")) - 1
        if hi <= lo:
            raise ValueError(f"one {what} holds more than {limit} stored entries: it cannot be indexed with int32")
        cuts.append(min(hi, n))
    return cuts


def wide_blocks(ctx: "Ctx", m: WideCSR, along_major: bool) -> list:
    """[(lo, hi, DeviceCSR)]: m cut into int32-indexed blocks of at most WIDE_NNZ_LIMIT entries along its rows
    (along_major) or its columns.  Every block has the SHAPE of m and holds the entries of rows (columns) [lo, hi); all
    other rows (columns) are empty in it, so row and column ids, norm vectors, selectors and target lists keep their
    meaning and the int32 kernels run on a block unchanged.  A row block is a zero-copy view: its indptr is
    clamp(indptr, first, last) - first, its indices / values are slices.  A column block filters every row block by the
    column range and stacks the pieces.  Explicit zeros are dropped like upload_stored does (s_plus.pyx:210-211)."""
    torch, lib, limit = ctx.torch, ctx.lib, int(WIDE_NNZ_LIMIT)

    def row_block(r0, r1, ip):
        first, last = int(ip[r0]), int(ip[r1])
        indptr = ctx.empty(m.n_rows + 1, torch.int32)
        _lib.check(lib.spy_csr_wide_block_indptr_dev(m.n_rows + 1, _ptr(m.indptr), first, last, _ptr(indptr), ctx.sptr))
        return DeviceCSR(m.n_rows, m.n_cols, indptr, m.indices[first:last], m.data[first:last], sorted_rows=m.sorted_rows)

    def build():
        ip = m.host_indptr()
        rows = _greedy_cuts(ip, limit, "row")
        if along_major:
            return [(r0, r1, filter_csr(ctx, row_block(r0, r1, ip), drop_zeros=True)) for r0, r1 in zip(rows[:-1], rows[1:])]
        counts = ctx.empty(max(m.n_cols, 1), torch.int32)[: m.n_cols]
        _lib.check(lib.spy_csr_col_count_dev(m.nnz, _ptr(m.indices), m.n_cols, _ptr(counts), ctx.sptr))
        prefix = np.zeros(m.n_cols + 1, dtype=np.int64)
        np.cumsum(counts.cpu().numpy(), out=prefix[1:])
        out = []
        cols = _greedy_cuts(prefix, limit, "column")
        for c0, c1 in zip(cols[:-1], cols[1:]):
            mask = ctx.zeros(max(m.n_cols, 1), torch.uint8)
            mask[c0:c1] = 1
            indptr = ctx.zeros(m.n_rows + 1, torch.int32)
            idx, val = [], []
            for r0, r1 in zip(rows[:-1], rows[1:]):
                piece = filter_csr(ctx, row_block(r0, r1, ip), col_mask=mask, drop_zeros=True)
                _lib.check(lib.spy_csr_indptr_add_dev(m.n_rows + 1, _ptr(piece.indptr), _ptr(indptr), ctx.sptr))
                idx.append(piece.indices)
                val.append(piece.data)
            out.append((c0, c1, DeviceCSR(m.n_rows, m.n_cols, indptr, torch.cat(idx), torch.cat(val), sorted_rows=m.sorted_rows)))
        return out
    return m.cached(("blocks", bool(along_major), limit), build)


def operand_blocks(ctx: "Ctx", stored: DeviceCSR, transposed: bool, axis: int) -> list:
    """[(lo, hi, DeviceMatrix)] covering logical axis `axis` (0: rows, 1: columns) of the operand (stored, transposed); a
    matrix that fits int32 indexing is its own single block."""
    n = (stored.n_cols, stored.n_rows)[axis] if transposed else (stored.n_rows, stored.n_cols)[axis]
    if not isinstance(stored, WideCSR):
        return [(0, n, DeviceMatrix(stored, transposed))]
    return [(lo, hi, DeviceMatrix(b, transposed)) for lo, hi, b in wide_blocks(ctx, stored, along_major=(axis == 0) != transposed)]


class DeviceMatrix:
    """A sparse matrix resident in HBM, usable wherever the similarity functions take ``matrix1`` /
    ``matrix2`` (chain calls without host round-trips).  ``stored`` is the CSR of the
    matrix when ``transposed`` is False and the CSR of its transpose (== its CSC) when True, so ``.T``
    is free, exactly like scipy's."""

    def __init__(self, stored: DeviceCSR, transposed: bool):
        self.stored = stored
        self.transposed = bool(transposed)

    @property
    def shape(self):
        s = self.stored
        return (s.n_cols, s.n_rows) if self.transposed else (s.n_rows, s.n_cols)

    @property
    def nnz(self) -> int:
        return self.stored.nnz

    @property
    def T(self) -> "DeviceMatrix":
        return DeviceMatrix(self.stored, not self.transposed)

    @property
    def device(self):
        return self.stored.indptr.device


def _is_matrix(m) -> bool:
    return sp.issparse(m) or isinstance(m, DeviceMatrix)


_VAL_CODES = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64,
              np.dtype(np.int32): _lib.VAL_I32, np.dtype(np.int64): _lib.VAL_I64}


def _upload_values(ctx: Ctx, data: np.ndarray, binary: bool):
    """matrix.data.astype(float32) / ones for binary (s_plus_utils.pyx:281-308) on the device."""
    torch = ctx.torch
    n = data.shape[0]
    if binary:
        out = ctx.empty(n, torch.float32)
        _lib.check(ctx.lib.spy_cast_values_dev(n, _ptr(out), _lib.F32, 1, _ptr(out), ctx.sptr))
        return out
    dt = np.dtype(data.dtype)
    if dt == np.float32:
        return ctx.h2d(data)
    if dt not in _VAL_CODES:
        return ctx.h2d(data.astype(np.float32))
    raw = ctx.h2d(data)
    out = ctx.empty(n, torch.float32)
    _lib.check(ctx.lib.spy_cast_values_dev(n, _ptr(raw), _VAL_CODES[dt], 0, _ptr(out), ctx.sptr))
    return out


def _upload_index(ctx: Ctx, arr: np.ndarray):
    """An index array as int32 on the device (s_plus.pyx:241-244).  A large int64 array is uploaded as it is and narrowed
    by a kernel: the host-side astype(int32) of 2e8 entries costs several times the extra PCIe traffic."""
    if arr.dtype == np.int64 and arr.nbytes >= STAGED_H2D_MIN_BYTES:
        wide = ctx.h2d(arr)
        out = ctx.empty(arr.shape[0], ctx.torch.int32)
        _lib.check(ctx.lib.spy_narrow_index_dev(arr.shape[0], _ptr(wide), _ptr(out), ctx.sptr))
        return out
    if arr.dtype != np.int32:
        arr = arr.astype(np.int32)
    return ctx.h2d(arr)


def transpose_csr(ctx: Ctx, m: DeviceCSR, sort: bool = True) -> DeviceCSR:
    """CSR -> CSR of the transpose (scipy csr_tocsc behind s_plus.pyx:205-206); rows sorted like scipy's unless
    sort=False (enough for the left operand of the similarity kernel, and a third of the transposition's time)."""
    torch = ctx.torch
    counts = ctx.empty(max(m.n_cols, 1), torch.int32)[: m.n_cols]
    _lib.check(ctx.lib.spy_csr_col_count_dev(m.nnz, _ptr(m.indices), m.n_cols, _ptr(counts), ctx.sptr))
    t_indptr = ctx.scan_i32(counts)
    t_indices = ctx.empty(m.nnz, torch.int32)
    t_data = ctx.empty(m.nnz, torch.float32)
    cursor = ctx.empty(max(m.n_cols, 1), torch.int32)
    _lib.check(ctx.lib.spy_csr_transpose_dev(m.n_rows, m.n_cols, _ptr(m.indptr), _ptr(m.indices), _ptr(m.data),
                                             _ptr(t_indptr), _ptr(t_indices), _ptr(t_data), _ptr(cursor), 1 if sort else 0,
                                             ctx.sptr))
    return DeviceCSR(m.n_cols, m.n_rows, t_indptr, t_indices, t_data, sorted_rows=bool(sort), scatter_order=not sort)


def cached_transpose(ctx: Ctx, m: DeviceCSR, sort: bool = True) -> DeviceCSR:
    """transpose_csr through the handle's cache (a row-sorted transpose also serves callers that do not need the order)."""
    if not CACHE_OPERANDS:
        return transpose_csr(ctx, m, sort=sort)
    hit = m.cache.get("T_sorted")
    if hit is None and not sort:
        hit = m.cache.get("T")
    if hit is not None:
        return hit[1]
    return m.cached("T_sorted" if sort else "T", lambda: transpose_csr(ctx, m, sort=sort))


def sorted_rows(ctx: Ctx, m: DeviceCSR) -> DeviceCSR:
    """m with ascending column ids inside every row (sort_indices, s_plus_utils.pyx:344,573).  The caller's buffers are
    never permuted: an unsorted matrix is cloned, the clone sorted, and the result kept on the handle."""
    if m.sorted_rows:
        return m

    def build():
        c = DeviceCSR(m.n_rows, m.n_cols, m.indptr, m.indices.clone(), m.data.clone(), sorted_rows=True)
        _lib.check(ctx.lib.spy_csr_sort_rows_dev(c.n_rows, _ptr(c.indptr), _ptr(c.indices), _ptr(c.data), ctx.sptr))
        return c
    return m.cached("sorted", build) if CACHE_OPERANDS else build()


def filter_csr(ctx: Ctx, m: DeviceCSR, col_mask=None, drop_zeros=False, values=None) -> DeviceCSR:
    """Keep entries with mask[col] != 0 and/or value != 0; returns m itself when nothing is dropped."""
    torch = ctx.torch
    data = m.data if values is None else values
    counts = ctx.empty(max(m.n_rows, 1), torch.int32)[: m.n_rows]
    _lib.check(ctx.lib.spy_csr_filter_count_dev(m.n_rows, _ptr(m.indptr), _ptr(m.indices), _ptr(data),
                                                _ptr(col_mask), 1 if drop_zeros else 0, _ptr(counts), ctx.sptr))
    new_indptr = ctx.scan_i32(counts)
    kept = int(new_indptr[-1].item()) if m.n_rows > 0 else 0
    if kept == m.nnz and values is None:
        return m
    new_indices = ctx.empty(kept, torch.int32)
    new_data = ctx.empty(kept, torch.float32)
    _lib.check(ctx.lib.spy_csr_filter_compact_dev(m.n_rows, _ptr(m.indptr), _ptr(m.indices), _ptr(data),
                                                  _ptr(col_mask), 1 if drop_zeros else 0, _ptr(new_indptr),
                                                  _ptr(new_indices), _ptr(new_data), ctx.sptr))
    return DeviceCSR(m.n_rows, m.n_cols, new_indptr, new_indices, new_data, sorted_rows=m.sorted_rows, scatter_order=m.scatter_order)


def upload_stored(ctx: Ctx, matrix):
    """Upload a scipy sparse matrix in its STORED orientation, zero-free, values as float32.

    Returns (DeviceCSR over the stored major axis, transposed_flag): a CSR matrix is uploaded as is
    (flag False); a CSC matrix -- what ``X.T`` of a CSR is -- is uploaded as the CSR of its transpose
    (flag True) without any host-side conversion; other formats go through scipy's ``tocsr`` first.
    Covers s_plus.pyx:205-211 (tocsr + eliminate_zeros) and :237-244 (float32 / int32 views).
    """
    if isinstance(matrix, DeviceMatrix):
        if matrix.device != ctx.device:
            raise ValueError(f"DeviceMatrix lives on {matrix.device}, the call runs on {ctx.device}")
        return ctx.consume(matrix.stored), matrix.transposed
    fmt = getattr(matrix, "format", None)
    if fmt not in ("csr", "csc"):
        matrix = matrix.tocsr()
        fmt = "csr"
    major, minor = (matrix.shape if fmt == "csr" else matrix.shape[::-1])
    if max(matrix.shape) > INT32_MAX:
        raise ValueError("matrix dimensions exceed int32, which the similarity kernel (like the reference) uses")
    if matrix.nnz > WIDE_NNZ_LIMIT:  # 64-bit indptr; computed in int32-indexed blocks (wide_blocks)
        return WideCSR(major, minor, ctx.h2d(np.asarray(matrix.indptr, dtype=np.int64)), _upload_index(ctx, matrix.indices),
                       _upload_values(ctx, matrix.data, binary=False),
                       sorted_rows=bool(getattr(matrix, "_has_sorted_indices", False))), fmt == "csc"
    m = DeviceCSR(major, minor, _upload_index(ctx, matrix.indptr), _upload_index(ctx, matrix.indices),
                  _upload_values(ctx, matrix.data, binary=False),
                  sorted_rows=bool(getattr(matrix, "_has_sorted_indices", False)))
    before = m.nnz
    m = filter_csr(ctx, m, drop_zeros=True)  # eliminate_zeros (s_plus.pyx:210-211)
    if m.nnz != before and fmt == "csr":
        matrix.eliminate_zeros()  # the reference mutates a CSR argument in place; keep that side effect
    return m, fmt == "csc"


def binarize(ctx: Ctx, m: DeviceCSR) -> DeviceCSR:
    """binary=True: every stored value becomes 1.0 (s_plus_utils.pyx:301-304)."""
    ones = ctx.empty(m.nnz, ctx.torch.float32)
    _lib.check(ctx.lib.spy_cast_values_dev(m.nnz, _ptr(ones), _lib.F32, 1, _ptr(ones), ctx.sptr))
    return DeviceCSR(m.n_rows, m.n_cols, m.indptr, m.indices, ones, sorted_rows=m.sorted_rows, scatter_order=m.scatter_order)


def upload_pair(ctx: Ctx, matrix1, matrix2):
    """A = matrix1 and B = matrix2 as device CSR.  With matrix2=None (B = matrix1.T, s_plus.pyx:169-170)
    the data crosses PCIe once and is transposed once on the GPU, whichever of CSR / CSC matrix1 is."""
    s1, t1 = upload_stored(ctx, matrix1)
    if isinstance(s1, WideCSR):
        return s1, s1  # (refused by the caller: s_plus cuts such operands into blocks first)
    if matrix2 is None:  # A's row order is free; B is sorted later only if the plan has several panels
        other = cached_transpose(ctx, s1, sort=False)
        return (other, s1) if t1 else (s1, other)
    A = cached_transpose(ctx, s1, sort=False) if t1 else s1
    s2, t2 = upload_stored(ctx, matrix2)
    if isinstance(s2, WideCSR):
        return A, s2
    B = cached_transpose(ctx, s2) if t2 else s2
    return A, B


# --------------------------------------------------------------------------------------------
# validation (s_plus_utils.pyx:19-125): same checks, same exception types, same order
# --------------------------------------------------------------------------------------------
def validate_inputs(matrix1, matrix2, weight_depop_matrix1, weight_depop_matrix2, k, target_rows,
                    filter_cols, target_cols, verbose, format_output) -> None:
    if not _is_matrix(matrix1):
        raise TypeError("matrix1 must be a sparse matrix")
    if not _is_matrix(matrix2):
        raise TypeError("matrix2 must be a sparse matrix")
    if matrix1.shape[1] != matrix2.shape[0]:
        raise ValueError(f"Incompatible matrix shapes: matrix1.shape[1]={matrix1.shape[1]} "
                         f"must equal matrix2.shape[0]={matrix2.shape[0]}")
    if k < 1:
        raise ValueError(f"k must be >= 1, got {k}")
    for name, w, n in (("weight_depop_matrix1", weight_depop_matrix1, matrix1.shape[0]),
                       ("weight_depop_matrix2", weight_depop_matrix2, matrix2.shape[1])):
        if isinstance(w, str):
            ok = w in ("none", "sum") or len(w) == n
        else:
            ok = len(w) == n
        if not ok:
            raise ValueError(f'{name} must be array of length {n} or one of ("none", "sum"), got length {len(w)}')
    if target_rows is not None and len(target_rows) > matrix1.shape[0]:
        raise ValueError(f"target_rows length ({len(target_rows)}) cannot exceed matrix1.shape[0] ({matrix1.shape[0]})")
    for name, cols in (("filter_cols", filter_cols), ("target_cols", target_cols)):
        if cols is None:
            continue
        if not (_is_matrix(cols) or isinstance(cols, (list, np.ndarray))):
            raise TypeError(f"{name} must be a sparse matrix, list, numpy array, or None")
        if isinstance(cols, DeviceMatrix):
            if cols.nnz != 0 and cols.shape != (matrix1.shape[0], matrix2.shape[1]):
                raise ValueError(f"{name} shape {cols.shape} does not match expected shape "
                                 f"{(matrix1.shape[0], matrix2.shape[1])}")
        elif sp.issparse(cols) and cols.data.shape[0] != 0:
            expected = (matrix1.shape[0], matrix2.shape[1])
            if cols.shape != expected:
                raise ValueError(f"{name} shape {cols.shape} does not match expected shape {expected}")
    if not isinstance(verbose, bool):
        raise TypeError(f"verbose must be boolean, got {type(verbose).__name__}")
    if format_output not in ("coo", "csr"):
        raise ValueError(f"format_output must be 'coo' or 'csr', got '{format_output}'")


def selector_mode(cols) -> int:
    """s_plus_utils.pyx:311-361 (a DeviceMatrix is a sparse matrix that already lives in HBM)."""
    if isinstance(cols, DeviceMatrix):
        return MODE_MATRIX if cols.nnz != 0 else MODE_NONE
    if sp.issparse(cols) and cols.data.shape[0] != 0:
        return MODE_MATRIX
    if isinstance(cols, (list, np.ndarray)) and len(cols) != 0:
        return MODE_ARRAY
    return MODE_NONE


def keep_mask(filter_cols, target_cols, n_cols: int) -> np.ndarray:
    """List-mode column set target \\ filter, out-of-range ids dropped (s_plus_utils.pyx:364-421)."""
    if selector_mode(target_cols) == MODE_ARRAY:
        mask = np.zeros(n_cols, dtype=np.uint8)
        t = np.asarray(target_cols, dtype=np.int32)
        mask[t[(t >= 0) & (t < n_cols)]] = 1
    else:
        mask = np.ones(n_cols, dtype=np.uint8)
    if selector_mode(filter_cols) == MODE_ARRAY:
        f = np.asarray(filter_cols, dtype=np.int32)
        mask[f[(f >= 0) & (f < n_cols)]] = 0
    return mask


@dataclass
class KnnJob:
    """Everything resident on the device for one similarity call."""
    ctx: Ctx
    A: DeviceCSR
    B: DeviceCSR
    targets: object
    n_targets: int
    k: int
    n_rows: int
    n_cols: int
    params: dict
    vectors: dict = field(default_factory=dict)
    filter_mode: int = MODE_NONE
    filter_m: tuple = (None, None)
    target_mode: int = MODE_NONE
    target_m: tuple = (None, None)
    args: Optional[_lib.KnnArgs] = None
    keep: list = field(default_factory=list)
    out_cols: object = None
    out_vals: object = None
    out_counts: object = None
    unique_targets: bool = True
    tuning: dict = field(default_factory=dict)
    shard: object = None      # sharded.ShardPlan when the target rows are split over ranks
    exchange: object = None   # sharded.SlabExchange when the full result is gathered on every rank
    targets_np: object = None  # the FULL target list on the host (sharded runs)
    targets_key: object = None  # ("all", n) / ("range", lo, hi, n) when the target list is (a range of) all rows, else None
    unit_values: bool = False   # binary=True: every stored value of A and B is 1.0 (the stream engine counts with integer adds)

    # ---- norm vectors (s_plus.pyx:259-269) ------------------------------------------------
    def build_vectors(self, weight_depop_matrix1, weight_depop_matrix2, p1, p2, c1, c2, additive_shrink):
        ctx, lib, torch = self.ctx, self.ctx.lib, self.ctx.torch
        A, B, P = self.A, self.B, self.params
        v = {}
        if P["l1"] != 0 or P["l2"] != 0:  # _build_squared_norms, s_plus_utils.pyx:169-201
            sq1 = A.cached("row_sq", lambda: self._row_sum(A, 1))
            sq2 = B.cached("col_sq", lambda: self._col_sum(B, 1))
        if P["l1"] != 0:
            v["Xt"], v["Yt"] = sq1, sq2
        if P["l2"] != 0:  # _build_cosine_normalization, s_plus_utils.pyx:204-228
            v["Xc"] = ctx.empty(A.n_rows, torch.float32)
            v["Yc"] = ctx.empty(B.n_cols, torch.float32)
            _lib.check(lib.spy_pow_shift_dev(A.n_rows, _ptr(sq1), _lib.F32, additive_shrink, c1, _ptr(v["Xc"]), ctx.sptr))
            _lib.check(lib.spy_pow_shift_dev(B.n_cols, _ptr(sq2), _lib.F32, additive_shrink, c2, _ptr(v["Yc"]), ctx.sptr))
        if P["l3"] != 0:  # _build_depop_normalization, s_plus_utils.pyx:231-278
            v["Xd"] = self._depop(weight_depop_matrix1, p1, axis=1)
            v["Yd"] = self._depop(weight_depop_matrix2, p2, axis=0)
        self.vectors = v

    def _row_sum(self, m, square):
        ctx = self.ctx
        out = ctx.empty(m.n_rows, ctx.torch.float32)
        _lib.check(ctx.lib.spy_csr_row_sum_dev(m.n_rows, _ptr(m.indptr), _ptr(m.data), square, _ptr(out), ctx.sptr))
        return out

    def _col_sum(self, m, square):
        ctx = self.ctx
        out = ctx.empty(m.n_cols, ctx.torch.float32)
        acc = ctx.empty(max(m.n_cols, 1), ctx.torch.float64)
        _lib.check(ctx.lib.spy_csr_col_sum_dev(m.nnz, _ptr(m.indices), _ptr(m.data), square, m.n_cols, _ptr(acc), _ptr(out), ctx.sptr))
        return out

    def _depop(self, spec, p, axis):
        ctx, lib, torch = self.ctx, self.ctx.lib, self.ctx.torch
        m = self.A if axis == 1 else self.B
        n = m.n_rows if axis == 1 else m.n_cols
        out = ctx.empty(n, torch.float32)
        if ctx.torch.is_tensor(spec):  # already on the device (DeviceMatrix pipelines)
            wd = spec.to(ctx.device)
            if wd.dtype not in (torch.float32, torch.float64):
                wd = wd.to(torch.float32)
            wd = wd.contiguous().reshape(-1)
            code = _lib.F32 if wd.dtype == torch.float32 else _lib.F64
            _lib.check(lib.spy_pow_shift_dev(n, _ptr(wd), code, 0.0, p, _ptr(out), ctx.sptr))
            self.keep.append(wd)
        elif isinstance(spec, (list, np.ndarray)):
            w = np.asarray(spec)
            if w.dtype not in (np.float32, np.float64):
                w = w.astype(np.float32)
            wd = ctx.h2d(w.ravel())
            code = _lib.F32 if w.dtype == np.float32 else _lib.F64
            _lib.check(lib.spy_pow_shift_dev(n, _ptr(wd), code, 0.0, p, _ptr(out), ctx.sptr))
            self.keep.append(wd)
        elif spec == "none":
            out.fill_(1.0)
        elif spec == "sum":
            s = m.cached("row_sum", lambda: self._row_sum(m, 0)) if axis == 1 else m.cached("col_sum", lambda: self._col_sum(m, 0))
            _lib.check(lib.spy_pow_shift_dev(n, _ptr(s), _lib.F32, 0.0, p, _ptr(out), ctx.sptr))
        else:
            raise ValueError(f"Invalid depopularization weights: {spec}")
        return out

    # ---- selectors (s_plus.pyx:284-295) -----------------------------------------------------
    def build_selectors(self, filter_cols, target_cols, raw_b_values):
        ctx = self.ctx
        self.filter_mode = selector_mode(filter_cols)
        self.target_mode = selector_mode(target_cols)
        for which, cols, mode in (("filter_m", filter_cols, self.filter_mode), ("target_m", target_cols, self.target_mode)):
            if mode == MODE_MATRIX and isinstance(cols, DeviceMatrix):  # already in HBM: rows sorted on the device
                m = transpose_csr(ctx, cols.stored) if cols.transposed else cols.stored
                m = filter_csr(ctx, m, drop_zeros=True)
                if not m.sorted_rows:
                    m = DeviceCSR(m.n_rows, m.n_cols, m.indptr, m.indices.clone(), m.data.clone(), sorted_rows=False)
                    _lib.check(ctx.lib.spy_csr_sort_rows_dev(m.n_rows, _ptr(m.indptr), _ptr(m.indices), _ptr(m.data), ctx.sptr))
                setattr(self, which, (m.indptr, m.indices))
            elif mode == MODE_MATRIX:  # per-row lists, sorted for the in-kernel range search
                c = cols.tocsr()
                c.eliminate_zeros()
                c.sort_indices()
                setattr(self, which, (_upload_index(ctx, c.indptr), _upload_index(ctx, c.indices)))
        if self.filter_mode == MODE_ARRAY or self.target_mode == MODE_ARRAY:
            mask = ctx.h2d(keep_mask(filter_cols, target_cols, self.n_cols))
            # the reference re-reads the restored (non-binary) values here: SURVEY 8a / a11
            self.B = filter_csr(ctx, self.B, col_mask=mask, values=raw_b_values)
            self.keep.append(mask)

    # ---- multi-GPU: keep this rank's share of the target rows -----------------------
    def shard_targets(self, spec, targets_np):
        """Cut the target list into world contiguous ranges of equal work (scalar products) and keep range
        `rank`.  Every rank computes the same cut from the same inputs; nothing is communicated."""
        ctx, torch = self.ctx, self.ctx.torch
        rank, world = spec.resolve()
        def cut():
            if self.n_targets > 0:
                work = ctx.empty(self.n_targets, torch.int64)
                _lib.check(ctx.lib.spy_knn_row_work_dev(self.n_targets, _ptr(self.targets), _ptr(self.A.indptr),
                                                        _ptr(self.A.indices), _ptr(self.B.indptr), _ptr(work), ctx.sptr))
                work_np = work.cpu().numpy()
            else:
                work_np = np.zeros(0, dtype=np.int64)
            return _sharded.balanced_bounds(work_np, world)
        bounds = self.A.cached(("cut", self.targets_key, world), cut, partner=self.B) if self.targets_key is not None else cut()
        self.shard = _sharded.ShardPlan(rank, world, bounds)
        self.targets_np = targets_np
        self.targets = self.targets[self.shard.lo: self.shard.hi]
        self.n_targets = self.shard.n_local
        if self.targets_key is not None:
            self.targets_key = ("range", self.shard.lo, self.shard.hi, self.targets_key[1])
        if spec.gather:
            self.exchange = _sharded.SlabExchange(self.shard, self.k, torch, ctx.device)
            self.gather_group = spec.group

    def gather(self):
        """All-gather the per-rank slabs in place; afterwards the job describes the FULL result."""
        ctx, torch = self.ctx, self.ctx.torch
        plan = self.shard
        ctx.sync()  # the collective runs on NCCL's stream: the kernel's writes must have landed
        cols, vals, counts = self.exchange.all_gather(self.gather_group)
        padded = plan.padded_targets(self.targets_np)
        if self.unique_targets:  # assemble straight from the padded slab (targets < 0 are skipped)
            self.out_cols, self.out_vals, self.out_counts = cols, vals, counts
            self.targets = ctx.h2d(padded)
            self.n_targets = plan.world * plan.n_max
            self.padded = True
        else:  # duplicates take the host assembly path: give it a contiguous slab
            k, m = self.k, plan.n_max
            n = [plan.bounds[p + 1] - plan.bounds[p] for p in range(plan.world)]
            self.out_cols = torch.cat([cols[p * m * k: (p * m + n[p]) * k] for p in range(plan.world)])
            self.out_vals = torch.cat([vals[p * m * k: (p * m + n[p]) * k] for p in range(plan.world)])
            self.out_counts = torch.cat([counts[p * m: p * m + n[p]] for p in range(plan.world)])
            self.targets = ctx.h2d(self.targets_np)
            self.n_targets = int(self.targets_np.shape[0])

    def compact_padded(self):
        """Drop the padding rows of a gathered slab (COO keeps every slab entry, s_plus.pyx:351-353)."""
        if not getattr(self, "padded", False):
            return
        plan, k, m, torch = self.shard, self.k, self.shard.n_max, self.ctx.torch
        n = [plan.bounds[p + 1] - plan.bounds[p] for p in range(plan.world)]
        self.out_cols = torch.cat([self.out_cols[p * m * k: (p * m + n[p]) * k] for p in range(plan.world)])
        self.out_vals = torch.cat([self.out_vals[p * m * k: (p * m + n[p]) * k] for p in range(plan.world)])
        self.out_counts = torch.cat([self.out_counts[p * m: p * m + n[p]] for p in range(plan.world)])
        self.targets = self.ctx.h2d(self.targets_np)
        self.n_targets = int(self.targets_np.shape[0])
        self.padded = False

    # ---- plan + split points ------------------------------------------------------------------
    def plan(self):
        ctx, lib, torch = self.ctx, self.ctx.lib, self.ctx.torch
        A, B, P, v = self.A, self.B, self.params, self.vectors
        a = _lib.KnnArgs()
        a.n_targets = self.n_targets
        a.targets = _ptr(self.targets)
        a.a_rows, a.a_indptr, a.a_indices, a.a_data = A.n_rows, _ptr(A.indptr), _ptr(A.indices), _ptr(A.data)
        a.b_rows, a.n_cols = B.n_rows, self.n_cols
        a.b_indptr, a.b_indices, a.b_data = _ptr(B.indptr), _ptr(B.indices), _ptr(B.data)
        a.Xtversky, a.Ytversky = _ptr(v.get("Xt")), _ptr(v.get("Yt"))
        a.Xcosine, a.Ycosine = _ptr(v.get("Xc")), _ptr(v.get("Yc"))
        a.Xdepop, a.Ydepop = _ptr(v.get("Xd")), _ptr(v.get("Yd"))
        for name in ("a1", "l1", "l2", "l3", "t1", "t2", "stabilized_shrink", "bayesian_shrink", "threshold"):
            setattr(a, name, P[name])
        a.k = self.k
        a.filter_mode, a.filter_indptr, a.filter_indices = self.filter_mode, _ptr(self.filter_m[0]), _ptr(self.filter_m[1])
        a.target_mode, a.target_indptr, a.target_indices = self.target_mode, _ptr(self.target_m[0]), _ptr(self.target_m[1])
        a.threads = int(self.tuning.get("threads", 0))
        a.panel_width = int(self.tuning.get("panel_width", 0))
        a.group = int(self.tuning.get("group", 0))
        a.b_nnz = B.nnz
        a.a_nnz = A.nnz
        a.unit_values = 1 if (self.unit_values and self.tuning.get("unit_values", True)) else 0  # (tuning: A/B measurements)
        if self.tuning.get("tie_mode", "deterministic") == "reference":
            self._alloc_outputs(a)
            self.args = a
            return
        # "engine": that kernel generation or an error; "engine_prefer": that one when it covers the configuration
        eng = self.tuning.get("engine", 0)
        prefer = self.tuning.get("engine_prefer") if not eng else None
        want = prefer if prefer is not None else eng
        a.engine = _lib.ENGINES[want] if isinstance(want, str) else int(want)
        # "drain_warps": 8 or 16, the build of the stream kernel (default: the planner's choice from the products per panel)
        if a.engine == _lib.ENGINE_STREAM and self.tuning.get("drain_warps"):
            a.group = int(self.tuning["drain_warps"])
        rc = lib.spy_knn_plan(C.byref(a), ctx.index)
        if rc == _lib.ERR_UNSUPPORTED and prefer is not None:
            a.engine = _lib.ENGINE_AUTO if _lib.ENGINES.get(prefer, prefer) == _lib.ENGINE_FLAT else _lib.ENGINE_FLAT
            a.group = int(self.tuning.get("group", 0))
            rc = lib.spy_knn_plan(C.byref(a), ctx.index)
        _lib.check(rc)
        if a.n_panels > 1:
            # the panel split needs ascending columns inside every row of B (a sorted CLONE when it is not: the
            # caller's handle is never permuted)
            B = self.B = sorted_rows(ctx, B)
            a.b_indptr, a.b_indices, a.b_data = _ptr(B.indptr), _ptr(B.indices), _ptr(B.data)

            def build_split():
                split = ctx.empty(B.n_rows * a.split_stride, torch.int32)
                _lib.check(lib.spy_knn_build_split_dev(B.n_rows, _ptr(B.indptr), _ptr(B.indices), a.panel_width, a.n_panels,
                                                       a.split_stride, _ptr(split), ctx.sptr))
                return split
            split = B.cached(("split", int(a.panel_width), int(a.n_panels), int(a.split_stride)), build_split)
            a.b_split = _ptr(split)
            self.keep.append(split)
        if a.engine == _lib.ENGINE_STREAM:
            n_p, stride = int(a.n_panels), int(a.split_stride)
            split_ptr = a.b_split

            def build_chunks():  # every (row, panel) segment of B as whole 16-byte chunks of two (column, value) pairs
                n_seg = B.n_rows * n_p
                cnt = ctx.empty(max(n_seg, 1), torch.int32)[: n_seg]
                _lib.check(lib.spy_knn_chunk_counts_dev(B.n_rows, _ptr(B.indptr), split_ptr, stride, n_p, _ptr(cnt), ctx.sptr))
                chunk_indptr = ctx.scan_i32(cnt)
                n_chunks = int(chunk_indptr[-1].item()) if n_seg > 0 else 0
                chunks = ctx.empty(max(n_chunks, 1) * 4, torch.int32)
                _lib.check(lib.spy_knn_pad_chunks_dev(B.n_rows, _ptr(B.indptr), _ptr(B.indices), _ptr(B.data), split_ptr, stride, n_p,
                                                      _ptr(chunk_indptr), _ptr(chunks), ctx.sptr, int(a.panel_width)))
                return chunk_indptr, chunks
            chunk_indptr, chunks = B.cached(("chunks", int(a.panel_width), n_p), build_chunks)
            a.b_chunk_indptr, a.b_chunks = _ptr(chunk_indptr), _ptr(chunks)

            def build_tables():  # chunk range of every (entry of a target row, panel)
                rl = ctx.empty(max(self.n_targets, 1), torch.int32)[: self.n_targets]
                _lib.check(lib.spy_knn_row_lengths_dev(self.n_targets, _ptr(self.targets), _ptr(A.indptr), _ptr(rl), ctx.sptr))
                toff = ctx.scan_i64(rl)
                n_entries = int(toff[-1].item()) if self.n_targets > 0 else 0
                aexp = ctx.empty(max(n_entries, 1) * int(a.n_panels) * 2, torch.int32)
                a.toff, a.n_entries, a.aexp = _ptr(toff), n_entries, _ptr(aexp)
                _lib.check(lib.spy_knn_build_aexp_dev(C.byref(a), ctx.sptr))
                return toff, n_entries, aexp
            if self.targets_key is not None:  # all rows, or this rank's contiguous range of them: reusable
                toff, n_entries, aexp = A.cached(("aexp", self.targets_key, int(a.panel_width), int(a.n_panels)), build_tables, partner=B)
            else:
                toff, n_entries, aexp = build_tables()
            a.toff, a.n_entries, a.aexp = _ptr(toff), n_entries, _ptr(aexp)
            self.keep += [chunk_indptr, chunks, toff, aexp]
        else:
            def build_pairs():  # 8-byte (column, value) stream layout of B
                pairs = ctx.empty((max(B.nnz, 1) + 1) * 2, torch.int32)  # +1: the kernel reads 16-byte words
                _lib.check(lib.spy_knn_pack_pairs_dev(B.nnz, _ptr(B.indices), _ptr(B.data), _ptr(pairs), ctx.sptr))
                return pairs
            pairs = B.cached("pairs", build_pairs)
            a.b_pairs = _ptr(pairs)
            self.keep.append(pairs)
        self._alloc_outputs(a)
        sb = int(lib.spy_knn_scratch_bytes(C.byref(a), ctx.index))
        if sb < 0:
            _lib.check(sb)
        self.scratch = ctx.empty(sb, torch.uint8)
        self.scratch_bytes = sb
        self.args = a

    def _alloc_outputs(self, a):
        ctx, torch = self.ctx, self.ctx.torch
        slab = self.n_targets * self.k
        if self.exchange is not None:  # the kernel writes into this rank's slice of the all-gather buffer
            self.out_cols, self.out_vals, self.out_counts = self.exchange.local()
        else:
            self.out_cols = ctx.empty(slab, torch.int32)
            self.out_vals = ctx.empty(slab, torch.float32)
            self.out_counts = ctx.empty(max(self.n_targets, 1), torch.int32)
        a.out_rows = None
        a.out_cols, a.out_values, a.out_counts = _ptr(self.out_cols), _ptr(self.out_vals), _ptr(self.out_counts)

    # ---- the reference's own order (tuning={"tie_mode": "reference"}) --------------------------------------------
    def run_reference_order(self, block_size):
        """Same result as run(), computed in the reference's order (csrc/knn_reforder.cu): float sums entry by entry of the
        target row, ties at the k-th value resolved like its heap -- including the blocked path, where matrix2's columns
        are permuted by popularity first (s_plus.pyx:218-225, 308-346; s_plus_utils.pyx:493-618) and the heap compares the
        permuted ids.  Slower by design; for callers that need the reference's index sets on tie-heavy data."""
        ctx, lib, torch = self.ctx, self.ctx.lib, self.ctx.torch
        a, n_cols = self.args, self.n_cols
        A = sorted_rows(ctx, self.A)  # scipy's tocsr of matrix1 yields ascending columns; an unsorted GPU transposition does not
        B = sorted_rows(ctx, self.B) if self.B.scatter_order else self.B  # otherwise the caller's stored order, like the reference
        bs = 0 if block_size is None else (262144 if block_size == 0 else int(block_size))  # s_plus.pyx:218-225, s_plus.h:33
        blocking = bs > 0 and n_cols > bs
        keep = [A, B]
        back = None
        v = dict(self.vectors)
        f_m, t_m = self.filter_m, self.target_m
        if blocking:  # _reorder_columns_by_popularity
            col_nnz = torch.bincount(B.indices.long(), minlength=n_cols)
            back = torch.sort(col_nnz, descending=True, stable=True).indices.to(torch.int32)
            if bool((back == torch.arange(n_cols, device=back.device, dtype=torch.int32)).all()):
                back = None
            else:
                fwd = torch.empty(n_cols, dtype=torch.int32, device=back.device)
                fwd[back.long()] = torch.arange(n_cols, device=back.device, dtype=torch.int32)

                def permuted(indptr, indices, data, n_rows):
                    m = DeviceCSR(n_rows, n_cols, indptr, fwd[indices.long()].contiguous(),
                                  data.clone() if data is not None else ctx.zeros(indices.numel(), torch.float32))
                    _lib.check(lib.spy_csr_sort_rows_dev(n_rows, _ptr(m.indptr), _ptr(m.indices), _ptr(m.data), ctx.sptr))
                    return m
                B = permuted(B.indptr, B.indices, B.data, B.n_rows)
                for name in ("Yt", "Yc", "Yd"):
                    if name in v:
                        v[name] = v[name][back.long()].contiguous()
                if self.filter_mode == MODE_MATRIX:
                    m = permuted(f_m[0], f_m[1], None, self.n_rows)
                    f_m = (m.indptr, m.indices)
                if self.target_mode == MODE_MATRIX:
                    m = permuted(t_m[0], t_m[1], None, self.n_rows)
                    t_m = (m.indptr, m.indices)
                keep += [B, fwd]
        a.a_indptr, a.a_indices, a.a_data = _ptr(A.indptr), _ptr(A.indices), _ptr(A.data)
        a.b_indptr, a.b_indices, a.b_data = _ptr(B.indptr), _ptr(B.indices), _ptr(B.data)
        a.Ytversky, a.Ycosine, a.Ydepop = _ptr(v.get("Yt")), _ptr(v.get("Yc")), _ptr(v.get("Yd"))
        a.filter_indptr, a.filter_indices = _ptr(f_m[0]), _ptr(f_m[1])
        a.target_indptr, a.target_indices = _ptr(t_m[0]), _ptr(t_m[1])
        sb = int(lib.spy_knn_reforder_scratch_bytes(C.byref(a)))
        if sb < 0:
            _lib.check(sb)
        scratch = ctx.empty(sb, torch.uint8)
        _lib.check(lib.spy_knn_topk_reforder_dev(C.byref(a), bs if blocking else 0, _ptr(back), _ptr(scratch), sb, ctx.sptr))
        ctx.sync()  # the permuted operands and the scratch are released on return
        del keep, v, f_m, t_m

    # ---- the hot kernel ---------------------------------------------------------------------------
    def run(self):
        trace = KERNEL_TRACE
        if trace is not None:  # bench / profiling: CUDA events on the launching stream around the hot kernel only
            ev0 = self.ctx.torch.cuda.Event(enable_timing=True)
            ev1 = self.ctx.torch.cuda.Event(enable_timing=True)
            ev0.record(self.ctx.stream)
        _lib.check(self.ctx.lib.spy_knn_topk_dev(C.byref(self.args), _ptr(self.scratch), self.scratch_bytes, self.ctx.sptr))
        if trace is not None:
            ev1.record(self.ctx.stream)
            trace.append(dict(start=ev0, end=ev1, n_targets=self.n_targets, k=self.k, n_panels=int(self.args.n_panels),
                              panel_width=int(self.args.panel_width), threads=int(self.args.threads),
                              group=int(self.args.group), engine=int(self.args.engine)))

    # ---- output (s_plus.pyx:405-424) --------------------------------------------------------------
    def assemble_device(self, format_output):
        """Device part of the output assembly; returns device tensors ready for the D2H copy."""
        ctx, lib, torch = self.ctx, self.ctx.lib, self.ctx.torch
        k, nt = self.k, self.n_targets
        if format_output == "coo":
            self.compact_padded()
            nt = self.n_targets
            rows = ctx.empty(nt * k, torch.int32)
            _lib.check(lib.spy_slab_fill_rows_dev(nt, k, _ptr(self.targets), _ptr(self.out_counts), _ptr(rows), ctx.sptr))
            return ("coo", rows, self.out_cols, self.out_vals)
        if not self.unique_targets:
            return ("slab", self.out_cols, self.out_vals, self.out_counts)
        row_nnz = ctx.zeros(max(self.n_rows, 1), torch.int32)[: self.n_rows]
        _lib.check(lib.spy_slab_row_nnz_dev(nt, k, _ptr(self.out_vals), _ptr(self.out_counts), _ptr(self.targets),
                                            _ptr(row_nnz), ctx.sptr))
        indptr = ctx.scan_i64(row_nnz)
        nnz = int(indptr[-1].item())
        # get_index_dtype(max(len(values), n_cols)) on the PADDED slab length, utils.pyx:165-168
        n_full = nt if self.targets_np is None or self.exchange is None else int(self.targets_np.shape[0])
        idx64 = max(n_full * k, self.n_cols) > INT32_MAX
        indices = ctx.empty(nnz, torch.int64 if idx64 else torch.int32)
        data = ctx.empty(nnz, torch.float32)
        _lib.check(lib.spy_slab_compact_dev(nt, k, _ptr(self.out_cols), _ptr(self.out_vals), _ptr(self.out_counts),
                                            _ptr(self.targets), _ptr(indptr), _ptr(indices),
                                            _lib.I64 if idx64 else _lib.I32, _ptr(data), ctx.sptr))
        if not idx64:
            indptr = indptr.to(torch.int32)
        return ("csr", indptr, indices, data)

    def to_device_matrix(self) -> DeviceMatrix:
        """The result as a CSR DeviceMatrix (zero-free, rows in best-first order, like format_output='csr')."""
        if not self.unique_targets:
            raise ValueError("on_device=True needs unique target_rows")
        if self.n_targets == 0:
            z = self.ctx.zeros(self.n_rows + 1, self.ctx.torch.int32)
            e = self.ctx.empty(0, self.ctx.torch.int32)
            return DeviceMatrix(DeviceCSR(self.n_rows, self.n_cols, z, e, self.ctx.empty(0, self.ctx.torch.float32)), False)
        _, indptr, indices, data = self.assemble_device("csr")
        if indices.dtype != self.ctx.torch.int32:
            raise ValueError("result exceeds int32 indexing; fetch it with on_device=False")
        return DeviceMatrix(self.ctx.produced(DeviceCSR(self.n_rows, self.n_cols, indptr, indices, data, sorted_rows=False)), False)

    def to_host(self, assembled):
        ctx = self.ctx
        kind = assembled[0]
        shape = (self.n_rows, self.n_cols)
        host = [ctx.d2h(t) for t in assembled[1:]]
        ctx.sync()
        host = [h.numpy() for h in host]
        if kind == "coo":
            rows, cols, vals = host
            return sp.coo_array((vals, (rows, cols)), shape=shape, dtype=np.float32)
        if kind == "csr":
            indptr, indices, data = host
            return sp.csr_array((data, indices, indptr), shape=shape, dtype=np.float32)
        # duplicate target rows: stable counting sort by row on the host, duplicates carried over
        cols, vals, counts = host
        k = self.k
        targets = self.targets.cpu().numpy()
        valid = (np.arange(k, dtype=np.int64)[None, :] < counts[: self.n_targets, None]).ravel()
        rows = np.repeat(targets, k)[valid]
        cols, vals = cols[valid], vals[valid]
        nz = vals != 0
        rows, cols, vals = rows[nz], cols[nz], vals[nz]
        order = np.argsort(rows, kind="stable")
        indptr = np.zeros(self.n_rows + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=self.n_rows), out=indptr[1:])
        idx_dtype = np.int64 if max(self.n_targets * k, self.n_cols) > INT32_MAX else np.int32
        return sp.csr_array((vals[order], cols[order].astype(idx_dtype), indptr.astype(idx_dtype)), shape=shape, dtype=np.float32)


def prepare_job(matrix1, matrix2=None, weight_depop_matrix1="none", weight_depop_matrix2="none",
                p1=0.0, p2=0.0, a1=1.0, l1=0.0, l2=0.0, l3=0.0, t1=1.0, t2=1.0, c1=0.5, c2=0.5, k=100,
                stabilized_shrink=0.0, bayesian_shrink=0.0, additive_shrink=0.0, threshold=0.0, binary=False,
                target_rows=None, filter_cols=None, target_cols=None, verbose=True, format_output="csr",
                num_threads=0, block_size=0, device=None, tuning=None) -> KnnJob:
    """Validate, upload and pre-process: everything up to (not including) the hot kernel."""
    matrix2_given = matrix2
    if matrix2 is None:  # s_plus.pyx:169-170
        matrix2 = matrix1.T if _is_matrix(matrix1) else None
    validate_inputs(matrix1, matrix2, weight_depop_matrix1, weight_depop_matrix2, k, target_rows,
                    filter_cols, target_cols, verbose, format_output)
    k = int(min(k, matrix2.shape[1]))  # s_plus.pyx:187-188
    n_rows, n_cols = int(matrix1.shape[0]), int(matrix2.shape[1])
    if target_rows is None:  # s_plus.pyx:191-196
        targets_np = np.arange(n_rows, dtype=np.int32)
        unique = True
    else:
        targets_np = np.ascontiguousarray(np.asarray(target_rows, dtype=np.int32))
        unique = np.unique(targets_np).shape[0] == targets_np.shape[0]
        if targets_np.size and (targets_np.min() < 0 or targets_np.max() >= n_rows):
            # undefined behaviour in the reference (no bounds check, s_plus.h:344-346); refuse instead
            raise ValueError(f"target_rows must lie in [0, {n_rows})")
    ctx = Ctx(device)
    array_mode = selector_mode(filter_cols) == MODE_ARRAY or selector_mode(target_cols) == MODE_ARRAY
    A, B = upload_pair(ctx, matrix1, matrix2_given)
    if isinstance(A, WideCSR) or isinstance(B, WideCSR):
        raise ValueError("a matrix beyond int32 stored entries is computed block by block: call s_plus (or a similarity function)")
    raw_b = B.data if (binary and array_mode) else None
    if binary:
        A, B = binarize(ctx, A), binarize(ctx, B)
    f32 = lambda x: float(np.float32(x))
    params = dict(a1=f32(a1), l1=f32(l1), l2=f32(l2), l3=f32(l3), t1=f32(t1), t2=f32(t2),
                  stabilized_shrink=f32(stabilized_shrink), bayesian_shrink=f32(bayesian_shrink), threshold=f32(threshold))
    job = KnnJob(ctx=ctx, A=A, B=B, targets=ctx.h2d(targets_np), n_targets=int(targets_np.shape[0]), k=k,
                 n_rows=n_rows, n_cols=n_cols, params=params, unique_targets=unique, tuning={**DEFAULT_TUNING, **dict(tuning or {})},
                 targets_key=("all", n_rows) if target_rows is None else None,
                 unit_values=bool(binary) and raw_b is None)  # (list-mode selectors restore B's raw values: the a11 quirk)
    job.build_vectors(weight_depop_matrix1, weight_depop_matrix2, f32(p1), f32(p2), f32(c1), f32(c2), f32(additive_shrink))
    job.build_selectors(filter_cols, target_cols, raw_b)
    spec = _sharded.active()
    if spec is not None:
        job.shard_targets(spec, targets_np)
    job.plan()
    return job


def _is_wide(m) -> bool:
    """True for an operand whose stored entries do not fit int32 indexing (scipy matrix or DeviceMatrix)."""
    if isinstance(m, DeviceMatrix):
        return isinstance(m.stored, WideCSR)
    return sp.issparse(m) and m.nnz > WIDE_NNZ_LIMIT


def _s_plus_wide(matrix1, matrix2, kw, target_rows, device, on_device):
    """s_plus for operands with more than WIDE_NNZ_LIMIT stored entries.

    Target rows are independent (s_plus.h:337-451), so matrix1 is taken one block of ROWS at a time; and the k best
    columns of a row are the k best among the k best of every block of COLUMNS of matrix2 (TopK keeps the k largest
    values, s_plus.h:45-59), so matrix2 is taken one block of columns at a time and the slabs are merged
    (spy_slab_merge_dev).  Every block keeps the shape of its matrix (wide_blocks): ids, norm vectors, depop weights and
    selectors are passed through unchanged, and each (row block, column block) call is an ordinary int32 job -- the
    column sums of a column block and the row sums of a row block are complete, which is all computeSimilarity needs
    (s_plus.h:129-156)."""
    m2 = matrix2 if matrix2 is not None else matrix1.T
    validate_inputs(matrix1, m2, kw["weight_depop_matrix1"], kw["weight_depop_matrix2"], kw["k"], target_rows,
                    kw["filter_cols"], kw["target_cols"], kw["verbose"], kw["format_output"])
    if _sharded.active() is not None:
        raise NotImplementedError("sharded calls on matrices beyond int32 stored entries: shard the target rows yourself (target_rows)")
    if dict(kw["tuning"] or {}).get("tie_mode", DEFAULT_TUNING.get("tie_mode", "deterministic")) == "reference":
        raise NotImplementedError("tie_mode='reference' is limited to int32-indexed matrices, like the reference itself")
    for name in ("filter_cols", "target_cols"):
        if _is_wide(kw[name]):
            raise NotImplementedError(f"{name} beyond int32 stored entries is not supported")
    ctx = Ctx(device)
    torch, lib = ctx.torch, ctx.lib
    s1, t1 = upload_stored(ctx, matrix1)
    s2, t2 = (s1, not t1) if matrix2 is None else upload_stored(ctx, matrix2)
    a_blocks = operand_blocks(ctx, s1, t1, axis=0)
    b_blocks = operand_blocks(ctx, s2, t2, axis=1)
    n_rows, n_cols = int(matrix1.shape[0]), int(m2.shape[1])
    k = int(min(kw["k"], n_cols))
    if target_rows is None:
        targets_np = np.arange(n_rows, dtype=np.int32)
    else:
        targets_np = np.ascontiguousarray(np.asarray(target_rows, dtype=np.int32))
        if targets_np.size and (targets_np.min() < 0 or targets_np.max() >= n_rows):
            raise ValueError(f"target_rows must lie in [0, {n_rows})")
    n_t = int(targets_np.shape[0])
    unique = target_rows is None or np.unique(targets_np).shape[0] == n_t
    whole = len(a_blocks) == 1
    slab = None if whole else (ctx.zeros(n_t * k, torch.int32), ctx.zeros(n_t * k, torch.float32), ctx.zeros(max(n_t, 1), torch.int32))
    for r0, r1, a_blk in a_blocks:
        if whole:
            sel, rows = None, target_rows
        else:
            sel = np.nonzero((targets_np >= r0) & (targets_np < r1))[0]
            rows = targets_np[sel]
        n_loc = n_t if whole else int(sel.shape[0])
        if n_loc == 0:
            continue
        acc = None
        for _, _, b_blk in b_blocks:
            job = prepare_job(a_blk, b_blk, target_rows=rows, device=ctx.device, **kw)
            job.run()
            cur = (job.out_cols, job.out_vals, job.out_counts)
            if acc is None:
                acc = cur
            else:
                out = (ctx.empty(n_loc * k, torch.int32), ctx.empty(n_loc * k, torch.float32), ctx.empty(n_loc, torch.int32))
                _lib.check(lib.spy_slab_merge_dev(n_loc, k, _ptr(acc[0]), _ptr(acc[1]), _ptr(acc[2]), _ptr(cur[0]), _ptr(cur[1]),
                                                  _ptr(cur[2]), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), ctx.sptr))
                acc = out
            ctx.sync()  # the block job's tables are released before the next block builds its own
            del job
        if whole:
            slab = acc
        else:
            at = ctx.h2d(sel.astype(np.int64))
            slab[0].view(n_t, k).index_copy_(0, at, acc[0].view(n_loc, k))
            slab[1].view(n_t, k).index_copy_(0, at, acc[1].view(n_loc, k))
            slab[2][:n_t].index_copy_(0, at, acc[2][:n_loc])
    if slab is None:  # no target rows at all
        slab = (ctx.zeros(0, torch.int32), ctx.zeros(0, torch.float32), ctx.zeros(1, torch.int32))
    fin = KnnJob(ctx=ctx, A=None, B=None, targets=ctx.h2d(targets_np), n_targets=n_t, k=k, n_rows=n_rows, n_cols=n_cols,
                 params={}, unique_targets=unique)
    fin.out_cols, fin.out_vals, fin.out_counts = slab
    if on_device:
        return fin.to_device_matrix()
    return fin.to_host(fin.assemble_device(kw["format_output"]))


@preserve_device
def s_plus(matrix1, matrix2=None, weight_depop_matrix1="none", weight_depop_matrix2="none",
           p1=0.0, p2=0.0, a1=1.0, l1=0.0, l2=0.0, l3=0.0, t1=1.0, t2=1.0, c1=0.5, c2=0.5, k=100,
           stabilized_shrink=0.0, bayesian_shrink=0.0, additive_shrink=0.0, threshold=0.0, binary=False,
           target_rows=None, filter_cols=None, target_cols=None, verbose=True, format_output="csr",
           num_threads=0, block_size=0, device=None, tuning=None, on_device=False):
    """Top-K similarity between the rows of matrix1 and the columns of matrix2.

    Same arguments, defaults and result as the reference's ``cython_code.s_plus.s_plus``
    (s_plus.pyx:95-123).  ``num_threads`` and ``block_size`` are accepted for compatibility: they
    steer the reference's OpenMP team and CPU-cache blocking and have no meaning on the GPU
    (the shared-memory panel width is planned by the library).  ``verbose`` is validated and ignored.
    """
    if _is_wide(matrix1) or _is_wide(matrix2):
        return _s_plus_wide(matrix1, matrix2, dict(
            weight_depop_matrix1=weight_depop_matrix1, weight_depop_matrix2=weight_depop_matrix2, p1=p1, p2=p2, a1=a1, l1=l1,
            l2=l2, l3=l3, t1=t1, t2=t2, c1=c1, c2=c2, k=k, stabilized_shrink=stabilized_shrink, bayesian_shrink=bayesian_shrink,
            additive_shrink=additive_shrink, threshold=threshold, binary=binary, filter_cols=filter_cols, target_cols=target_cols,
            verbose=verbose, format_output=format_output, num_threads=num_threads, block_size=block_size, tuning=tuning),
            target_rows, device, on_device)
    job = prepare_job(matrix1, matrix2, weight_depop_matrix1, weight_depop_matrix2, p1, p2, a1, l1, l2, l3, t1, t2,
                      c1, c2, k, stabilized_shrink, bayesian_shrink, additive_shrink, threshold, binary,
                      target_rows, filter_cols, target_cols, verbose, format_output, num_threads, block_size,
                      device, tuning)
    if job.n_targets > 0:
        if job.tuning.get("tie_mode", "deterministic") == "reference":
            job.run_reference_order(block_size)
        else:
            job.run()
    if job.exchange is not None:
        job.gather()
    if on_device:
        return job.to_device_matrix()
    return job.to_host(job.assemble_device(format_output))


@preserve_device
def to_device(matrix, device=None) -> DeviceMatrix:
    """Upload a scipy sparse matrix once (zero-free, float32 values, int32 indices) and return a handle the
    similarity functions accept in place of ``matrix1`` / ``matrix2``."""
    if isinstance(matrix, DeviceMatrix):
        return matrix
    if not sp.issparse(matrix):
        raise TypeError("matrix must be a sparse matrix")
    ctx = Ctx(device)
    stored, transposed = upload_stored(ctx, matrix)
    ctx.sync()
    return DeviceMatrix(stored, transposed)


@preserve_device
def axis_sum(m: DeviceMatrix, axis: int):
    """``matrix.sum(axis)`` of a DeviceMatrix as a float32 device tensor (similarity.py:479)."""
    ctx = Ctx(m.device)
    torch, lib, s = ctx.torch, ctx.lib, ctx.consume(m.stored)
    along_stored_rows = (axis == 1) != m.transposed  # summing over the stored minor axis
    if isinstance(s, WideCSR):  # block by block: every block has the shape of the matrix, the sums add up
        total = None
        for _, _, b in wide_blocks(ctx, s, along_major=True):
            part = axis_sum(DeviceMatrix(b, m.transposed), axis)
            total = part if total is None else total.add_(part)
        return total
    if along_stored_rows:
        out = ctx.empty(s.n_rows, torch.float32)
        _lib.check(lib.spy_csr_row_sum_dev(s.n_rows, _ptr(s.indptr), _ptr(s.data), 0, _ptr(out), ctx.sptr))
    else:
        out = ctx.empty(s.n_cols, torch.float32)
        acc = ctx.empty(max(s.n_cols, 1), torch.float64)
        _lib.check(lib.spy_csr_col_sum_dev(s.nnz, _ptr(s.indices), _ptr(s.data), 0, s.n_cols, _ptr(acc), _ptr(out), ctx.sptr))
        ctx.sync()
    return out


@preserve_device
def pow_values_(m: DeviceMatrix, p: float) -> DeviceMatrix:
    """``m.data = m.data ** p`` in place on the device (similarity.py:411-415, 480-483)."""
    ctx = Ctx(m.device)
    d = m.stored.data
    ctx.consume(m.stored)
    _lib.check(ctx.lib.spy_pow_shift_dev(d.numel(), _ptr(d), _lib.F32, 0.0, float(p), _ptr(d), ctx.sptr))
    ctx.produced(m.stored)
    m.stored.invalidate()  # the values changed: cached transposes / norms / stream layouts are stale
    return m


def to_host(m: DeviceMatrix):
    """DeviceMatrix -> scipy csr_array (or csc_array when the handle is a transposed view)."""
    s = m.stored
    arrs = [t.cpu().numpy() for t in (s.data, s.indices, s.indptr)]
    if arrs[1].dtype != arrs[2].dtype:  # a 64-bit indptr: scipy wants one index dtype
        arrs[1] = arrs[1].astype(arrs[2].dtype)
    if m.transposed:
        return sp.csc_array(tuple(arrs), shape=m.shape, dtype=np.float32)
    return sp.csr_array(tuple(arrs), shape=m.shape, dtype=np.float32)
