"""Public normalization functions -- drop-in for ``similaripy.normalization`` (reference
similaripy/normalization.py:91-218).  The row loops of normalization.pyx run as CUDA kernels
(csrc/normalize.cu) on the matrix's own value dtype (float32 / float64) and index dtype
(int32 / int64); the Python side reproduces ``_prepare_csr`` / ``_finalize_csr`` exactly:
dtype coercion, copy unless ``inplace``, ``axis=0`` via transpose, always a CSR result.
"""
from __future__ import annotations

from math import e

import numpy as np
import scipy.sparse as sps

from . import _lib
from ._engine import Ctx, DeviceCSR, DeviceMatrix, WideCSR, _ptr, preserve_device, transpose_csr

_NORMALIZATIONS = ("l1", "l2", "max")
_TF_MODES = ("binary", "raw", "sqrt", "freq", "log")
_IDF_MODES = ("unary", "base", "smooth", "prob", "bm25")


def _check_matrix(X):
    """normalization.py:23-40."""
    if not sps.issparse(X):
        raise TypeError("X must be a sparse matrix")
    if X.data.dtype not in (np.float32, np.float64):
        X = sps.csr_array(X, dtype=np.float32)
    return X


def _prepare_csr(X, axis: int, inplace: bool):
    """normalization.py:43-66."""
    if axis not in (0, 1):
        raise ValueError(f"axis must be 0 or 1, got {axis}")
    X = _check_matrix(X)
    if not inplace:
        X = X.copy()
    if axis == 0:
        X = X.T
    return X.tocsr()


def _finalize_csr(X, axis: int):
    """normalization.py:69-73."""
    if axis == 0:
        X = X.T
    return X.tocsr()


def _validate_modes(tf_mode: str, idf_mode: str) -> None:
    """normalization.py:76-86."""
    if tf_mode not in _TF_MODES:
        raise ValueError(f"tf_mode must be one of {_TF_MODES}, got '{tf_mode}'")
    if idf_mode not in _IDF_MODES:
        raise ValueError(f"idf_mode must be one of {_IDF_MODES}, got '{idf_mode}'")


class _DeviceRows:
    """The CSR arrays of X on the device; ``finish`` copies the (in-place modified) values back into X.data."""

    def __init__(self, X, device, need_indices: bool):
        self.ctx = ctx = Ctx(device)
        self.X = X
        if X.indptr.dtype not in (np.int32, np.int64):  # Cython's `integral` also takes int16; widen it
            X.indptr = X.indptr.astype(np.int32)
            X.indices = X.indices.astype(np.int32)
        if X.indices.dtype != X.indptr.dtype:
            X.indices = X.indices.astype(X.indptr.dtype)
        self.val_code = _lib.F32 if X.data.dtype == np.float32 else _lib.F64
        self.idx_code = _lib.I32 if X.indptr.dtype == np.int32 else _lib.I64
        self.data = ctx.h2d(X.data)
        self.indptr = ctx.h2d(X.indptr)
        self.indices = ctx.h2d(X.indices) if need_indices else None

    def finish(self):
        ctx = self.ctx
        host = ctx.d2h(self.data)
        ctx.sync()
        out = host.numpy()
        if self.X.data.flags.writeable:
            self.X.data[...] = out  # in place: `inplace=True` callers hold references to X.data
        else:
            self.X.data = out.copy()


def _device_rows(X: DeviceMatrix, axis: int, inplace: bool):
    """DeviceMatrix counterpart of _prepare_csr: the CSR whose ROWS are to be normalised, with values that
    may be overwritten.  Returns (ctx, csr, transposed_flag_of_result)."""
    if axis not in (0, 1):
        raise ValueError(f"axis must be 0 or 1, got {axis}")
    ctx = Ctx(X.device)
    want_transposed = (axis == 0)  # axis=0 normalises the rows of X.T
    s = ctx.consume(X.stored)
    if isinstance(s, WideCSR):  # beyond int32 stored entries (64-bit indptr): the stored orientation only
        if X.transposed != want_transposed:
            raise NotImplementedError("a DeviceMatrix beyond int32 stored entries is normalised along its stored rows only; "
                                      "normalise the scipy matrix (64-bit indices are supported there) before to_device")
        if not inplace:
            s = WideCSR(s.n_rows, s.n_cols, s.indptr, s.indices, s.data.clone(), sorted_rows=s.sorted_rows)
        else:
            s.invalidate()
    elif X.transposed != want_transposed:  # stored orientation is the other one: transpose on the device (a copy)
        s = transpose_csr(ctx, s)
    elif not inplace:
        s = DeviceCSR(s.n_rows, s.n_cols, s.indptr, s.indices, s.data.clone(), sorted_rows=s.sorted_rows, scatter_order=s.scatter_order)
    else:
        s.invalidate()  # values are about to be overwritten in place: prepared operands cached on the handle are stale
    return ctx, s, want_transposed


def _device_weighting(X: DeviceMatrix, axis, inplace, bm25_args, tf_mode, idf_mode, logbase):
    _validate_modes(tf_mode, idf_mode)
    if isinstance(X.stored, WideCSR):
        raise NotImplementedError("tfidf / bm25 on a DeviceMatrix beyond int32 stored entries: weight the scipy matrix "
                                  "(64-bit indices are supported there) before to_device")
    ctx, s, flag = _device_rows(X, axis, inplace)
    lib = ctx.lib
    scratch = ctx.empty(lib.spy_tfidf_scratch_bytes(s.n_rows, s.n_cols, _lib.F32), ctx.torch.uint8)
    tf, idf = _lib.TF_MODES[tf_mode], _lib.IDF_MODES[idf_mode]
    if bm25_args is None:
        _lib.check(lib.spy_tfidf_dev(s.n_rows, s.n_cols, _ptr(s.data), _lib.F32, _ptr(s.indices), _ptr(s.indptr),
                                     _lib.I32, tf, idf, float(logbase), _ptr(scratch), ctx.sptr))
    else:
        k1, b, delta = bm25_args
        _lib.check(lib.spy_bm25plus_dev(s.n_rows, s.n_cols, _ptr(s.data), _lib.F32, _ptr(s.indices), _ptr(s.indptr),
                                        _lib.I32, float(k1), float(b), float(delta), tf, idf, float(logbase),
                                        _ptr(scratch), ctx.sptr))
    ctx.sync()  # scratch is released on return
    return DeviceMatrix(s, flag)


@preserve_device
def normalize(X, norm: str = "l2", axis: int = 1, inplace: bool = False, *, device=None):
    """Row (axis=1) or column (axis=0) l1 / l2 / max normalisation (normalization.py:91-113)."""
    if norm not in _NORMALIZATIONS:
        raise ValueError(f"norm must be one of {_NORMALIZATIONS}, got '{norm}'")
    if isinstance(X, DeviceMatrix):
        ctx, s, flag = _device_rows(X, axis, inplace)
        _lib.check(ctx.lib.spy_normalize_rows_dev(_NORMALIZATIONS.index(norm), s.n_rows, _ptr(s.data), _lib.F32,
                                                  _ptr(s.indptr), _lib.I64 if isinstance(s, WideCSR) else _lib.I32, ctx.sptr))
        return DeviceMatrix(ctx.produced(s), flag)
    X = _prepare_csr(X, axis, inplace)
    d = _DeviceRows(X, device, need_indices=False)
    _lib.check(d.ctx.lib.spy_normalize_rows_dev(_NORMALIZATIONS.index(norm), X.shape[0], _ptr(d.data), d.val_code,
                                                _ptr(d.indptr), d.idx_code, d.ctx.sptr))
    d.finish()
    return _finalize_csr(X, axis)


def _weighting(X, axis, inplace, device, bm25_args, tf_mode, idf_mode, logbase):
    if isinstance(X, DeviceMatrix):
        return _device_weighting(X, axis, inplace, bm25_args, tf_mode, idf_mode, logbase)
    _validate_modes(tf_mode, idf_mode)
    X = _prepare_csr(X, axis, inplace)
    d = _DeviceRows(X, device, need_indices=True)
    ctx, lib = d.ctx, d.ctx.lib
    n_rows, n_cols = X.shape
    scratch = ctx.empty(lib.spy_tfidf_scratch_bytes(n_rows, n_cols, d.val_code), ctx.torch.uint8)
    tf, idf = _lib.TF_MODES[tf_mode], _lib.IDF_MODES[idf_mode]
    if bm25_args is None:
        _lib.check(lib.spy_tfidf_dev(n_rows, n_cols, _ptr(d.data), d.val_code, _ptr(d.indices), _ptr(d.indptr),
                                     d.idx_code, tf, idf, float(logbase), _ptr(scratch), ctx.sptr))
    else:
        k1, b, delta = bm25_args
        _lib.check(lib.spy_bm25plus_dev(n_rows, n_cols, _ptr(d.data), d.val_code, _ptr(d.indices), _ptr(d.indptr),
                                        d.idx_code, float(k1), float(b), float(delta), tf, idf, float(logbase),
                                        _ptr(scratch), ctx.sptr))
    d.finish()
    return _finalize_csr(X, axis)


@preserve_device
def bm25(X, axis: int = 1, k1: float = 1.2, b: float = 0.75, logbase: float = e, tf_mode: str = "raw",
         idf_mode: str = "bm25", inplace: bool = False, *, device=None):
    """BM25 weighting (normalization.py:116-149): BM25+ with delta = 0."""
    return _weighting(X, axis, inplace, device, (k1, b, 0.0), tf_mode, idf_mode, logbase)


@preserve_device
def bm25plus(X, axis: int = 1, k1: float = 1.2, b: float = 0.75, delta: float = 1.0, logbase: float = e,
             tf_mode: str = "raw", idf_mode: str = "bm25", inplace: bool = False, *, device=None):
    """BM25+ weighting (normalization.py:152-187)."""
    return _weighting(X, axis, inplace, device, (k1, b, delta), tf_mode, idf_mode, logbase)


@preserve_device
def tfidf(X, axis: int = 1, logbase: float = e, tf_mode: str = "sqrt", idf_mode: str = "smooth",
          inplace: bool = False, *, device=None):
    """TF-IDF weighting (normalization.py:190-218)."""
    return _weighting(X, axis, inplace, device, None, tf_mode, idf_mode, logbase)
