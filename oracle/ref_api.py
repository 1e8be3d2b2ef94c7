"""Thin access layer to the compiled, UNMODIFIED reference under oracle/_ref -- TEST INFRASTRUCTURE ONLY.

``oracle/_ref/similaripy_ref/cython_code/*.so`` are the reference's own Cython modules
(s_plus.pyx, s_plus_utils.pyx, normalization.pyx, utils.pyx) built by build_ref.py with the
reference's flags.  The reference's pure-Python wrappers (similaripy/similarity.py,
similaripy/normalization.py) only choose constants; they are restated, not copied, by
``oracle.similarity(..., core=...)`` and by the small functions below, so that nothing here
needs /root/reference at run time (the GPU box does not have it).
"""
from __future__ import annotations

import importlib
import os
import sys
from math import e as _E

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")


def available() -> bool:
    from . import build_ref
    return build_ref.built()


def _mod(name):
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    return importlib.import_module(f"similaripy_ref.cython_code.{name}")


def s_plus_core(*args, **kwargs):
    """The reference's Cython driver s_plus.s_plus (s_plus.pyx:95-433) itself."""
    return _mod("s_plus").s_plus(*args, **kwargs)


def num_threads() -> int:
    return int(_mod("utils").get_num_threads())


def _prepare_csr(X, axis, inplace):  # restates normalization.py:23-66
    if axis not in (0, 1):
        raise ValueError(f"axis must be 0 or 1, got {axis}")
    if not sp.issparse(X):
        raise TypeError("X must be a sparse matrix")
    if X.data.dtype not in (np.float32, np.float64):
        X = sp.csr_array(X, dtype=np.float32)
    if not inplace:
        X = X.copy()
    if axis == 0:
        X = X.T
    return X.tocsr()


def _fin(X, axis):
    return (X.T if axis == 0 else X).tocsr()


def normalize(X, norm="l2", axis=1, inplace=False):
    X = _prepare_csr(X, axis, inplace)
    getattr(_mod("normalization"), f"inplace_normalize_csr_{norm}")(
        shape=X.shape, data=X.data, indices=X.indices, indptr=X.indptr)
    return _fin(X, axis)


def bm25plus(X, axis=1, k1=1.2, b=0.75, delta=1.0, logbase=_E, tf_mode="raw", idf_mode="bm25", inplace=False):
    X = _prepare_csr(X, axis, inplace)
    _mod("normalization").inplace_normalize_csr_bm25plus(
        shape=X.shape, data=X.data, indices=X.indices, indptr=X.indptr, k1=k1, b=b, delta=delta,
        tf_mode=tf_mode, idf_mode=idf_mode, logbase=logbase)
    return _fin(X, axis)


def bm25(X, axis=1, k1=1.2, b=0.75, logbase=_E, tf_mode="raw", idf_mode="bm25", inplace=False):
    return bm25plus(X, axis, k1, b, 0.0, logbase, tf_mode, idf_mode, inplace)


def tfidf(X, axis=1, logbase=_E, tf_mode="sqrt", idf_mode="smooth", inplace=False):
    X = _prepare_csr(X, axis, inplace)
    _mod("normalization").inplace_normalize_csr_tfidf(
        shape=X.shape, data=X.data, indices=X.indices, indptr=X.indptr,
        tf_mode=tf_mode, idf_mode=idf_mode, logbase=logbase)
    return _fin(X, axis)


def similarity(name, matrix1, matrix2=None, **kw):
    """One of the nine public similarities, computed by the compiled reference."""
    from . import oracle
    kw.setdefault("verbose", False)
    return oracle.similarity(name, matrix1, matrix2, core=s_plus_core, normalize_fn=normalize, **kw)
