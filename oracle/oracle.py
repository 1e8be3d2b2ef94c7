"""CPU oracle for the similaripy sparse-KNN hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``similaripy_b200`` must never import it (tests/test_boundary.py enforces that).

It restates, in numpy + the C file ``spy_oracle.c``, what the reference does on
this path; every function cites the reference ``file:line`` it follows (paths are
relative to /root/reference).  It is pinned against the reference itself
(``oracle/_ref``, built by ``build_ref.py``) in tests/test_oracle_vs_ref.py and
against the golden fixtures in tests/golden/.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from math import e as _E

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libspy_oracle.so")
_lib = None

MODE_NONE, MODE_ARRAY, MODE_MATRIX = 0, 1, 2
DEFAULT_BLOCK_SIZE = 262144  # s_plus.h:33

_TF = {"binary": 0, "raw": 1, "sqrt": 2, "freq": 3, "log": 4}       # normalization.pyx:12-17
_IDF = {"unary": 0, "base": 1, "smooth": 2, "prob": 3, "bm25": 4}   # normalization.pyx:19-24


def build(force: bool = False) -> str:
    """Compile spy_oracle.c into libspy_oracle.so (gcc -O2 -fopenmp, no fast-math)."""
    src = os.path.join(_HERE, "spy_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        cc = os.environ.get("ORACLE_CC", "/usr/bin/gcc")
        if not os.path.exists(cc):
            cc = "gcc"
        subprocess.check_call([cc, "-O2", "-fPIC", "-fopenmp", "-std=c99", "-shared", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.spy_oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- #
# s_plus_utils.pyx restatements                                                #
# --------------------------------------------------------------------------- #
def validate(matrix1, matrix2, w1, w2, k, target_rows, filter_cols, target_cols, verbose, format_output):
    """s_plus_utils.pyx:19-125 (same exception types, same order of checks)."""
    if not sp.issparse(matrix1):
        raise TypeError("matrix1 must be a sparse matrix")
    if not sp.issparse(matrix2):
        raise TypeError("matrix2 must be a sparse matrix")
    if matrix1.shape[1] != matrix2.shape[0]:
        raise ValueError("Incompatible matrix shapes")
    if k < 1:
        raise ValueError("k must be >= 1")
    if not (len(w1) == matrix1.shape[0] or w1 in ("none", "sum")):
        raise ValueError("weight_depop_matrix1")
    if not (len(w2) == matrix2.shape[1] or w2 in ("none", "sum")):
        raise ValueError("weight_depop_matrix2")
    if target_rows is not None and len(target_rows) > matrix1.shape[0]:
        raise ValueError("target_rows length")
    for name, c in (("filter_cols", filter_cols), ("target_cols", target_cols)):
        if c is not None:
            if not (sp.issparse(c) or isinstance(c, (list, np.ndarray))):
                raise TypeError(f"{name} must be a sparse matrix, list, numpy array, or None")
            if sp.issparse(c) and c.data.shape[0] != 0 and c.shape != (matrix1.shape[0], matrix2.shape[1]):
                raise ValueError(f"{name} shape")
    if not isinstance(verbose, bool):
        raise TypeError("verbose must be boolean")
    if format_output not in ("coo", "csr"):
        raise ValueError("format_output must be 'coo' or 'csr'")


def csr_sum(matrix, axis):
    """s_plus_utils.pyx:128-166: row sums by reduceat in fp32, column sums by bincount in fp64."""
    data = matrix.data.astype(np.float32, copy=False)
    indices = matrix.indices.astype(np.int32, copy=False)
    indptr = matrix.indptr.astype(np.int32, copy=False)
    if axis == 1:
        if data.shape[0] == 0:
            return np.zeros(matrix.shape[0], dtype=np.float32)
        starts = np.minimum(indptr[:-1], data.shape[0] - 1)
        res = np.add.reduceat(data, starts).astype(np.float32, copy=False)
        res[np.diff(indptr) == 0] = 0.0
        return res
    return np.bincount(indices, weights=data, minlength=matrix.shape[1]).astype(np.float32, copy=False)


def squared_norms(m1, m2):
    """s_plus_utils.pyx:169-201."""
    sq1 = sp.csr_array((np.square(m1.data, dtype=np.float32), m1.indices, m1.indptr), shape=m1.shape)
    sq2 = sp.csr_array((np.square(m2.data, dtype=np.float32), m2.indices, m2.indptr), shape=m2.shape)
    return csr_sum(sq1, 1), csr_sum(sq2, 0)


def depop(matrix, spec, p, axis):
    """s_plus_utils.pyx:231-278."""
    n = matrix.shape[0] if axis == 1 else matrix.shape[1]
    if isinstance(spec, (list, np.ndarray)):
        return np.power(spec, np.float32(p), dtype=np.float32)
    if spec == "none":
        return np.ones(n, dtype=np.float32)
    if spec == "sum":
        return np.power(csr_sum(matrix, axis), np.float32(p), dtype=np.float32)
    raise ValueError("Invalid weight spec")


def column_selector(cols):
    """s_plus_utils.pyx:311-361."""
    empty = np.array([], dtype=np.int32)
    if sp.issparse(cols) and cols.data.shape[0] != 0:
        c = cols.tocsr()
        c.eliminate_zeros()
        c.sort_indices()
        return MODE_MATRIX, np.array(c.indptr, dtype=np.int32), np.array(c.indices, dtype=np.int32)
    if isinstance(cols, (list, np.ndarray)) and len(cols) != 0:
        return MODE_ARRAY, empty, empty
    return MODE_NONE, empty, empty


def keep_columns(filter_cols, target_cols, n_cols):
    """s_plus_utils.pyx:364-421: columns that survive list-mode filter/target."""
    f_list = isinstance(filter_cols, (list, np.ndarray)) and len(filter_cols) != 0
    t_list = isinstance(target_cols, (list, np.ndarray)) and len(target_cols) != 0
    if t_list:
        mask = np.zeros(n_cols, dtype=bool)
        t = np.asarray(target_cols, dtype=np.int32)
        mask[t[(t >= 0) & (t < n_cols)]] = True
    else:
        mask = np.ones(n_cols, dtype=bool)
    if f_list:
        f = np.asarray(filter_cols, dtype=np.int32)
        mask[f[(f >= 0) & (f < n_cols)]] = False
    return mask


def filter_matrix_columns(matrix, mask):
    """s_plus_utils.pyx:424-490: drop B entries whose column is not kept; ids preserved."""
    keep = mask[matrix.indices]
    csum = np.concatenate([[0], np.cumsum(keep, dtype=np.int64)])
    indptr = csum[matrix.indptr].astype(np.int32)
    return (matrix.data.astype(np.float32, copy=False)[keep],
            matrix.indices[keep].astype(np.int32, copy=False), indptr)


def reorder_by_popularity(b_data, b_indices, b_indptr, n_cols, Yt, Yc, Yd,
                          f_mode, f_indptr, f_indices, t_mode, t_indptr, t_indices):
    """s_plus_utils.pyx:493-618: stable descending-popularity column permutation for the blocked path."""
    col_nnz = np.bincount(b_indices, minlength=n_cols)
    back = np.argsort(-col_nnz, kind="stable").astype(np.int32)
    if np.array_equal(back, np.arange(n_cols, dtype=np.int32)):
        return b_data, b_indices, b_indptr, Yt, Yc, Yd, f_indptr, f_indices, t_indptr, t_indices, None
    fwd = np.empty(n_cols, dtype=np.int32)
    fwd[back] = np.arange(n_cols, dtype=np.int32)
    tmp = sp.csr_array((b_data.copy(), fwd[b_indices].astype(np.int32), b_indptr.copy()),
                       shape=(len(b_indptr) - 1, n_cols))
    tmp.sort_indices()
    out = [np.asarray(tmp.data, np.float32), np.asarray(tmp.indices, np.int32), np.asarray(tmp.indptr, np.int32)]
    out += [None if y is None else y[back] for y in (Yt, Yc, Yd)]
    for mode, ip, ix in ((f_mode, f_indptr, f_indices), (t_mode, t_indptr, t_indices)):
        if mode == MODE_MATRIX and len(ix) > 0:
            m = sp.csr_array((np.ones(len(ix), np.float32), fwd[ix].astype(np.int32), ip.copy()),
                             shape=(len(ip) - 1, n_cols))
            m.sort_indices()
            out += [np.asarray(m.indptr, np.int32), np.asarray(m.indices, np.int32)]
        else:
            out += [ip, ix]
    return (*out, back)


# --------------------------------------------------------------------------- #
# the kernel call (s_plus.h:265-453 restated in spy_oracle.c)                  #
# --------------------------------------------------------------------------- #
def knn_kernel(targets, a, b, Xt, Yt, Xc, Yc, Xd, Yd, a1, l1, l2, l3, t1, t2, stab, bayes, thr, k, n_cols,
               f_mode, f_indptr, f_indices, t_mode, t_indptr, t_indices, num_threads=0, block_size=0):
    """a, b are (data f32, indices i32, indptr i32) triples.  Returns zero-padded slab (rows, cols, values)."""
    n_t = int(targets.shape[0])
    rows = np.zeros(n_t * k, dtype=np.int32)
    cols = np.zeros(n_t * k, dtype=np.int32)
    vals = np.zeros(n_t * k, dtype=np.float32)
    f32 = lambda x: ctypes.c_float(float(np.float32(x)))
    keep = [np.ascontiguousarray(x) for x in (targets, *a, *b, Xt, Yt, Xc, Yc, Xd, Yd,
                                              f_indptr, f_indices, t_indptr, t_indices)]
    (targets, ad, ai, ap, bd, bi, bp, Xt, Yt, Xc, Yc, Xd, Yd, f_indptr, f_indices, t_indptr, t_indices) = keep
    lib().spy_oracle_knn(
        ctypes.c_int32(n_t), _p(targets), _p(ad), _p(ai), _p(ap), _p(bd), _p(bi), _p(bp),
        _p(Xt), _p(Yt), _p(Xc), _p(Yc), _p(Xd), _p(Yd),
        f32(a1), f32(l1), f32(l2), f32(l3), f32(t1), f32(t2), f32(stab), f32(bayes), f32(thr),
        ctypes.c_int32(k), ctypes.c_int32(n_cols),
        ctypes.c_int32(f_mode), _p(f_indptr), _p(f_indices),
        ctypes.c_int32(t_mode), _p(t_indptr), _p(t_indices),
        _p(rows), _p(cols), _p(vals), ctypes.c_int32(num_threads), ctypes.c_int32(block_size))
    return rows, cols, vals


def s_plus_core(matrix1, matrix2=None, weight_depop_matrix1="none", weight_depop_matrix2="none",
                p1=0.0, p2=0.0, a1=1.0, l1=0.0, l2=0.0, l3=0.0, t1=1.0, t2=1.0, c1=0.5, c2=0.5, k=100,
                stabilized_shrink=0.0, bayesian_shrink=0.0, additive_shrink=0.0, threshold=0.0, binary=False,
                target_rows=None, filter_cols=None, target_cols=None, verbose=True, format_output="csr",
                num_threads=0, block_size=0, return_slab=False):
    """Restates the Cython driver s_plus.pyx:95-433 step by step."""
    if matrix2 is None:                                                        # s_plus.pyx:169-170
        matrix2 = matrix1.T
    validate(matrix1, matrix2, weight_depop_matrix1, weight_depop_matrix2, k, target_rows,
             filter_cols, target_cols, verbose, format_output)
    k = int(min(k, matrix2.shape[1]))                                           # s_plus.pyx:187-188
    if target_rows is None:                                                     # s_plus.pyx:191-196
        targets = np.arange(matrix1.shape[0], dtype=np.int32)
    else:
        targets = np.ascontiguousarray(np.asarray(target_rows, dtype=np.int32))
    matrix1 = matrix1.tocsr()                                                   # s_plus.pyx:205-211
    matrix2 = matrix2.tocsr()
    matrix1.eliminate_zeros()
    matrix2.eliminate_zeros()
    n_rows, n_cols = matrix1.shape[0], matrix2.shape[1]
    if block_size is None:                                                      # s_plus.pyx:218-225
        bs = 0
    elif block_size == 0:
        bs = DEFAULT_BLOCK_SIZE
    else:
        bs = int(block_size)
    use_blocking = bs > 0 and n_cols > bs

    orig1, orig2 = matrix1.data, matrix2.data                                   # s_plus.pyx:230-234
    if binary:                                                                  # s_plus_utils.pyx:281-308
        matrix1.data = np.ones(matrix1.data.shape[0], dtype=np.float32)
        matrix2.data = np.ones(matrix2.data.shape[0], dtype=np.float32)
    else:
        matrix1.data = matrix1.data.astype(np.float32, copy=False)
        matrix2.data = matrix2.data.astype(np.float32, copy=False)
    a = (matrix1.data, matrix1.indices.astype(np.int32, copy=False), matrix1.indptr.astype(np.int32, copy=False))
    b = (matrix2.data, matrix2.indices.astype(np.int32, copy=False), matrix2.indptr.astype(np.int32, copy=False))
    empty = np.array([], dtype=np.float32)
    Xt = Yt = Xc = Yc = Xd = Yd = empty
    l1, l2, l3, c1, c2 = (np.float32(x) for x in (l1, l2, l3, c1, c2))
    if l1 != 0 or l2 != 0:                                                      # s_plus.pyx:259-269
        sq1, sq2 = squared_norms(matrix1, matrix2)
    if l1 != 0:
        Xt, Yt = sq1, sq2
    if l2 != 0:                                                                 # s_plus_utils.pyx:204-228
        h = np.float32(additive_shrink)
        Xc = np.power(sq1 + h, c1, dtype=np.float32)
        Yc = np.power(sq2 + h, c2, dtype=np.float32)
    if l3 != 0:
        Xd = depop(matrix1, weight_depop_matrix1, p1, 1)
        Yd = depop(matrix2, weight_depop_matrix2, p2, 0)
    matrix1.data, matrix2.data = orig1, orig2                                   # s_plus.pyx:272

    f_mode, f_indptr, f_indices = column_selector(filter_cols)                  # s_plus.pyx:284-295
    t_mode, t_indptr, t_indices = column_selector(target_cols)
    if f_mode == MODE_ARRAY or t_mode == MODE_ARRAY:
        mask = keep_columns(filter_cols, target_cols, n_cols)
        # NB reference quirk: values are re-read from the restored, non-binary matrix2
        b = filter_matrix_columns(matrix2, mask)

    back = None
    if use_blocking:                                                            # s_plus.pyx:308-346
        (bd, bi, bp, Yt2, Yc2, Yd2, f_indptr, f_indices, t_indptr, t_indices, back) = reorder_by_popularity(
            np.asarray(b[0]), np.asarray(b[1]), np.asarray(b[2]), n_cols,
            Yt if l1 != 0 else None, Yc if l2 != 0 else None, Yd if l3 != 0 else None,
            f_mode, f_indptr, f_indices, t_mode, t_indptr, t_indices)
        b = (bd.astype(np.float32, copy=False), bi.astype(np.int32, copy=False), bp.astype(np.int32, copy=False))
        if l1 != 0:
            Yt = Yt2.astype(np.float32, copy=False)
        if l2 != 0:
            Yc = Yc2.astype(np.float32, copy=False)
        if l3 != 0:
            Yd = Yd2.astype(np.float32, copy=False)

    rows, cols, vals = knn_kernel(targets, a, b, Xt, Yt, Xc, Yc, Xd, Yd, a1, l1, l2, l3, t1, t2,
                                  stabilized_shrink, bayesian_shrink, threshold, k, n_cols,
                                  f_mode, f_indptr, f_indices, t_mode, t_indptr, t_indices,
                                  num_threads=num_threads, block_size=bs)
    if back is not None:                                                        # s_plus.pyx:387-392
        nz = (cols != 0) | (vals != 0)
        cols[nz] = back[cols[nz]]
    if return_slab:
        return rows, cols, vals, k
    if format_output == "coo":                                                  # utils.pyx:43-64
        return sp.coo_array((vals, (rows, cols)), shape=(n_rows, n_cols), dtype=np.float32)
    return slab_to_csr(rows, cols, vals, n_rows, n_cols)


def slab_to_csr(rows, cols, vals, n_rows, n_cols):
    """utils.pyx:67-173 + coo_to_csr.h:28-71 + eliminate_zeros (s_plus.pyx:424)."""
    nnz = int(vals.shape[0])
    idx_dtype = np.int32 if max(nnz, n_cols) <= np.iinfo(np.int32).max else np.int64
    Bp = np.zeros(n_rows + 1, dtype=np.int64)
    Bj = np.zeros(nnz, dtype=np.int64)
    Bx = np.zeros(nnz, dtype=np.float32)
    if nnz:
        lib().spy_oracle_coo_to_csr(ctypes.c_int32(n_rows), ctypes.c_int64(nnz), _p(rows), _p(cols), _p(vals),
                                    _p(Bp), _p(Bj), _p(Bx))
    res = sp.csr_array((Bx, Bj.astype(idx_dtype), Bp.astype(idx_dtype)), shape=(n_rows, n_cols), dtype=np.float32)
    res.eliminate_zeros()
    return res


# --------------------------------------------------------------------------- #
# similarity.py restatement (parameter presets only)                           #
# --------------------------------------------------------------------------- #
def shrink_values(shrink, shrink_type):
    """similarity.py:595-617."""
    if shrink_type == "stabilized":
        return shrink, 0.0, 0.0
    if shrink_type == "bayesian":
        return 0.0, shrink, 0.0
    if shrink_type == "additive":
        return 0.0, 0.0, shrink
    raise ValueError("shrink_type must be one of 'stabilized', 'bayesian', or 'additive'")


def _common(shrink, shrink_type, kw):
    s, b_, a_ = shrink_values(shrink, shrink_type)
    kw.setdefault("format_output", "coo")
    return dict(stabilized_shrink=s, bayesian_shrink=b_, additive_shrink=a_, **kw)


def similarity(name, matrix1, matrix2=None, *, core=None, normalize_fn=None, shrink=0.0, shrink_type="stabilized",
               alpha=None, beta=None, l1=0.5, l2=0.5, l3=0.0, t1=1.0, t2=1.0, c1=0.5, c2=0.5,
               pop1="none", pop2="none", beta1=0.0, beta2=0.0, **kw):
    """The nine public presets (similarity.py:9-592), parameterised over the core driver so the same
    table serves the oracle (core=s_plus_core) and the compiled reference (core=_ref s_plus)."""
    core = core or s_plus_core
    normalize_fn = normalize_fn or normalize
    base = _common(shrink, shrink_type, kw)
    if name == "dot_product":
        return core(matrix1, matrix2=matrix2, **base)
    if name == "cosine":
        return core(matrix1, matrix2=matrix2, l2=1, c1=0.5, c2=0.5, **base)
    if name == "asymmetric_cosine":
        a_ = 0.5 if alpha is None else alpha
        return core(matrix1, matrix2=matrix2, l2=1, c1=a_, c2=1 - a_, **base)
    if name == "tversky":
        return core(matrix1, matrix2=matrix2, l1=1, t1=1.0 if alpha is None else alpha,
                    t2=1.0 if beta is None else beta, **base)
    if name == "jaccard":
        return core(matrix1, matrix2=matrix2, l1=1, t1=1, t2=1, **base)
    if name == "dice":
        return core(matrix1, matrix2=matrix2, l1=1, t1=0.5, t2=0.5, **base)
    if name in ("p3alpha", "rp3beta"):                                          # similarity.py:410-415, 477-483
        a_ = 1.0 if alpha is None else alpha
        if matrix2 is None:
            matrix2 = matrix1.T
        extra = {}
        if name == "rp3beta":
            extra = dict(weight_depop_matrix2=np.asarray(matrix2.sum(axis=0)).ravel(),
                         p2=1.0 if beta is None else beta, l3=1)
        m1 = normalize_fn(matrix1, norm="l1", axis=1, inplace=False)
        m1.data = np.power(m1.data, a_)
        m2 = normalize_fn(matrix2, norm="l1", axis=1, inplace=False)
        m2.data = np.power(m2.data, a_)
        return core(matrix1=m1, matrix2=m2, **extra, **base)
    if name == "s_plus":
        return core(matrix1, matrix2=matrix2, l1=l1, l2=l2, l3=l3, t1=t1, t2=t2, c1=c1, c2=c2,
                    a1=1.0 if alpha is None else alpha, weight_depop_matrix1=pop1, weight_depop_matrix2=pop2,
                    p1=beta1, p2=beta2, **base)
    raise ValueError(name)


# --------------------------------------------------------------------------- #
# normalization.py restatement                                                 #
# --------------------------------------------------------------------------- #
def _prepare_csr(X, axis, inplace):
    """normalization.py:23-66."""
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


def _finalize_csr(X, axis):
    return (X.T if axis == 0 else X).tocsr()


def _suffix(X):
    ft = "f32" if X.data.dtype == np.float32 else "f64"
    it = {np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}[X.indptr.dtype]
    return f"{ft}_{it}", (ctypes.c_float if ft == "f32" else ctypes.c_double)


def normalize(X, norm="l2", axis=1, inplace=False):
    """normalization.py:91-113 -> normalization.pyx:97-197."""
    if norm not in ("l1", "l2", "max"):
        raise ValueError("norm must be one of ('l1', 'l2', 'max')")
    X = _prepare_csr(X, axis, inplace)
    suf, _ = _suffix(X)
    getattr(lib(), f"spy_oracle_{norm}_{suf}")(ctypes.c_int64(X.shape[0]), _p(X.data), _p(X.indptr))
    return _finalize_csr(X, axis)


def _check_modes(tf_mode, idf_mode):
    if tf_mode not in _TF:
        raise ValueError(f"tf_mode must be one of {tuple(_TF)}, got '{tf_mode}'")
    if idf_mode not in _IDF:
        raise ValueError(f"idf_mode must be one of {tuple(_IDF)}, got '{idf_mode}'")


def bm25plus(X, axis=1, k1=1.2, b=0.75, delta=1.0, logbase=_E, tf_mode="raw", idf_mode="bm25", inplace=False):
    """normalization.py:152-187 -> normalization.pyx:260-334."""
    _check_modes(tf_mode, idf_mode)
    X = _prepare_csr(X, axis, inplace)
    suf, ct = _suffix(X)
    getattr(lib(), f"spy_oracle_bm25plus_{suf}")(
        ctypes.c_int64(X.shape[0]), ctypes.c_int64(X.shape[1]), _p(X.data), _p(X.indices), _p(X.indptr),
        ct(k1), ct(b), ct(delta), ctypes.c_int(_TF[tf_mode]), ctypes.c_int(_IDF[idf_mode]), ct(logbase))
    return _finalize_csr(X, axis)


def bm25(X, axis=1, k1=1.2, b=0.75, logbase=_E, tf_mode="raw", idf_mode="bm25", inplace=False):
    """normalization.py:116-149 (bm25plus with delta=0)."""
    return bm25plus(X, axis=axis, k1=k1, b=b, delta=0.0, logbase=logbase, tf_mode=tf_mode,
                    idf_mode=idf_mode, inplace=inplace)


def tfidf(X, axis=1, logbase=_E, tf_mode="sqrt", idf_mode="smooth", inplace=False):
    """normalization.py:190-218 -> normalization.pyx:200-257."""
    _check_modes(tf_mode, idf_mode)
    X = _prepare_csr(X, axis, inplace)
    suf, ct = _suffix(X)
    getattr(lib(), f"spy_oracle_tfidf_{suf}")(
        ctypes.c_int64(X.shape[0]), ctypes.c_int64(X.shape[1]), _p(X.data), _p(X.indices), _p(X.indptr),
        ctypes.c_int(_TF[tf_mode]), ctypes.c_int(_IDF[idf_mode]), ct(logbase))
    return _finalize_csr(X, axis)


def max_threads() -> int:
    return int(lib().spy_oracle_max_threads())
