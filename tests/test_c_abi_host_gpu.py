"""The drop-in seam itself: ``spy_knn_topk_host`` called through ctypes with NumPy-owned HOST buffers -- exactly the
data the reference's Cython call site holds (similaripy/cython_code/s_plus.pyx:359-384) -- against the oracle's
restatement of ``compute_similarities_parallel`` (s_plus.h:265-453) on the same arrays.  The binding used here is the
one INTEGRATION.md shows for a maintainer of the reference."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle
from parity import assert_topk_parity, random_csr
from similaripy_b200 import _lib

pytestmark = pytest.mark.gpu

i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)


def _ptr(a):
    return None if a is None or a.size == 0 else a.ctypes.data_as(C.c_void_p)


def _host_call(targets, A, B, vecs, scal, k, filt=None, targ=None, threads=0, panel_width=0, with_rows=True, engine=0,
               devices=None):
    """Fill struct spy_knn_args with host pointers and call the library; returns the slab (rows, cols, vals, counts)."""
    lib = _lib.load()
    a = _lib.KnnArgs()
    keep = []

    def arr(x, dt):
        x = np.ascontiguousarray(x, dtype=dt)
        keep.append(x)
        return x
    targets = arr(targets, np.int32)
    a.n_targets, a.targets = len(targets), _ptr(targets)
    a.a_rows = A.shape[0]
    a.a_indptr, a.a_indices, a.a_data = _ptr(arr(A.indptr, np.int32)), _ptr(arr(A.indices, np.int32)), _ptr(arr(A.data, np.float32))
    a.b_rows, a.n_cols = B.shape[0], B.shape[1]
    a.b_indptr, a.b_indices, a.b_data = _ptr(arr(B.indptr, np.int32)), _ptr(arr(B.indices, np.int32)), _ptr(arr(B.data, np.float32))
    for name in ("Xtversky", "Ytversky", "Xcosine", "Ycosine", "Xdepop", "Ydepop"):
        v = vecs.get(name)
        setattr(a, name, None if v is None else _ptr(arr(v, np.float32)))
    for name, val in scal.items():
        setattr(a, name, float(np.float32(val)))
    a.k = k
    for mode_name, m in (("filter", filt), ("target", targ)):
        if m is not None:
            m = m.tocsr(); m.sort_indices()
            setattr(a, mode_name + "_mode", _lib.SEL_MATRIX)
            setattr(a, mode_name + "_indptr", _ptr(arr(m.indptr, np.int32)))
            setattr(a, mode_name + "_indices", _ptr(arr(m.indices, np.int32)))
    n = len(targets) * k
    rows, cols = np.full(n, -7, np.int32), np.full(n, -7, np.int32)
    vals, counts = np.full(n, np.nan, np.float32), np.full(len(targets), -7, np.int32)
    a.out_rows = _ptr(rows) if with_rows else None
    a.out_cols, a.out_values, a.out_counts = _ptr(cols), _ptr(vals), _ptr(counts)
    a.threads, a.panel_width, a.engine = threads, panel_width, engine
    if devices is None:
        _lib.check(lib.spy_knn_topk_host(C.byref(a), 0))
        return rows, cols, vals, counts
    dev = np.asarray(devices, dtype=np.int32)
    bounds = np.full(len(dev) + 1, -1, np.int32)
    _lib.check(lib.spy_knn_topk_multi_host(C.byref(a), _ptr(dev), len(dev), 1, _ptr(bounds)))
    return rows, cols, vals, counts, bounds


def _slab_to_csr(cols, vals, counts, targets, k, shape):
    r = np.repeat(np.asarray(targets), k)
    valid = (np.arange(k)[None, :] < counts[:, None]).ravel()
    return sp.csr_array((vals[valid], (r[valid], cols[valid])), shape=shape)


SCAL0 = dict(a1=1.0, l1=0.0, l2=0.0, l3=0.0, t1=1.0, t2=1.0, stabilized_shrink=0.0, bayesian_shrink=0.0, threshold=0.0)


def _oracle_slab(targets, A, B, vecs, scal, k, filt=None, targ=None):
    e = np.zeros(1, np.float32); ei = np.zeros(1, np.int32)
    g = lambda n: np.ascontiguousarray(vecs[n], np.float32) if vecs.get(n) is not None else e
    fm, fp, fi = (0, ei, ei)
    tm, tp, ti = (0, ei, ei)
    if filt is not None:
        f = filt.tocsr(); f.sort_indices(); fm, fp, fi = 2, f.indptr.astype(np.int32), f.indices.astype(np.int32)
    if targ is not None:
        t = targ.tocsr(); t.sort_indices(); tm, tp, ti = 2, t.indptr.astype(np.int32), t.indices.astype(np.int32)
    rows, cols, vals = oracle.knn_kernel(
        np.asarray(targets, np.int32), (A.data.astype(np.float32), A.indices.astype(np.int32), A.indptr.astype(np.int32)),
        (B.data.astype(np.float32), B.indices.astype(np.int32), B.indptr.astype(np.int32)),
        g("Xtversky"), g("Ytversky"), g("Xcosine"), g("Ycosine"), g("Xdepop"), g("Ydepop"),
        scal["a1"], scal["l1"], scal["l2"], scal["l3"], scal["t1"], scal["t2"], scal["stabilized_shrink"],
        scal["bayesian_shrink"], scal["threshold"], k, B.shape[1], fm, fp, fi, tm, tp, ti)
    return sp.csr_array((vals, (rows, cols)), shape=(A.shape[0], B.shape[1]))


@pytest.mark.parametrize("engine,panel_width", [(1, 0), (1, 256), (2, 0), (0, 0)], ids=["flat", "flat-panels", "stream", "auto"])
def test_host_entry_cosine_like(engine, panel_width):
    urm = random_csr(700, 900, 0.04, 3)
    A, B = urm.T.tocsr(), urm
    A.sort_indices(); B.sort_indices()
    sqa = np.asarray(A.multiply(A).sum(axis=1)).ravel().astype(np.float32)
    sqb = np.asarray(B.multiply(B).sum(axis=0)).ravel().astype(np.float32)
    vecs = dict(Xcosine=np.sqrt(sqa + 2.0).astype(np.float32), Ycosine=np.sqrt(sqb + 2.0).astype(np.float32))
    scal = dict(SCAL0, l2=1.0, stabilized_shrink=3.0)
    targets = np.arange(0, 900, 2, dtype=np.int32)
    k = 30
    rows, cols, vals, counts = _host_call(targets, A, B, vecs, scal, k, panel_width=panel_width, engine=engine)
    got = _slab_to_csr(cols, vals, counts, targets, k, (A.shape[0], B.shape[1]))
    ref = _oracle_slab(targets, A, B, vecs, scal, k)
    assert_topk_parity(ref, got, k=k, rtol=1e-5, what="host ABI cosine-like")
    # slab contract: rows[i*k+j] = target for j < count, tail zero-filled (s_plus.h:443-450, s_plus.pyx:351-353)
    for i in (0, 17, len(targets) - 1):
        c = counts[i]
        assert np.all(rows[i * k: i * k + c] == targets[i]) and np.all(rows[i * k + c: (i + 1) * k] == 0)
        assert np.all(cols[i * k + c: (i + 1) * k] == 0) and np.all(vals[i * k + c: (i + 1) * k] == 0)
        assert np.all(np.diff(vals[i * k: i * k + c]) <= 0)  # best first


def test_host_entry_dot_with_filter_and_target_matrices():
    urm = random_csr(300, 500, 0.05, 8)
    S = random_csr(500, 500, 0.06, 9)
    targets = np.array([5, 5, 299, 0, 123], dtype=np.int32)  # duplicates allowed at this level
    k = 12
    rows, cols, vals, counts = _host_call(targets, urm, S, {}, SCAL0, k, filt=urm, with_rows=False)
    got = _slab_to_csr(cols, vals, counts, targets, k, (300, 500))
    ref = _oracle_slab(targets, urm, S, {}, SCAL0, k, filt=urm)
    # duplicates: compare the de-duplicated rows (scipy sums duplicate coordinates)
    uniq = np.array([5, 299, 0, 123], dtype=np.int32)
    r2, c2, v2, n2 = _host_call(uniq, urm, S, {}, SCAL0, k, filt=urm)
    got_u = _slab_to_csr(c2, v2, n2, uniq, k, (300, 500))
    ref_u = _oracle_slab(uniq, urm, S, {}, SCAL0, k, filt=urm)
    assert_topk_parity(ref_u, got_u, k=k, rtol=1e-5, what="host ABI dot + filter matrix")
    assert counts[0] == counts[1] and np.array_equal(cols[:k], cols[k:2 * k])
    tm = random_csr(300, 500, 0.3, 10)
    r3, c3, v3, n3 = _host_call(uniq, urm, S, {}, SCAL0, k, targ=tm)
    assert_topk_parity(_oracle_slab(uniq, urm, S, {}, SCAL0, k, targ=tm), _slab_to_csr(c3, v3, n3, uniq, k, (300, 500)),
                       k=k, rtol=1e-5, what="host ABI dot + target matrix")
    del got, ref


def test_host_entry_error_convention():
    urm = random_csr(50, 60, 0.1, 1)
    with pytest.raises(_lib.SimilaripyB200Error):
        _host_call(np.arange(5, dtype=np.int32), urm.T.tocsr(), urm, {}, dict(SCAL0, l2=1.0), 5)  # l2 != 0 without Xcosine/Ycosine
    with pytest.raises(_lib.SimilaripyB200Error):
        _host_call(np.arange(5, dtype=np.int32), urm.T.tocsr(), urm, {}, SCAL0, 5, threads=333)
    msg = _lib.load().spy_last_error().decode()
    assert "threads" in msg


@pytest.mark.parametrize("engine,panel_width", [(1, 256), (2, 2048)], ids=["flat", "stream"])
def test_host_entry_sorts_unsorted_b_rows(engine, panel_width):
    """The reference accepts unsorted rows of B on its unblocked path (s_plus.h:418-438); the panel split needs them
    ascending, so the host entry sorts its device copy (the caller's arrays are untouched)."""
    rng = np.random.default_rng(5)
    urm = random_csr(400, 5000, 0.02, 11)
    A, B = urm.T.tocsr(), urm.copy()
    for r in range(B.shape[0]):  # shuffle every row of B in place
        s, e = B.indptr[r], B.indptr[r + 1]
        perm = rng.permutation(e - s)
        B.indices[s:e] = B.indices[s:e][perm]
        B.data[s:e] = B.data[s:e][perm]
    B.has_sorted_indices = False
    before = B.indices.copy()
    targets = np.arange(0, 5000, 7, dtype=np.int32)
    k = 20
    rows, cols, vals, counts = _host_call(targets, A, B, {}, SCAL0, k, panel_width=panel_width, engine=engine)
    assert np.array_equal(B.indices, before)
    got = _slab_to_csr(cols, vals, counts, targets, k, (A.shape[0], B.shape[1]))
    ref = _oracle_slab(targets, A, B, {}, SCAL0, k)
    assert_topk_parity(ref, got, k=k, rtol=1e-5, what="host ABI, unsorted B")


def test_multi_device_host_entry_matches_single_device():
    """spy_knn_topk_multi_host with a device list (the same GPU three times on a one-GPU box: three host threads, three
    ranges) assembles the same slab as the single-device call."""
    urm = random_csr(800, 700, 0.04, 21)
    A, B = urm.T.tocsr(), urm
    sqa = np.asarray(A.multiply(A).sum(axis=1)).ravel().astype(np.float32)
    sqb = np.asarray(B.multiply(B).sum(axis=0)).ravel().astype(np.float32)
    vecs = dict(Xcosine=np.sqrt(sqa).astype(np.float32), Ycosine=np.sqrt(sqb).astype(np.float32))
    scal = dict(SCAL0, l2=1.0)
    targets = np.arange(700, dtype=np.int32)
    k = 25
    r1, c1, v1, n1 = _host_call(targets, A, B, vecs, scal, k)
    n_dev = max(1, _lib.device_count())
    devices = [i % n_dev for i in range(3)]
    r3, c3, v3, n3, bounds = _host_call(targets, A, B, vecs, scal, k, devices=devices)
    assert bounds[0] == 0 and bounds[-1] == len(targets) and np.all(np.diff(bounds) > 0)
    work = np.diff(A.indptr)[targets] + 1
    shares = [work[bounds[i]: bounds[i + 1]].sum() / work.sum() for i in range(3)]
    assert max(shares) < 0.40  # cut by work, not by row count
    assert np.array_equal(n1, n3) and np.array_equal(r1, r3)
    ref = _oracle_slab(targets, A, B, vecs, scal, k)
    assert_topk_parity(ref, _slab_to_csr(c3, v3, n3, targets, k, (A.shape[0], B.shape[1])), k=k, rtol=1e-5, what="multi-device host ABI")
    with pytest.raises(_lib.SimilaripyB200Error):
        _host_call(targets, A, B, vecs, scal, k, devices=[0, 99])


@pytest.mark.parametrize("val,idx", [(np.float32, np.int32), (np.float64, np.int64), (np.float32, np.int64)])
def test_host_normalizer_entries(val, idx):
    """spy_normalize_rows_host / spy_tfidf_host / spy_bm25plus_host with the arrays of a scipy CSR matrix, as the reference's
    Cython functions receive them (normalization.pyx:97-102, 200-208, 260-271), against the oracle."""
    lib = _lib.load()
    m = random_csr(300, 200, 0.06, 31).astype(val)
    m.indices, m.indptr = m.indices.astype(idx), m.indptr.astype(idx)
    vd = _lib.F32 if val == np.float32 else _lib.F64
    xd = _lib.I32 if idx == np.int32 else _lib.I64
    rtol = 1e-5 if val == np.float32 else 1e-12
    for code, norm in enumerate(("l1", "l2", "max")):
        d = m.data.copy()
        _lib.check(lib.spy_normalize_rows_host(code, m.shape[0], _ptr(d), vd, _ptr(m.indptr), xd, 0))
        np.testing.assert_allclose(d, oracle.normalize(m, norm=norm).data, rtol=rtol)
    d = m.data.copy()
    _lib.check(lib.spy_tfidf_host(m.shape[0], m.shape[1], _ptr(d), vd, _ptr(m.indices), _ptr(m.indptr), xd,
                                  _lib.TF_MODES["sqrt"], _lib.IDF_MODES["smooth"], float(np.e), 0))
    np.testing.assert_allclose(d, oracle.tfidf(m).data, rtol=max(rtol, 1e-6))
    d = m.data.copy()
    _lib.check(lib.spy_bm25plus_host(m.shape[0], m.shape[1], _ptr(d), vd, _ptr(m.indices), _ptr(m.indptr), xd, 1.2, 0.75, 1.0,
                                     _lib.TF_MODES["raw"], _lib.IDF_MODES["bm25"], float(np.e), 0))
    np.testing.assert_allclose(d, oracle.bm25plus(m).data, rtol=max(rtol, 1e-5))
    assert lib.spy_normalize_rows_host(9, m.shape[0], _ptr(d), vd, _ptr(m.indptr), xd, 0) < 0
