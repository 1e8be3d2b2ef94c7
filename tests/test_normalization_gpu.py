"""GPU: in-place CSR normalizers (csrc/normalize.cu through normalization.py) against the oracle on seeded
inputs: every tf/idf mode, float32/float64 x int32/int64, axis 0/1, inplace semantics, empty rows, and the
device-resident (DeviceMatrix) variants.  Tolerance: 1e-5 relative for float32 (north_star), 1e-12 float64."""
import numpy as np
import pytest
import scipy.sparse as sp

import similaripy_b200 as sim
from oracle import oracle
from parity import random_csr

pytestmark = pytest.mark.gpu


def _tol(dtype):
    return 1e-5 if dtype == np.float32 else 1e-12


def _with_empty_rows(m):
    m = m.tolil()
    m[3, :] = 0
    m[17, :] = 0
    m = m.tocsr()
    m.eliminate_zeros()
    return m


@pytest.mark.parametrize("norm", ["l1", "l2", "max"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("itype", [np.int32, np.int64])
@pytest.mark.parametrize("axis", [0, 1])
def test_normalize(norm, dtype, itype, axis):
    m = _with_empty_rows(random_csr(500, 300, 0.05, seed=1, dtype=dtype))
    m.data -= dtype(0.3)  # mixed signs: l1 uses |x|, max skips rows whose max <= 0
    m = sp.csr_array((m.data, m.indices.astype(itype), m.indptr.astype(itype)), shape=m.shape)
    got = sim.normalize(m.copy(), norm=norm, axis=axis)
    ref = oracle.normalize(m.copy(), norm=norm, axis=axis)
    np.testing.assert_array_equal(got.indices, ref.indices)
    np.testing.assert_allclose(got.data, ref.data, rtol=_tol(dtype))


@pytest.mark.parametrize("tf_mode", ["binary", "raw", "sqrt", "freq", "log"])
@pytest.mark.parametrize("idf_mode", ["unary", "base", "smooth", "prob", "bm25"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_tfidf_and_bm25_modes(tf_mode, idf_mode, dtype):
    m = _with_empty_rows(random_csr(400, 120, 0.08, seed=2, dtype=dtype))
    m.data *= dtype(3.0)
    for fn, kw in (("tfidf", {}), ("bm25", {}), ("bm25plus", dict(delta=0.8, k1=1.7, b=0.6, logbase=2.0))):
        got = getattr(sim, fn)(m.copy(), tf_mode=tf_mode, idf_mode=idf_mode, **kw)
        ref = getattr(oracle, fn)(m.copy(), tf_mode=tf_mode, idf_mode=idf_mode, **kw)
        np.testing.assert_allclose(got.data, ref.data, rtol=_tol(dtype), atol=1e-30, err_msg=f"{fn} {tf_mode} {idf_mode}")


def test_inplace_semantics():
    m = random_csr(200, 100, 0.05, seed=3)
    keep = m.data.copy()
    out = sim.normalize(m, norm="l2", inplace=False)
    np.testing.assert_array_equal(m.data, keep)          # untouched
    assert out is not m
    data_ref = m.data
    out2 = sim.bm25(m, inplace=True)
    assert np.shares_memory(out2.data, data_ref)          # same buffer mutated (normalization.py:62-66)
    np.testing.assert_allclose(out2.data, oracle.bm25(sp.csr_array((keep, m.indices, m.indptr), shape=m.shape)).data, rtol=1e-5)


def test_integer_input_becomes_float32():
    m = random_csr(100, 60, 0.1, seed=4, integer=True).astype(np.int64)
    got = sim.tfidf(m)
    assert got.dtype == np.float32 and got.format == "csr"
    np.testing.assert_allclose(got.data, oracle.tfidf(m).data, rtol=1e-5)


def test_bm25_large_sequential_mean():
    """avg_doc_len is a sequential float32 running sum in the reference (normalization.pyx:297-323)."""
    m = random_csr(200_000, 500, 0.01, seed=5)
    got = sim.bm25(m.copy())
    ref = oracle.bm25(m.copy())
    np.testing.assert_allclose(got.data, ref.data, rtol=1e-5)


def test_device_matrix_normalizers():
    m = random_csr(600, 250, 0.04, seed=6)
    d = sim.to_device(m)
    for fn, kw in (("bm25", {}), ("tfidf", {}), ("normalize", dict(norm="l1")), ("bm25", dict(axis=0))):
        got = sim.to_host(getattr(sim, fn)(d, **kw)).tocsr()
        ref = getattr(oracle, fn)(m.copy(), **kw)
        got.sort_indices(); ref.sort_indices()
        np.testing.assert_array_equal(got.indices, ref.indices)
        np.testing.assert_allclose(got.data, ref.data, rtol=1e-5)
    np.testing.assert_array_equal(sim.to_host(d).data, m.data)  # inplace=False left the handle untouched
