"""CPU: the oracle (oracle/oracle.py + spy_oracle.c) against the golden vectors generated from the unmodified
reference (tests/golden/make_golden.py).  This is what pins the oracle: the raw slab must match the
reference's entry by entry -- same columns IN THE SAME HEAP ORDER (which pins the tie rule and the
first-touch candidate order), values within 2e-6 relative (the reference is built with -ffast-math, the
oracle is not), bit-exact on integer-valued data."""
import numpy as np
import pytest

import golden_io
from oracle import oracle
from parity import assert_topk_parity

EXACT = {"dot_int", "jaccard_binary"} | {f"jaccard_binary_block_{b}" for b in (None, 32, 64)}


def _slab(case):
    kw = dict(case["kw"])
    kw.pop("format_output", None)
    res = oracle.similarity(case["fn"], case["m1"].copy(), None if case["m2"] is None else case["m2"].copy(),
                            format_output="coo", **kw)
    return res.row, res.col, res.data


@pytest.mark.parametrize("name", golden_io.similarity_cases())
def test_similarity_slab_matches_reference(name):
    case = golden_io.load_similarity(name)
    rows, cols, vals = _slab(case)
    grows, gcols, gvals = case["slab"]
    assert rows.shape == grows.shape
    if name in EXACT or name.startswith("dot_int"):
        np.testing.assert_array_equal(vals, gvals)
        np.testing.assert_array_equal(cols, gcols)
        np.testing.assert_array_equal(rows, grows)
        return
    # float data: the same entries; a position may differ only where two values are within rounding of each other
    np.testing.assert_allclose(np.sort(vals), np.sort(gvals), rtol=2e-6, atol=1e-9)
    same = cols == gcols
    if not same.all():
        k = case["k"]
        for i in np.unique(np.nonzero(~same)[0] // k):
            a = dict(zip(cols[i * k:(i + 1) * k].tolist(), vals[i * k:(i + 1) * k].tolist()))
            b = dict(zip(gcols[i * k:(i + 1) * k].tolist(), gvals[i * k:(i + 1) * k].tolist()))
            common = set(a) & set(b)
            for c in common:
                assert abs(a[c] - b[c]) <= 2e-6 * abs(b[c]) + 1e-9
            lo = min(min(a.values()), min(b.values()))
            for c in (set(a) ^ set(b)):  # only near-ties at the k boundary may swap
                v = a.get(c, b.get(c))
                assert abs(v - lo) <= 4e-6 * abs(lo) + 1e-9, f"{name}: row {i} column {c} differs away from the boundary"
    import scipy.sparse as sp
    got = sp.coo_array((vals, (rows, cols)), shape=case["shape"]).tocsr()
    assert_topk_parity(case["ref_csr"], got, k=case["k"], rtol=2e-6, atol=1e-9, what=name)


@pytest.mark.parametrize("name", golden_io.normalization_cases())
def test_normalization_matches_reference(name):
    case = golden_io.load_normalization(name)
    got = getattr(oracle, case["fn"])(case["m"].copy(), **case["kw"])
    ref = case["out"]
    np.testing.assert_array_equal(got.indptr, ref.indptr)
    np.testing.assert_array_equal(got.indices, ref.indices)
    rtol = 2e-6 if ref.data.dtype == np.float32 else 1e-12
    np.testing.assert_allclose(got.data, ref.data, rtol=rtol, atol=1e-30)
    assert got.data.dtype == ref.data.dtype
