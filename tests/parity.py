"""Tie-aware parity definition for top-k similarity results.

Two results agree when, for every row,
  (i)   the sorted value lists have the same length and agree within ``rtol`` (bit-exact with rtol=0);
  (ii)  for rows that are "full" (k entries), every entry whose value is clearly above that row's smallest kept
        value (by more than 2 * rtol) on one side is present on the other side; a row with fewer than k
        entries was not truncated, so its id set must match completely.
The reference itself is order dependent on exact ties at the k boundary (its blocked and unblocked
modes disagree with each other on binary data), which is why boundary ties are excluded from (ii).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _rows(m):
    m = m.tocsr() if m.format != "csr" else m
    m = m.copy()
    m.eliminate_zeros()
    return m


def assert_topk_parity(ref, got, k, rtol=1e-5, atol=0.0, what=""):
    ref, got = _rows(ref), _rows(got)
    assert ref.shape == got.shape, f"{what}: shape {got.shape} != {ref.shape}"
    n_full = n_band = 0
    for r in range(ref.shape[0]):
        rs, re = ref.indptr[r], ref.indptr[r + 1]
        gs, ge = got.indptr[r], got.indptr[r + 1]
        rv, gv = ref.data[rs:re], got.data[gs:ge]
        ri, gi = ref.indices[rs:re], got.indices[gs:ge]
        assert rv.shape[0] == gv.shape[0], f"{what}: row {r} has {gv.shape[0]} entries, reference {rv.shape[0]}"
        if rv.shape[0] == 0:
            continue
        np.testing.assert_allclose(np.sort(gv), np.sort(rv), rtol=rtol, atol=atol,
                                   err_msg=f"{what}: row {r} sorted values differ")
        if rv.shape[0] >= k:  # truncated row: entries tied with the k-th value (within rtol) may legitimately differ
            n_full += 1
            lo = min(rv.min(), gv.min())
            band = abs(lo) * max(rtol, 1e-7) + atol
            n_band += int((rv <= lo + band).sum())
            # every entry CLEARLY above the band on one side must be present on the other side (anywhere: an entry
            # sitting on the band's edge can fall on different sides of it in the two results by one ulp)
            r_strict, g_strict = set(ri[rv > lo + 2 * band].tolist()), set(gi[gv > lo + 2 * band].tolist())
            r_all, g_all = set(ri.tolist()), set(gi.tolist())
            assert r_strict <= g_all and g_strict <= r_all, (
                f"{what}: row {r} column ids differ outside the boundary band: "
                f"only-ref={sorted(r_strict - g_all)[:8]} only-got={sorted(g_strict - r_all)[:8]}")
        else:
            rset, gset = set(ri.tolist()), set(gi.tolist())
            assert rset == gset, (f"{what}: row {r} column ids differ: "
                                  f"only-ref={sorted(rset - gset)[:8]} only-got={sorted(gset - rset)[:8]}")
    return dict(full_rows=n_full, band_entries=n_band)


def assert_same_matrix(ref, got, rtol=1e-6, atol=0.0, what=""):
    """Entry-wise equality of two sparse matrices (same structure after sorting indices)."""
    ref, got = ref.tocsr().copy(), got.tocsr().copy()
    ref.sort_indices(); got.sort_indices()
    assert ref.shape == got.shape, what
    np.testing.assert_array_equal(got.indptr, ref.indptr, err_msg=f"{what}: indptr")
    np.testing.assert_array_equal(got.indices, ref.indices, err_msg=f"{what}: indices")
    np.testing.assert_allclose(got.data, ref.data, rtol=rtol, atol=atol, err_msg=f"{what}: data")


def check_sum(x):
    """The reference tests' tie-invariant digest: sum over rows of (row sum)^2 (tests/test_similarity.py:8-14)."""
    aux = np.asarray(x.sum(axis=1)).ravel().astype(np.float64)
    return float(np.sum(aux * aux))


def random_csr(n_rows, n_cols, density, seed, dtype=np.float32, integer=False):
    rng = np.random.default_rng(seed)
    m = sp.random_array((n_rows, n_cols), density=density, format="csr", dtype=np.float32, random_state=rng)
    if integer:
        m.data = np.floor(m.data * 5).astype(np.float32) + 1.0
    return m.astype(dtype)
