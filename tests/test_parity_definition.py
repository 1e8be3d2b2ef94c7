"""The parity definition itself (tests/parity.py) on hand-made rows: what it must accept
(ties at the k-th value resolved differently, an entry on the edge of the tie band) and what it must reject."""
import numpy as np
import pytest
import scipy.sparse as sp

from parity import assert_topk_parity


def _row(cols, vals, n_cols=64):
    cols, vals = np.asarray(cols, np.int32), np.asarray(vals, np.float32)
    return sp.csr_array((vals, cols, np.array([0, len(cols)], np.int32)), shape=(1, n_cols))


K = 4


def test_identical_rows_pass():
    a = _row([1, 2, 3, 4], [4.0, 3.0, 2.0, 1.0])
    info = assert_topk_parity(a, a, K)
    assert info["full_rows"] == 1


def test_order_inside_a_row_is_irrelevant():
    assert_topk_parity(_row([1, 2, 3, 4], [4.0, 3.0, 2.0, 1.0]), _row([4, 3, 2, 1], [1.0, 2.0, 3.0, 4.0]), K)


def test_exact_tie_at_the_boundary_may_pick_either_column():
    # columns 4 and 9 tie for the last place: the reference's choice depends on its traversal order (s_plus.h:45-59)
    assert_topk_parity(_row([1, 2, 3, 4], [4.0, 3.0, 2.0, 1.0]), _row([1, 2, 3, 9], [4.0, 3.0, 2.0, 1.0]), K, rtol=0)


def test_entry_on_the_edge_of_the_tie_band_passes():
    # column 3 is within 2 * rtol of the k-th value on one side and just outside rtol on the other: one ulp apart
    lo = np.float32(1.0)
    edge_hi = np.float32(lo * (1 + 1.05e-5))
    edge_lo = np.float32(lo * (1 + 0.95e-5))
    assert_topk_parity(_row([1, 2, 3, 4], [4.0, 3.0, edge_hi, lo]), _row([1, 2, 3, 4], [4.0, 3.0, edge_lo, lo]), K)
    assert_topk_parity(_row([1, 2, 3, 4], [4.0, 3.0, edge_hi, lo]), _row([1, 2, 9, 4], [4.0, 3.0, edge_lo, lo]), K)


def test_missing_column_above_the_band_fails():
    with pytest.raises(AssertionError, match="outside the boundary band"):
        assert_topk_parity(_row([1, 2, 3, 4], [4.0, 3.0, 2.0, 1.0]), _row([1, 2, 7, 4], [4.0, 3.0, 2.0, 1.0]), K)


def test_value_mismatch_fails():
    with pytest.raises(AssertionError, match="sorted values differ"):
        assert_topk_parity(_row([1, 2, 3, 4], [4.0, 3.0, 2.0, 1.0]), _row([1, 2, 3, 4], [4.0, 3.0, 2.001, 1.0]), K)


def test_integer_data_is_bit_exact():
    with pytest.raises(AssertionError, match="sorted values differ"):
        assert_topk_parity(_row([1, 2, 3, 4], [4.0, 3.0, 2.0, 1.0]), _row([1, 2, 3, 4], [4.0, 3.0, np.float32(2.0000002), 1.0]), K, rtol=0)


def test_short_row_needs_identical_columns():
    # fewer than k entries: nothing was truncated, so there is no boundary to excuse a different column
    with pytest.raises(AssertionError, match="column ids differ"):
        assert_topk_parity(_row([1, 2, 3], [3.0, 2.0, 1.0]), _row([1, 2, 9], [3.0, 2.0, 1.0]), K)


def test_row_length_mismatch_fails():
    with pytest.raises(AssertionError, match="entries"):
        assert_topk_parity(_row([1, 2, 3, 4], [4.0, 3.0, 2.0, 1.0]), _row([1, 2, 3], [4.0, 3.0, 2.0]), K)
