"""Host logic of the block path for matrices beyond int32 stored entries (similaripy_b200/_engine.py: _greedy_cuts,
wide_blocks): the cuts respect the limit, and the clamp / stack arithmetic the CUDA helpers implement
(include/similaripy_b200.h: spy_csr_wide_block_indptr_dev, spy_csr_indptr_add_dev, spy_slab_merge_dev) describes the
matrices scipy builds for the same blocks."""
import numpy as np
import pytest
import scipy.sparse as sp

from similaripy_b200 import _engine


def test_greedy_cuts_respect_the_limit():
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 50, size=1000)
    prefix = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for limit in (49, 50, 333, 10_000, 10**9):
        cuts = _engine._greedy_cuts(prefix, limit, "row")
        assert cuts[0] == 0 and cuts[-1] == 1000 and all(b > a for a, b in zip(cuts, cuts[1:]))
        sizes = [prefix[b] - prefix[a] for a, b in zip(cuts, cuts[1:])]
        assert max(sizes) <= limit
        # greedy: a block could not have taken the next row as well
        for (a, b), s in zip(zip(cuts, cuts[1:]), sizes):
            assert b == 1000 or s + lens[b] > limit
    with pytest.raises(ValueError, match="more than 40 stored entries"):
        _engine._greedy_cuts(prefix, 40, "row")
    assert _engine._greedy_cuts(np.zeros(1, np.int64), 10, "row") == [0]


def test_row_block_is_a_clamped_view_and_column_blocks_stack():
    m = sp.random_array((60, 45), density=0.2, format="csr", dtype=np.float32, random_state=np.random.default_rng(1))
    ip = m.indptr.astype(np.int64)
    cuts = _engine._greedy_cuts(ip, m.nnz // 3 + 5, "row")
    assert len(cuts) >= 3
    total = sp.csr_array(m.shape, dtype=np.float32)
    pieces = []
    for r0, r1 in zip(cuts, cuts[1:]):
        lo, hi = ip[r0], ip[r1]
        blk_indptr = (np.clip(ip, lo, hi) - lo).astype(np.int32)  # spy_csr_wide_block_indptr_dev
        blk = sp.csr_array((m.data[lo:hi], m.indices[lo:hi], blk_indptr), shape=m.shape)
        want = m.copy().tolil(); want[:r0] = 0; want[r1:] = 0
        assert (blk != want.tocsr()).nnz == 0
        total = total + blk
        keep = (blk.indices >= 10) & (blk.indices < 30)  # a column block of this row block
        cnt = np.add.reduceat(np.concatenate([keep, [False]]).astype(np.int64), np.minimum(blk_indptr[:-1], keep.shape[0]))
        cnt[blk_indptr[:-1] == blk_indptr[1:]] = 0
        pieces.append((np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32), blk.indices[keep], blk.data[keep]))
    assert (total != m).nnz == 0
    stacked = sp.csr_array((np.concatenate([p[2] for p in pieces]), np.concatenate([p[1] for p in pieces]),
                            np.sum([p[0] for p in pieces], axis=0)), shape=m.shape)  # spy_csr_indptr_add_dev
    want = m.tocsc()[:, 10:30]
    assert (stacked[:, 10:30] != want).nnz == 0 and stacked[:, :10].nnz == 0 and stacked[:, 30:].nnz == 0
