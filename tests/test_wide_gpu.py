"""Matrices beyond int32 stored entries: 64-bit indptr handles computed in int32-indexed blocks.

The reference narrows indptr / indices to int32 (s_plus.pyx:241-244) and cannot run such inputs at all, so the oracle is
asked for the same call on the SAME matrix at a size it can hold: the block path is forced on small matrices by lowering
``_engine.WIDE_NNZ_LIMIT`` (every combination of row blocks of matrix1 x column blocks of matrix2, both orientations of
both operands, selectors, depop weights, duplicate target rows, COO padding), and one test runs a real matrix with
2.2e9 stored entries, checked against a dense restatement of the selected rows in torch."""
import numpy as np
import pytest
import scipy.sparse as sp

import similaripy_b200 as sim
from similaripy_b200 import _engine, _lib
from oracle import oracle
from parity import assert_topk_parity, random_csr

pytestmark = pytest.mark.gpu


@pytest.fixture
def small_limit(monkeypatch):
    def set_limit(n):
        monkeypatch.setattr(_engine, "WIDE_NNZ_LIMIT", int(n))
    return set_limit


def _both(name, m, m2=None, k=30, **kw):
    got = getattr(sim, name)(m, m2, k=k, verbose=False, **kw)
    ref = oracle.similarity(name, m, m2, k=k, verbose=False, **kw)
    return ref, got


@pytest.mark.parametrize("name,kw", [
    ("dot_product", {}), ("cosine", {}), ("jaccard", {}), ("tversky", dict(alpha=0.8, beta=0.4)),
    ("asymmetric_cosine", dict(alpha=0.2)), ("rp3beta", dict(alpha=0.8, beta=0.4)),
    ("s_plus", dict(l1=0.5, l2=0.5, l3=1, t1=1, t2=1, c1=0.5, c2=0.5, alpha=1, beta1=0, beta2=0, pop1="none", pop2="sum")),
])
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_blocks_item_item(small_limit, name, kw, fmt):
    m = random_csr(700, 500, 0.03, seed=21)  # 10 500 entries -> 4 row blocks x 4 column blocks
    small_limit(m.nnz // 4 + 200)
    m = m.asformat(fmt)
    assert _engine._is_wide(m)
    ref, got = _both(name, m, k=25, format_output="csr", **kw)
    assert_topk_parity(ref, got, k=25, rtol=1e-5, what=f"wide {name} {fmt}")


@pytest.mark.parametrize("fmt1", ["csr", "csc"])
@pytest.mark.parametrize("fmt2", ["csr", "csc"])
def test_blocks_two_matrices_selectors_and_weights(small_limit, fmt1, fmt2):
    rng = np.random.default_rng(5)
    m1 = random_csr(400, 300, 0.05, seed=22).asformat(fmt1)
    m2 = random_csr(300, 600, 0.04, seed=23).asformat(fmt2)
    small_limit(2500)  # m1: 6 000 entries, m2: 7 200 entries
    rows = [3, 399, 17, 200, 17, 250, 0]  # duplicates, both ends
    w1 = rng.random(400).astype(np.float32) + 0.5
    w2 = rng.random(600) + 0.5  # float64 weights
    fm = random_csr(400, 600, 0.01, seed=24)  # (a selector itself has to fit int32 indexing)
    big = random_csr(400, 600, 0.05, seed=24)
    with pytest.raises(NotImplementedError):
        sim.dot_product(m1, m2, k=20, filter_cols=big, verbose=False)
    for kw in (dict(), dict(target_rows=rows), dict(filter_cols=[1, 5, 599, 10_000], target_cols=list(range(0, 600, 2))),
               dict(filter_cols=fm, target_rows=rows), dict(target_cols=fm), dict(threshold=0.05, binary=True)):
        ref = oracle.similarity("s_plus", m1, m2, k=20, l1=0.3, l2=0.4, l3=0.5, t1=0.7, t2=0.2, c1=0.4, c2=0.6, pop1=w1, pop2=w2,
                                beta1=0.3, beta2=0.6, shrink=2.0, format_output="csr", verbose=False, **kw)
        got = sim.s_plus(m1, m2, k=20, l1=0.3, l2=0.4, l3=0.5, t1=0.7, t2=0.2, c1=0.4, c2=0.6, pop1=w1, pop2=w2,
                         beta1=0.3, beta2=0.6, shrink=2.0, format_output="csr", verbose=False, **kw)
        assert_topk_parity(ref, got, k=20, rtol=1e-5, what=f"wide s_plus {fmt1}/{fmt2} {sorted(kw)}")


def test_blocks_integer_bit_exact_and_coo_padding(small_limit):
    m = random_csr(500, 260, 0.06, seed=25, integer=True)
    small_limit(3000)
    ref, got = _both("dot_product", m, k=600, format_output="coo")  # k > n_cols = 500: clipped; COO keeps the padding triples
    assert got.format == "coo" and got.data.shape[0] == ref.data.shape[0] == 500 * 500
    assert_topk_parity(ref, got, k=500, rtol=0.0, what="wide integer dot_product coo")
    ref, got = _both("dot_product", m, k=40, format_output="csr")
    assert_topk_parity(ref, got, k=40, rtol=0.0, what="wide integer dot_product")
    few = sim.dot_product(m, k=7, target_rows=[499, 2], format_output="coo", verbose=False)
    assert few.data.shape[0] == 14 and set(few.row[few.data != 0].tolist()) <= {2, 499}


def test_blocks_empty_and_missing_rows(small_limit):
    m = random_csr(300, 200, 0.05, seed=26).tolil()
    m[10:40] = 0  # a run of empty rows across a block border
    m = m.tocsr()
    small_limit(800)
    ref, got = _both("cosine", m, k=10, format_output="csr")
    assert_topk_parity(ref, got, k=10, rtol=1e-5, what="wide cosine with empty rows")
    none = sim.cosine(m, k=10, target_rows=[], format_output="csr", verbose=False)
    assert none.nnz == 0 and none.shape == (300, 300)


def test_wide_device_matrix_handles(small_limit):
    m = random_csr(600, 350, 0.04, seed=27)
    small_limit(2000)
    d = sim.to_device(m)
    assert isinstance(d.stored, _engine.WideCSR) and d.stored.indptr.dtype == _engine._torch().int64
    back = _engine.to_host(d)
    assert (back != m).nnz == 0
    np.testing.assert_allclose(_engine.axis_sum(d, 1).cpu().numpy(), np.asarray(m.sum(axis=1)).ravel(), rtol=1e-5)
    np.testing.assert_allclose(_engine.axis_sum(d, 0).cpu().numpy(), np.asarray(m.sum(axis=0)).ravel(), rtol=1e-5)
    np.testing.assert_allclose(_engine.axis_sum(d.T, 1).cpu().numpy(), np.asarray(m.sum(axis=0)).ravel(), rtol=1e-5)
    n = sim.normalize(d, norm="l2", axis=1)
    np.testing.assert_allclose(_engine.to_host(n).data, oracle.normalize(m, norm="l2", axis=1).data, rtol=1e-5)
    with pytest.raises(NotImplementedError):
        sim.normalize(d, norm="l2", axis=0)
    with pytest.raises(NotImplementedError):
        sim.bm25(d)
    ref = oracle.similarity("cosine", m.T.tocsr(), k=15, format_output="csr", verbose=False)
    got = sim.cosine(d.T, k=15, format_output="csr", verbose=False)           # item-item on a handle
    again = sim.cosine(d.T, k=15, format_output="csr", verbose=False)         # blocks and tables come from the handle's cache
    assert_topk_parity(ref, got, k=15, rtol=1e-5, what="wide handle item-item")
    assert_topk_parity(got, again, k=15, rtol=1e-6, what="wide handle, second call")  # (float sums differ by ulps between runs)
    on_dev = sim.cosine(d.T, k=15, verbose=False, on_device=True)
    assert_topk_parity(ref, _engine.to_host(on_dev), k=15, rtol=1e-5, what="wide handle on_device")
    with pytest.raises(NotImplementedError):
        sim.cosine(d.T, k=15, verbose=False, tuning=dict(tie_mode="reference"))
    with pytest.raises(NotImplementedError):
        with sim.sharded.shard_rows(gather=False, rank=0, world=2):
            sim.cosine(d.T, k=15, verbose=False)


def test_row_too_long_for_a_block(small_limit):
    m = random_csr(50, 400, 0.5, seed=28)
    small_limit(100)
    with pytest.raises(ValueError, match="stored entries"):
        sim.dot_product(m, k=5, verbose=False)


def test_slab_merge_kernel_against_numpy():
    torch = _engine._torch()
    ctx = _engine.Ctx(None)
    rng = np.random.default_rng(9)
    n, k = 257, 37
    def slab(col0):
        counts = rng.integers(0, k + 1, size=n).astype(np.int32)
        counts[:3] = (0, k, 1)
        cols = np.zeros((n, k), np.int32); vals = np.zeros((n, k), np.float32)
        for i in range(n):
            c = counts[i]
            cc = col0 + rng.choice(1000, size=c, replace=False)
            vv = np.round(rng.normal(size=c) * 3).astype(np.float32) / 2 + 0.0  # many ties, negatives, zeros (no -0.0)
            order = np.lexsort((cc, -vv))
            cols[i, :c], vals[i, :c] = cc[order], vv[order]
        return cols, vals, counts
    a, b = slab(0), slab(5000)
    dev = [ctx.h2d(x.ravel()) for x in (*a, *b)]
    out = (ctx.empty(n * k, torch.int32), ctx.empty(n * k, torch.float32), ctx.empty(n, torch.int32))
    _lib.check(ctx.lib.spy_slab_merge_dev(n, k, *[t.data_ptr() for t in dev], *[t.data_ptr() for t in out], ctx.sptr))
    oc, ov, on = out[0].cpu().numpy().reshape(n, k), out[1].cpu().numpy().reshape(n, k), out[2].cpu().numpy()
    for i in range(n):
        cc = np.concatenate([a[0][i, :a[2][i]], b[0][i, :b[2][i]]]); vv = np.concatenate([a[1][i, :a[2][i]], b[1][i, :b[2][i]]])
        order = np.lexsort((cc, -vv))[:k]
        assert on[i] == order.shape[0]
        np.testing.assert_array_equal(oc[i, :on[i]], cc[order]); np.testing.assert_array_equal(ov[i, :on[i]], vv[order])
        assert not oc[i, on[i]:].any() and not ov[i, on[i]:].any()


def test_real_size_beyond_int32_entries():
    """2.2e9 stored entries (2 200 000 x 100 000, 1000 per row, integer values): user-user dot product and cosine of 48
    rows spread over both row blocks, against torch on the very same device arrays."""
    torch = _engine._torch()
    if torch.cuda.mem_get_info()[1] < 120 * 2**30:
        pytest.skip("needs a 180 GB B200")
    dev = torch.device("cuda", torch.cuda.current_device())
    n_rows, n_cols, per_row = 2_200_000, 100_000, 1000
    step = n_cols // per_row
    nnz = n_rows * per_row
    assert nnz > np.iinfo(np.int32).max
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    data = torch.empty(nnz, dtype=torch.float32, device=dev)
    j = torch.arange(per_row, device=dev, dtype=torch.int64)[None, :]
    for r0 in range(0, n_rows, 100_000):  # row r: columns off(r) + step * j, ascending; values 1..7 from a hash
        r = torch.arange(r0, min(r0 + 100_000, n_rows), device=dev, dtype=torch.int64)[:, None]
        off = (r * 2654435761 >> 7) % step
        h = ((r * 1000003 + j * 7919) % 2147483647) * 48271 % 2147483647
        indices[r0 * per_row:(r0 + r.shape[0]) * per_row] = (off + step * j).to(torch.int32).reshape(-1)
        data[r0 * per_row:(r0 + r.shape[0]) * per_row] = ((h >> 9) % 7 + 1).to(torch.float32).reshape(-1)
    indptr = torch.arange(n_rows + 1, device=dev, dtype=torch.int64) * per_row
    d = _engine.DeviceMatrix(_engine.WideCSR(n_rows, n_cols, indptr, indices, data, sorted_rows=True), False)
    targets = np.unique(np.concatenate([np.arange(0, n_rows, n_rows // 40), [n_rows - 1, 1_073_741, 1_073_742]])).astype(np.int32)
    k = 50
    dot = sim.dot_product(d, k=k, target_rows=targets, format_output="csr", verbose=False)
    cos = sim.cosine(d, k=k, target_rows=targets, format_output="csr", verbose=False)
    assert dot.shape == (n_rows, n_rows) and len(_engine.wide_blocks(_engine.Ctx(dev), d.stored, True)) == 2
    norms = torch.zeros(n_rows, dtype=torch.float64, device=dev)
    for r0 in range(0, n_rows, 200_000):
        v = data[r0 * per_row:(r0 + 200_000) * per_row].double().view(-1, per_row)
        norms[r0:r0 + v.shape[0]] = (v * v).sum(1).sqrt()
    for t in targets.tolist():
        x = torch.zeros(n_cols, dtype=torch.float32, device=dev)
        x[indices[t * per_row:(t + 1) * per_row].long()] = data[t * per_row:(t + 1) * per_row]
        score = torch.empty(n_rows, dtype=torch.float32, device=dev)
        for r0 in range(0, n_rows, 200_000):
            sl = slice(r0 * per_row, min(r0 + 200_000, n_rows) * per_row)
            score[r0:r0 + 200_000] = (x[indices[sl].long()] * data[sl]).view(-1, per_row).sum(1)  # integers below 2^24: exact
        want = torch.topk(score, k).values.cpu().numpy()
        row = dot[[t]]
        got = np.sort(row.data)[::-1]
        np.testing.assert_array_equal(got, want, err_msg=f"dot product row {t}")
        sc = score.cpu().numpy()
        assert np.array_equal(sc[row.indices], row.data), f"dot product row {t}: a value sits at the wrong column"
        cs = (score.double() / (norms * norms[t])).cpu().numpy()
        crow = cos[[t]]
        np.testing.assert_allclose(np.sort(crow.data)[::-1], np.sort(cs)[::-1][:k], rtol=1e-5, err_msg=f"cosine row {t}")
        np.testing.assert_allclose(crow.data, cs[crow.indices], rtol=1e-5)


def test_int64_index_arrays_are_narrowed_on_the_device(monkeypatch):
    """scipy matrices with int64 indptr / indices below the int32 limits (s_plus.pyx:241-244 narrows them on the host):
    large arrays are uploaded as they are and narrowed by spy_narrow_index_dev."""
    m = random_csr(900, 700, 0.03, seed=31)
    m64 = sp.csr_array((m.data, m.indices.astype(np.int64), m.indptr.astype(np.int64)), shape=m.shape)
    want = sim.cosine(m, k=20, format_output="csr", verbose=False)
    monkeypatch.setattr(_engine, "STAGED_H2D_MIN_BYTES", 1)  # every array takes the large-array path
    before = _lib.launch_count()
    got = sim.cosine(m64, k=20, format_output="csr", verbose=False)
    assert _lib.launch_count() > before
    assert_topk_parity(want, got, k=20, rtol=1e-6, what="int64-indexed input")
    ref = oracle.similarity("cosine", m, k=20, format_output="csr", verbose=False)
    assert_topk_parity(ref, got, k=20, rtol=1e-5, what="int64-indexed input vs oracle")
    got_t = sim.cosine(m64.T, k=20, format_output="csr", verbose=False)  # CSC view with int64 arrays
    ref_t = oracle.similarity("cosine", m.T.tocsr(), k=20, format_output="csr", verbose=False)
    assert_topk_parity(ref_t, got_t, k=20, rtol=1e-5, what="int64-indexed CSC input vs oracle")
