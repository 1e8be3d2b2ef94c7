"""CPU: the oracle against the compiled, unmodified reference (oracle/_ref, built by oracle/build_ref.py)
on seeded inputs larger than the golden fixtures, including the reference's own test shapes
(tests/test_similarity.py:289-314: 1000x800 d=0.025 k=50; 400x50 full).  Skipped when oracle/_ref is not
built (it is built by __graft_entry__.build() wherever /root/reference exists and travels to the GPU box)."""
import numpy as np
import pytest

from oracle import oracle, ref_api
from parity import assert_topk_parity, check_sum, random_csr

pytestmark = pytest.mark.skipif(not ref_api.available(), reason="oracle/_ref (compiled reference) not built")

PRESETS = [
    ("dot_product", {}), ("cosine", {}), ("asymmetric_cosine", dict(alpha=0.2)), ("jaccard", {}), ("dice", {}),
    ("tversky", dict(alpha=0.8, beta=0.4)), ("p3alpha", dict(alpha=0.8)), ("rp3beta", dict(alpha=0.8, beta=0.4)),
    ("s_plus", dict(l1=0.5, l2=0.5, l3=1, t1=1, t2=1, c1=0.5, c2=0.5, alpha=1, beta1=0, beta2=0, pop1="none", pop2="sum")),
]


@pytest.mark.parametrize("name,kw", PRESETS, ids=[p[0] for p in PRESETS])
def test_reference_topk_shape(name, kw):
    m = random_csr(1000, 800, 0.025, seed=42)
    a = oracle.similarity(name, m.copy(), k=50, format_output="csr", **kw)
    b = ref_api.similarity(name, m.copy(), k=50, format_output="csr", **kw)
    assert_topk_parity(b, a, k=50, rtol=2e-6, what=name)
    np.testing.assert_allclose(check_sum(a), check_sum(b), rtol=1e-5)


@pytest.mark.parametrize("name,kw", PRESETS, ids=[p[0] for p in PRESETS])
def test_reference_full_shape(name, kw):
    m = random_csr(400, 50, 0.025, seed=42)
    a = oracle.similarity(name, m.copy(), k=400, format_output="csr", **kw)
    b = ref_api.similarity(name, m.copy(), k=400, format_output="csr", **kw)
    assert_topk_parity(b, a, k=400, rtol=2e-6, what=name)


@pytest.mark.parametrize("block_size", [None, 0, 64, 256])
def test_binary_ties_exact_slab(block_size):
    """Exact ties everywhere: the oracle must keep the same columns in the same heap order as the reference,
    in the unblocked and in the blocked + popularity-permuted modes (SURVEY 8c tie rule)."""
    m = random_csr(300, 120, 0.1, seed=7)
    for name in ("jaccard", "dot_product"):
        a = oracle.similarity(name, m.copy(), k=10, binary=True, format_output="coo", block_size=block_size)
        b = ref_api.similarity(name, m.copy(), k=10, binary=True, format_output="coo", block_size=block_size, num_threads=2)
        np.testing.assert_array_equal(a.row, b.row)
        np.testing.assert_array_equal(a.col, b.col)
        np.testing.assert_array_equal(a.data, b.data)


def test_default_blocking_kicks_in_above_262144_columns():
    """n_cols > 262144 with block_size=0 takes the reference's blocked path (s_plus.h:33,311)."""
    rng = np.random.default_rng(5)
    import scipy.sparse as sp
    n_cols = 300_000
    a = sp.random_array((40, 500), density=0.05, format="csr", dtype=np.float32, random_state=rng)
    b = sp.random_array((500, n_cols), density=2e-4, format="csr", dtype=np.float32, random_state=rng)
    x = oracle.similarity("cosine", a.copy(), b.copy(), k=20, format_output="csr")
    y = ref_api.similarity("cosine", a.copy(), b.copy(), k=20, format_output="csr")
    assert_topk_parity(y, x, k=20, rtol=2e-6, what="default blocking")


@pytest.mark.parametrize("fn,kw", [("normalize", dict(norm="l1")), ("normalize", dict(norm="l2")), ("normalize", dict(norm="max")),
                                   ("bm25", {}), ("bm25plus", dict(delta=0.7)), ("tfidf", {}),
                                   ("tfidf", dict(tf_mode="log", idf_mode="prob", logbase=2.0))])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_normalizers(fn, kw, dtype):
    m = random_csr(2000, 300, 0.03, seed=9, dtype=dtype)
    a = getattr(oracle, fn)(m.copy(), **kw)
    b = getattr(ref_api, fn)(m.copy(), **kw)
    np.testing.assert_allclose(a.data, b.data, rtol=2e-6 if dtype == np.float32 else 1e-12)
