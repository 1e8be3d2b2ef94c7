"""GPU parity tests: the CUDA path (through the public API -> C ABI) against the CPU oracle on the same
seeded inputs.  Values within rtol=1e-5 (north_star), bit-exact for integer-valued data; column-id sets
identical outside exact ties at the k boundary (parity.assert_topk_parity)."""
import numpy as np
import pytest
import scipy.sparse as sp

import similaripy_b200 as sim
from oracle import oracle
from parity import assert_topk_parity, check_sum, random_csr

pytestmark = pytest.mark.gpu

PRESETS = [
    ("dot_product", {}),
    ("cosine", {}),
    ("asymmetric_cosine", dict(alpha=0.2)),
    ("jaccard", {}),
    ("dice", {}),
    ("tversky", dict(alpha=0.8, beta=0.4)),
    ("p3alpha", dict(alpha=0.8)),
    ("rp3beta", dict(alpha=0.8, beta=0.4)),
    ("s_plus", dict(l1=0.5, l2=0.5, l3=1, t1=1, t2=1, c1=0.5, c2=0.5, alpha=1, beta1=0, beta2=0, pop1="none", pop2="sum")),
]


def both(name, m, m2=None, k=50, **kw):
    got = getattr(sim, name)(m.copy(), None if m2 is None else m2.copy(), k=k, verbose=False, format_output="csr", **kw)
    ref = oracle.similarity(name, m.copy(), None if m2 is None else m2.copy(), k=k, verbose=False, format_output="csr", **kw)
    return ref, got


@pytest.mark.parametrize("name,kw", PRESETS, ids=[p[0] for p in PRESETS])
def test_topk_all_similarities(name, kw):
    # shape of the reference's test_similarity_topk (tests/test_similarity.py:289-300)
    m = random_csr(1000, 800, 0.025, seed=42)
    ref, got = both(name, m, k=50, **kw)
    assert_topk_parity(ref, got, k=50, rtol=1e-5, what=name)
    np.testing.assert_allclose(check_sum(got), check_sum(ref), rtol=1e-5)


@pytest.mark.parametrize("name,kw", PRESETS, ids=[p[0] for p in PRESETS])
def test_full_rows_all_similarities(name, kw):
    # k = n_cols: no truncation, every entry must match (tests/test_similarity.py:303-314)
    m = random_csr(400, 50, 0.025, seed=42)
    ref, got = both(name, m, k=400, **kw)
    assert_topk_parity(ref, got, k=400, rtol=1e-5, what=name)


@pytest.mark.parametrize("mode", ["stabilized", "bayesian", "additive"])
def test_shrink_types(mode):
    m = random_csr(400, 50, 0.025, seed=42)
    for name in ("cosine", "tversky", "asymmetric_cosine"):
        ref, got = both(name, m, k=50, shrink=10, shrink_type=mode)
        assert_topk_parity(ref, got, k=50, rtol=1e-5, what=f"{name}/{mode}")


def test_integer_data_bit_exact():
    m = random_csr(600, 300, 0.05, seed=3, integer=True)
    ref, got = both("dot_product", m, k=20)
    assert_topk_parity(ref, got, k=20, rtol=0.0, what="integer dot_product")


def test_binary_ties():
    m = random_csr(300, 120, 0.1, seed=7)
    for name in ("jaccard", "dot_product", "cosine"):
        ref, got = both(name, m, k=10, binary=True)
        stats = assert_topk_parity(ref, got, k=10, rtol=1e-6, what=f"binary {name}")
        assert stats["full_rows"] > 0
