"""GPU: tuning={"tie_mode": "reference"} -- the reference's OWN order (csrc/knn_reforder.cu): float32 sums built entry by
entry like s_plus.h:418-438 (bit-identical values) and ties at the k-th value kept like TopK's heap keeps them
(s_plus.h:45-59), including the blocked path with its popularity permutation (s_plus_utils.pyx:493-618).
Checked as EXACT equality of every row's {(column, value bits)} set against the golden slabs the unmodified reference
produced and against the oracle on tie-heavy data."""
import numpy as np
import pytest
import scipy.sparse as sp

import golden_io
import similaripy_b200 as sim
from oracle import oracle
from parity import random_csr

pytestmark = pytest.mark.gpu
REF = {"tie_mode": "reference"}


def assert_identical_rows(ref, got, what, exact_values=True):
    """Same stored column set in every row; values compared bit for bit (exact_values) or to 2e-6 relative -- the
    denominators' norm vectors come from powf on the GPU and from numpy on the host, which may differ in the last bit."""
    ref, got = ref.tocsr().copy(), got.tocsr().copy()
    ref.eliminate_zeros(); got.eliminate_zeros()
    ref.sort_indices(); got.sort_indices()
    assert ref.shape == got.shape, what
    bad = [r for r in range(ref.shape[0])
           if not np.array_equal(ref.indices[ref.indptr[r]:ref.indptr[r + 1]], got.indices[got.indptr[r]:got.indptr[r + 1]])]
    assert not bad, f"{what}: the column sets of {len(bad)} of {ref.shape[0]} rows differ, first {bad[:5]}"
    if exact_values:
        assert np.array_equal(ref.data.view(np.uint32), got.data.view(np.uint32)), f"{what}: values are not bit-identical"
    else:
        np.testing.assert_allclose(got.data, ref.data, rtol=2e-6, err_msg=what)


@pytest.mark.parametrize("name", golden_io.similarity_cases())
def test_every_golden_slab_is_reproduced_exactly(name):
    g = golden_io.load_similarity(name)
    kw = dict(g["kw"])
    kw.setdefault("format_output", "csr")
    kw["format_output"] = "csr"
    got = getattr(sim, g["fn"])(g["m1"].copy(), None if g["m2"] is None else g["m2"].copy(), verbose=False, tuning=REF, **kw)
    # dot products (no denominator vectors) must agree bit for bit; the others to the last bit or two of powf
    assert_identical_rows(g["ref_csr"], got, f"golden {name}", exact_values=g["fn"] == "dot_product")


@pytest.mark.parametrize("block_size", [None, 0, 32, 64, 100])
@pytest.mark.parametrize("name", ["jaccard", "dot_product", "cosine", "dice"])
def test_binary_data_ties_everywhere(name, block_size):
    m = random_csr(300, 120, 0.1, seed=7)
    ref = oracle.similarity(name, m.copy(), k=10, binary=True, block_size=block_size, format_output="csr")
    got = getattr(sim, name)(m.copy(), k=10, binary=True, block_size=block_size, verbose=False, format_output="csr", tuning=REF)
    assert_identical_rows(ref, got, f"binary {name} block_size={block_size}", exact_values=name in ("jaccard", "dot_product", "dice"))


@pytest.mark.parametrize("block_size", [None, 48])
def test_integer_counts_with_selectors_and_rectangular_operands(block_size):
    a = random_csr(250, 180, 0.08, seed=3, integer=True)
    b = random_csr(180, 220, 0.08, seed=4, integer=True)
    filt = random_csr(250, 220, 0.1, seed=5)
    rows = np.arange(0, 250, 3, dtype=np.int32)
    for kw in (dict(), dict(filter_cols=filt), dict(target_cols=filt), dict(filter_cols=[1, 5, 9, 100], target_rows=rows)):
        ref = oracle.similarity("dot_product", a.copy(), b.copy(), k=15, block_size=block_size, format_output="csr", **kw)
        got = sim.dot_product(a.copy(), b.copy(), k=15, block_size=block_size, verbose=False, format_output="csr", tuning=REF, **kw)
        assert_identical_rows(ref, got, f"integer dot {sorted(kw)} block_size={block_size}")


@pytest.mark.parametrize("name,kw", [("cosine", {}), ("rp3beta", dict(alpha=0.8, beta=0.4)), ("tversky", dict(alpha=0.7, beta=0.3)),
                                     ("s_plus", dict(l1=0.5, l2=0.5, l3=1, pop2="sum", shrink=3.0))])
def test_float_data_same_sets_and_values(name, kw):
    """The sums are built in the reference's order; what is left is the last bit of the norm vectors (powf)."""
    m = random_csr(400, 300, 0.05, seed=11)
    ref = oracle.similarity(name, m.copy(), k=20, block_size=None, format_output="csr", **kw)
    got = getattr(sim, name)(m.copy(), k=20, block_size=None, verbose=False, format_output="csr", tuning=REF, **kw)
    assert_identical_rows(ref, got, f"float {name}", exact_values=False)


def test_mixed_sign_data_zero_sum_quirk():
    """A column whose partial sum returns to exactly 0 is enlisted twice by the reference (s_plus.h:112-117): an extra
    zero-valued candidate.  With threshold <= 0 and a large k it shows up as an explicit zero in the COO slab."""
    rng = np.random.default_rng(2)
    m = sp.random_array((60, 40), density=0.3, format="csr", dtype=np.float32, random_state=rng)
    m.data = rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=m.nnz)
    ref = oracle.similarity("dot_product", m.copy(), k=40, threshold=-1e9, block_size=None, format_output="csr")
    got = sim.dot_product(m.copy(), k=40, threshold=-1e9, block_size=None, verbose=False, format_output="csr", tuning=REF)
    assert_identical_rows(ref, got, "mixed-sign dot")
