"""GPU parity tests: the CUDA path (through the public API -> C ABI) against the CPU oracle on the same
seeded inputs.  Values within rtol=1e-5 (north_star), bit-exact for integer-valued data; column-id sets
identical outside exact ties at the k boundary (parity.assert_topk_parity)."""
import numpy as np
import pytest
import scipy.sparse as sp

import similaripy_b200 as sim
from oracle import oracle
from parity import assert_topk_parity, check_sum, random_csr

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("spy_engine")]

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


# ---- selectors, edge cases, README flow (reference tests/test_similarity.py:359-617) ----------------------
def test_readme_flow_bm25_cosine_recommend():
    urm = sp.random_array((1000, 2000), density=0.025, format="csr", dtype=np.float32,
                          random_state=np.random.default_rng(0))
    w = sim.bm25(urm)
    w_ref = oracle.bm25(urm)
    np.testing.assert_allclose(w.data, w_ref.data, rtol=1e-5)
    model = sim.cosine(w.T, k=50, verbose=False, format_output="csr")
    model_ref = oracle.similarity("cosine", w_ref.T.tocsr(), k=50, format_output="csr")
    assert_topk_parity(model_ref, model, k=50, rtol=1e-5, what="readme cosine")
    users = [1, 14, 8, 999]
    rec = sim.dot_product(urm, model_ref.T, k=100, target_rows=users, filter_cols=urm, verbose=False, format_output="csr")
    rec_ref = oracle.similarity("dot_product", urm, model_ref.T, k=100, target_rows=users, filter_cols=urm, format_output="csr")
    assert rec.shape == (1000, 2000)
    assert_topk_parity(rec_ref, rec, k=100, rtol=1e-5, what="readme recommend")
    assert set(np.unique(rec.tocoo().row).tolist()) <= set(users)
    assert rec.multiply(urm).nnz == 0  # seen items are filtered


@pytest.mark.parametrize("panel_width", [128, 256, 1024])
def test_multi_panel_matches_single_panel(panel_width):
    m = random_csr(500, 1500, 0.02, seed=11)
    for name, kw in (("cosine", {}), ("jaccard", {}), ("rp3beta", dict(alpha=0.7, beta=0.3))):
        ref = oracle.similarity(name, m.T.tocsr(), k=30, format_output="csr", **kw)
        got = getattr(sim, name)(m.T.tocsr(), k=30, verbose=False, format_output="csr",
                                 tuning=dict(panel_width=panel_width), **kw)
        assert_topk_parity(ref, got, k=30, rtol=1e-5, what=f"{name} W={panel_width}")


@pytest.mark.parametrize("threads", [512, 1024])
def test_launch_shapes(threads):
    m = random_csr(400, 300, 0.05, seed=12, integer=True)
    ref = oracle.similarity("dot_product", m, k=25, format_output="csr")
    got = sim.dot_product(m, k=25, verbose=False, format_output="csr", tuning=dict(threads=threads))
    assert_topk_parity(ref, got, k=25, rtol=0.0, what=f"threads={threads}")


def test_long_rows_span_several_staging_chunks():
    """A rows with more stored entries than threads per CTA (several staging chunks) and a few empty ones."""
    rng = np.random.default_rng(31)
    a = sp.random_array((40, 6000), density=0.5, format="csr", dtype=np.float32, random_state=rng)
    b = sp.random_array((6000, 700), density=0.01, format="csr", dtype=np.float32, random_state=rng)
    for threads in (512, 1024):
        ref = oracle.similarity("cosine", a.copy(), b.copy(), k=30, format_output="csr")
        got = sim.cosine(a.copy(), b.copy(), k=30, verbose=False, format_output="csr",
                         tuning=dict(threads=threads, panel_width=256))
        assert_topk_parity(ref, got, k=30, rtol=1e-5, what=f"long rows threads={threads}")


def test_large_k_uses_global_candidate_buffer():
    m = random_csr(200, 9000, 0.01, seed=13)
    ref = oracle.similarity("dot_product", m.T.tocsr(), k=5000, format_output="csr")
    got = sim.dot_product(m.T.tocsr(), k=5000, verbose=False, format_output="csr")
    assert_topk_parity(ref, got, k=5000, rtol=1e-5, what="k=5000")


def test_empty_and_degenerate_inputs():
    empty = sp.csr_array((50, 40), dtype=np.float32)
    got = sim.cosine(empty, k=5, verbose=False, format_output="csr")
    assert got.shape == (50, 50) and got.nnz == 0
    coo = sim.cosine(empty, k=5, verbose=False, format_output="coo")
    assert coo.data.shape[0] == 50 * 5 and not coo.data.any()     # the zero-padded slab, like the reference
    m = random_csr(60, 40, 0.1, seed=14)
    m = m.tolil(); m[5, :] = 0; m[:, 7] = 0; m = m.tocsr()         # empty row, empty column, explicit zeros dropped
    m.data[:3] = 0.0
    ref, got = both("cosine", m, k=7)
    assert_topk_parity(ref, got, k=7, rtol=1e-5, what="ragged")
    ref, got = both("dot_product", m, k=1000)                      # k clipped to n_cols (s_plus.pyx:187-188)
    assert_topk_parity(ref, got, k=60, rtol=1e-5, what="k clipped")
    got = sim.dot_product(m, k=3, target_rows=[], verbose=False, format_output="csr")
    assert got.nnz == 0 and got.shape == (60, 60)


def test_duplicate_target_rows_and_dtypes():
    m = random_csr(80, 50, 0.1, seed=15)
    rows = [4, 4, 9, 4]
    ref = oracle.similarity("cosine", m, k=6, target_rows=rows, format_output="csr")
    got = sim.cosine(m, k=6, target_rows=rows, verbose=False, format_output="csr")
    assert got.nnz == ref.nnz
    np.testing.assert_allclose(np.sort(got.data), np.sort(ref.data), rtol=1e-5)
    for dt in (np.float64, np.int32, np.int64):
        mi = random_csr(80, 50, 0.1, seed=15, integer=True).astype(dt)
        mi = sp.csr_array((mi.data, mi.indices.astype(np.int64), mi.indptr.astype(np.int64)), shape=mi.shape)
        ref = oracle.similarity("dot_product", mi, k=6, format_output="csr")
        got = sim.dot_product(mi, k=6, verbose=False, format_output="csr")
        assert got.dtype == np.float32
        assert_topk_parity(ref, got, k=6, rtol=0.0, what=f"dtype {dt}")


def test_selectors_against_oracle():
    m = random_csr(300, 200, 0.05, seed=16)
    fm = random_csr(300, 300, 0.2, seed=17)
    tm = random_csr(300, 300, 0.3, seed=18)
    cases = [dict(filter_cols=[0, 5, 7, 299]), dict(target_cols=np.arange(0, 300, 2)), dict(filter_cols=fm),
             dict(target_cols=tm), dict(filter_cols=fm, target_cols=tm), dict(filter_cols=[1, 2], target_cols=tm),
             dict(filter_cols=fm, target_cols=list(range(100))), dict(threshold=0.2), dict(binary=True, filter_cols=[3, 4])]
    for kw in cases:
        ref = oracle.similarity("cosine", m.copy(), k=15, format_output="csr", **kw)
        got = sim.cosine(m.copy(), k=15, verbose=False, format_output="csr", **kw)
        assert_topk_parity(ref, got, k=15, rtol=1e-5, what=str(list(kw)))


def test_input_side_effects_like_reference():
    """eliminate_zeros() mutates a CSR argument in place (s_plus.pyx:210-211); data is otherwise untouched."""
    m = random_csr(100, 80, 0.1, seed=19)
    m.data[::7] = 0.0
    nnz_before, data_before = m.nnz, m.data.copy()
    sim.cosine(m, k=5, verbose=False, binary=True)
    assert m.nnz < nnz_before
    np.testing.assert_array_equal(m.data, data_before[data_before != 0])


def test_device_matrix_chain_matches_host_path():
    urm = random_csr(900, 400, 0.03, seed=20)
    host_model = sim.cosine(sim.bm25(urm).T, k=20, verbose=False, format_output="csr")
    d = sim.to_device(urm)
    model = sim.cosine(sim.bm25(d).T, k=20, verbose=False, on_device=True)
    assert isinstance(model, sim.DeviceMatrix) and model.shape == (400, 400)
    assert_topk_parity(host_model, sim.to_host(model), k=20, rtol=1e-6, what="device chain cosine")
    rec = sim.dot_product(d, model.T, k=10, filter_cols=urm, target_rows=[0, 5, 899], verbose=False, format_output="csr")
    rec_ref = oracle.similarity("dot_product", urm, host_model.T, k=10, filter_cols=urm, target_rows=[0, 5, 899], format_output="csr")
    assert_topk_parity(rec_ref, rec, k=10, rtol=1e-5, what="device chain recommend")
    for name, kw in (("rp3beta", dict(alpha=0.8, beta=0.5)), ("p3alpha", dict(alpha=1.3))):
        ref = oracle.similarity(name, urm, k=12, format_output="csr", **kw)
        got = sim.to_host(getattr(sim, name)(d, k=12, verbose=False, on_device=True, **kw))
        assert_topk_parity(ref, got, k=12, rtol=1e-5, what=f"device {name}")


def test_handles_are_ordered_across_streams():
    """A DeviceMatrix produced on one stream and consumed by a later call on another stream: the consumer waits for the
    producer's event (Ctx.produced / Ctx.consume); the caller does not have to synchronise."""
    import torch
    urm = random_csr(3000, 1200, 0.02, seed=23)
    host_model = oracle.similarity("cosine", urm.T.tocsr(), k=20, format_output="csr")
    d = sim.to_device(urm)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        model = sim.cosine(d.T, k=20, verbose=False, on_device=True)   # queued on `side`, nothing synchronises
        sim.normalize(model, norm="l1", axis=1, inplace=True)          # in place, still on `side`
    assert model.stored.ready is not None
    rec = sim.dot_product(d, model.T, k=10, target_rows=[0, 7, 2999], verbose=False, format_output="csr")  # default stream
    want_model = oracle.normalize(host_model, norm="l1", axis=1)
    rec_ref = oracle.similarity("dot_product", urm, want_model.T.tocsr(), k=10, target_rows=[0, 7, 2999], format_output="csr")
    assert_topk_parity(rec_ref, rec, k=10, rtol=1e-4, what="handle consumed on another stream")


# ---- size-independent properties at a size the oracle cannot finish in seconds -------------------------
def test_properties_at_scale():
    rng = np.random.default_rng(21)
    m = sp.random_array((20_000, 30_000), density=0.002, format="csr", dtype=np.float32, random_state=rng)
    k = 40
    a = sim.cosine(m, k=k, verbose=False, format_output="csr")
    # (0) at most k per row, all > 0, values best-first inside each CSR row (our slab order) -- checked before
    #     any scipy call that sorts the indices of `a` in place (max / sum_duplicates do)
    assert np.diff(a.indptr).max() <= k and a.data.min() > 0
    for r in (0, 1234, 19_999):
        assert np.all(np.diff(a.data[a.indptr[r]:a.indptr[r + 1]]) <= 0)
    # (1) symmetry of cosine: s(i,j) == s(j,i) wherever both directions were kept
    at = a.T.tocsr()
    both_kept = a.multiply(at > 0)
    d = abs(both_kept - at.multiply(a > 0))
    assert d.max() <= 1e-5
    # (2) self-similarity is 1 and is the row maximum for every non-empty row
    diag = a.diagonal()
    nonempty = np.diff(m.indptr) > 0
    np.testing.assert_allclose(diag[nonempty], 1.0, rtol=1e-5)
    assert np.all(a.max(axis=1).toarray().ravel()[nonempty] <= 1.0 + 1e-5)
    # (4) idempotence of target_rows: a subset call returns exactly those rows
    sub = [5, 777, 19_999]
    b = sim.cosine(m, k=k, target_rows=sub, verbose=False, format_output="csr")
    b.sort_indices()  # `a` was index-sorted in place by scipy above
    for r in sub:
        np.testing.assert_array_equal(b.indices[b.indptr[r]:b.indptr[r + 1]], a.indices[a.indptr[r]:a.indptr[r + 1]])
        np.testing.assert_allclose(b.data[b.indptr[r]:b.indptr[r + 1]], a.data[a.indptr[r]:a.indptr[r + 1]], rtol=1e-6)
    # (5) linearity of dot_product in the values: scaling A by 2 scales every kept value by 2, same columns
    d1 = sim.dot_product(m, k=k, target_rows=sub, verbose=False, format_output="csr")
    m2 = m.copy(); m2.data *= 2
    d2 = sim.dot_product(m2, m.T.tocsr(), k=k, target_rows=sub, verbose=False, format_output="csr")
    np.testing.assert_array_equal(d1.indices, d2.indices)
    np.testing.assert_allclose(d2.data, 2 * d1.data, rtol=1e-6)
    # (6) spot rows against the oracle
    ref = oracle.similarity("cosine", m, k=k, target_rows=sub, format_output="csr")
    assert_topk_parity(ref, b, k=k, rtol=1e-5, what="scale spot rows")


# ---- dense candidate sets: buffer overflow, sampling rounds, coarse block-minimum filter, several panels -----------
_DENSE = None


def _dense_urm():
    """3000 users x 20003 items, 2 % dense: an item row of URM.T expands to ~24k products over ~14k distinct columns,
    far more than the 2048-entry candidate buffer (n_cols deliberately not a multiple of 4)."""
    global _DENSE
    if _DENSE is None:
        rng = np.random.default_rng(77)
        _DENSE = sp.random_array((3000, 20003), density=0.02, format="csr", dtype=np.float32, random_state=rng)
    return _DENSE


@pytest.mark.parametrize("name,kw", PRESETS + [("cosine", dict(shrink=5.0, shrink_type="bayesian")),
                                                ("tversky", dict(alpha=0.3, beta=0.9, shrink=2.0))],
                         ids=[p[0] for p in PRESETS] + ["cosine_bayes", "tversky_shrink"])
@pytest.mark.parametrize("tuning", [None, dict(panel_width=4096), dict(threads=512, panel_width=2048, group=16)],
                         ids=["plan", "panels5", "t512g16"])
def test_dense_candidates_overflow_paths(name, kw, tuning):
    urm = _dense_urm()
    a = urm.T.tocsr()
    rows = np.arange(0, a.shape[0], 97, dtype=np.int32)  # 207 target rows
    k = 100
    got = getattr(sim, name)(a, urm, k=k, target_rows=rows, verbose=False, format_output="csr", tuning=tuning, **kw)
    ref = oracle.similarity(name, a, urm, k=k, target_rows=rows, verbose=False, format_output="csr", **kw)
    stats = assert_topk_parity(ref, got, k=k, rtol=1e-5, what=f"dense {name} {tuning}")
    assert stats["full_rows"] >= 200


@pytest.mark.parametrize("k", [1, 7, 700, 3000, 5000])  # 5000: candidate buffer in global memory
def test_dense_candidates_k_extremes(k):
    urm = _dense_urm()
    a = urm.T.tocsr()
    rows = np.arange(5, a.shape[0], 401, dtype=np.int32)
    got = sim.cosine(a, urm, k=k, target_rows=rows, verbose=False, format_output="csr")
    ref = oracle.similarity("cosine", a, urm, k=k, target_rows=rows, verbose=False, format_output="csr")
    assert_topk_parity(ref, got, k=k, rtol=1e-5, what=f"dense k={k}")


def test_dense_candidates_binary_integer_exact():
    """Integer-valued data: dot products are order independent, so the values must be bit-exact."""
    urm = _dense_urm().copy()
    urm.data = np.floor(urm.data * 4).astype(np.float32) + 1.0
    a = urm.T.tocsr()
    rows = np.arange(3, a.shape[0], 211, dtype=np.int32)
    got = sim.dot_product(a, urm, k=150, target_rows=rows, verbose=False, format_output="csr", tuning=dict(panel_width=8192))
    ref = oracle.similarity("dot_product", a, urm, k=150, target_rows=rows, verbose=False, format_output="csr")
    assert_topk_parity(ref, got, k=150, rtol=0.0, what="dense integer dot")


# ---- skewed (power-law) data: a few very long rows and columns, most of them short -------------------------------------
def _zipf_urm(n_users, n_items, nnz, seed):
    rng = np.random.default_rng(seed)
    pu = 1.0 / np.arange(1, n_users + 1) ** 0.9
    pi = 1.0 / np.arange(1, n_items + 1) ** 1.1
    u = rng.choice(n_users, size=nnz, p=pu / pu.sum())
    i = rng.choice(n_items, size=nnz, p=pi / pi.sum())
    m = sp.csr_array((np.ones(nnz, dtype=np.float32), (rng.permutation(n_users)[u], rng.permutation(n_items)[i])),
                     shape=(n_users, n_items))
    m.sum_duplicates()
    m.data = (1.0 + rng.random(m.nnz)).astype(np.float32)  # counts replaced by ratings in [1, 2)
    return m


@pytest.mark.parametrize("name,kw", [("cosine", {}), ("rp3beta", dict(alpha=0.7, beta=0.5)), ("jaccard", dict(binary=True)),
                                     ("dot_product", {})], ids=["cosine", "rp3beta", "jaccard_binary", "dot"])
@pytest.mark.parametrize("tuning", [None, dict(threads=512, panel_width=1024)], ids=["plan", "panels"])
def test_power_law_data(name, kw, tuning):
    """Head items are bought by most users (A rows of several thousand entries: many staging chunks, dynamic batch
    claims), head users own hundreds of items (B segments far longer than the group width), the tail is nearly empty."""
    urm = _zipf_urm(6000, 5000, 150_000, 13)
    a = urm.T.tocsr()
    assert np.diff(a.indptr).max() > 2600  # longer than two staging chunks of 1280 entries
    rows = np.concatenate([np.argsort(-np.diff(a.indptr))[:40], np.arange(0, 5000, 53)]).astype(np.int32)
    rows = np.unique(rows)
    k = 60
    got = getattr(sim, name)(a, urm, k=k, target_rows=rows, verbose=False, format_output="csr", tuning=tuning, **kw)
    ref = oracle.similarity(name, a, urm, k=k, target_rows=rows, verbose=False, format_output="csr", **kw)
    # jaccard on binary data is all ties: values must agree exactly, columns only outside the tie band
    assert_topk_parity(ref, got, k=k, rtol=1e-5, what=f"zipf {name} {tuning}")
