"""Hand-over protocol of the stream kernel (csrc/knn_stream_kernel.cuh): dense snapshots (the whole of TMEM) and sparse
hand-overs ((column, sum) pairs, double buffered in two TMEM regions) in every order, in both builds of the kernel.
A panel is handed over densely when more than 6144 of its slots were touched, so the matrices here are wide enough for
that: target rows alternate between long rows (dense panels) and short ones (sparse panels), with runs of either kind,
empty rows in between and a last panel that is narrower than the others.  Checked against the oracle
(s_plus.h:265-453 restated in oracle/spy_oracle.c)."""
import numpy as np
import pytest
import scipy.sparse as sp

import similaripy_b200 as sim
from oracle import oracle
from parity import assert_topk_parity

pytestmark = pytest.mark.gpu


def _operands(seed, n_u=2000, n_cols=25000, b_per_row=250):
    rng = np.random.default_rng(seed)
    b = sp.random_array((n_u, n_cols), density=b_per_row / n_cols, format="csr", dtype=np.float32, random_state=rng)
    lens = []
    pattern = [150, 5, 150, 150, 3, 0, 7, 150, 1, 1, 150, 0, 0, 150, 4, 150]  # long = dense panels, short = sparse, 0 = empty row
    for i in range(160):
        lens.append(pattern[i % len(pattern)])
    rows, cols, vals = [], [], []
    for r, n in enumerate(lens):
        c = rng.choice(n_u, size=n, replace=False)
        rows += [r] * n; cols += c.tolist(); vals += (rng.random(n).astype(np.float32) + 0.1).tolist()
    a = sp.csr_array((np.asarray(vals, np.float32), (np.asarray(rows), np.asarray(cols))), shape=(len(lens), n_u))
    return a, b


@pytest.mark.parametrize("drain_warps", [8, 16])
@pytest.mark.parametrize("name,kw", [("dot_product", {}), ("cosine", {}), ("tversky", dict(alpha=0.7, beta=0.3)),
                                     ("jaccard", dict(binary=True)), ("cosine", dict(binary=True)), ("dot_product", dict(binary=True))],
                         ids=["dot", "cosine", "tversky", "jaccard_binary", "cosine_binary", "dot_binary"])
def test_dense_and_sparse_handovers_interleaved(drain_warps, name, kw):
    a, b = _operands(41)
    tuning = dict(engine="stream", drain_warps=drain_warps, panel_width=10240)  # 3 panels: 10240 + 10240 + 4520 columns
    ref = oracle.similarity(name, a, b, k=40, format_output="csr", verbose=False, **kw)
    got = getattr(sim, name)(a, b, k=40, format_output="csr", verbose=False, tuning=tuning, **kw)
    assert_topk_parity(ref, got, k=40, rtol=1e-5, what=f"{name} D={drain_warps}")
    if kw.get("binary"):  # the counting form of the panel (integer adds) against the float adds: bit-identical values
        flt = getattr(sim, name)(a, b, k=40, format_output="csr", verbose=False, tuning=dict(tuning, unit_values=False), **kw)
        assert np.array_equal(np.sort(flt.data), np.sort(got.data)) and flt.nnz == got.nnz
    # every long row must have produced more candidates than the sparse hand-over can list (else this test tests nothing)
    full = oracle.similarity("dot_product", a[[0]], b, k=25000, format_output="csr", verbose=False)
    assert full.nnz > 3 * 6144


@pytest.mark.parametrize("drain_warps", [8, 16])
def test_handover_with_matrix_filter_and_threshold(drain_warps):
    a, b = _operands(42)
    flt = sp.random_array((a.shape[0], b.shape[1]), density=0.02, format="csr", dtype=np.float32,
                          random_state=np.random.default_rng(43))
    tuning = dict(engine="stream", drain_warps=drain_warps, panel_width=10240)
    kw = dict(k=30, filter_cols=flt, threshold=0.2, format_output="csr", verbose=False)
    ref = oracle.similarity("dot_product", a, b, **kw)
    got = sim.dot_product(a, b, tuning=tuning, **kw)
    assert_topk_parity(ref, got, k=30, rtol=1e-5, what=f"filter + threshold D={drain_warps}")
    assert got.multiply(flt).nnz == 0
