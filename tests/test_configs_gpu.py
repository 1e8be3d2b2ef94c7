"""The BASELINE.json configurations at their full matrix sizes (synthetic, generated on the device).

The oracle cannot expand these in seconds, so parity is checked through (a) an independent float64 restatement in
torch of single rows -- dense accumulation of the row's expansion, the computeSimilarity formula
(similaripy/cython_code/s_plus.h:129-156) and a full sort -- fed from the SAME device-resident operands and norm
vectors the kernel gets, (b) size-independent properties (bounded row lengths, best-first order, symmetry, filters
respected, agreement between different launch plans), and for cfg2 (c) spot rows against the CPU oracle."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("spy_engine")]


def _gen(n_rows, n_cols, density, seed):
    import torch
    import bench
    import similaripy_b200 as sim
    from similaripy_b200 import _engine
    ip, ix, dv = bench.gen_urm_device(n_rows, n_cols, density, seed, torch.device("cuda", 0))
    return sim.DeviceMatrix(_engine.DeviceCSR(n_rows, n_cols, ip, ix, dv, sorted_rows=True), False)


def _row_reference(job, t, k):
    """Row t of the kernel's job in float64: returns (cols, vals) best-first (value desc, column asc)."""
    import torch
    A, B, P, v = job.A, job.B, job.params, job.vectors
    a0, a1 = int(A.indptr[t]), int(A.indptr[t + 1])
    us = A.indices[a0:a1].long()
    av = A.data[a0:a1].double()
    starts, ends = B.indptr[us].long(), B.indptr[us + 1].long()
    lens = ends - starts
    seg = torch.repeat_interleave(torch.arange(us.numel(), device=us.device), lens)
    pos = torch.arange(int(lens.sum()), device=us.device) - torch.repeat_interleave(torch.cumsum(lens, 0) - lens, lens)
    q = starts[seg] + pos
    acc = torch.zeros(job.n_cols, dtype=torch.float64, device=us.device)
    touched = torch.zeros(job.n_cols, dtype=torch.bool, device=us.device)
    cols = B.indices[q].long()
    acc.index_add_(0, cols, av[seg] * B.data[q].double())
    touched[cols] = True
    xy = acc
    z = torch.zeros((), dtype=torch.float64, device=us.device)
    vT = P["l1"] * (P["t1"] * (v["Xt"][t].double() - xy) + P["t2"] * (v["Yt"].double() - xy) + xy) if P["l1"] != 0 else z
    vC = P["l2"] * v["Xc"][t].double() * v["Yc"].double() if P["l2"] != 0 else z
    vD = P["l3"] * v["Xd"][t].double() * v["Yd"].double() if P["l3"] != 0 else z
    num = xy if P["a1"] == 1.0 else xy.pow(P["a1"])
    if any(P[n] != 0 for n in ("l1", "l2", "l3", "stabilized_shrink", "bayesian_shrink")):
        den = vT + vC + vD + P["stabilized_shrink"]
        val = torch.where(den != 0, num / den, torch.zeros_like(num))
        if P["bayesian_shrink"] != 0:
            val = val * num / (num + P["bayesian_shrink"])
    else:
        val = xy
    ok = touched & (val >= P["threshold"])
    idx = torch.nonzero(ok).ravel()
    vals = val[idx]
    order = torch.argsort(vals, descending=True, stable=True)[:k]  # idx ascending + stable => ties by column asc
    return idx[order].cpu().numpy(), vals[order].cpu().numpy()


def _check_rows(job, rows, k, what, rtol=1e-5):
    """Kernel slab rows against the float64 restatement, tie-aware at the k boundary."""
    import torch
    torch.cuda.synchronize()
    targets = job.targets.cpu().numpy()
    cols = job.out_cols.cpu().numpy().reshape(-1, k)
    vals = job.out_vals.cpu().numpy().reshape(-1, k)
    counts = job.out_counts.cpu().numpy()
    for i in rows:
        rc, rv = _row_reference(job, int(targets[i]), k)
        n = int(counts[i])
        assert n == len(rc), f"{what}: row {targets[i]} has {n} entries, restatement {len(rc)}"
        gv, gc = vals[i, :n], cols[i, :n]
        assert np.all(np.diff(gv) <= 0), f"{what}: row {targets[i]} not best-first"
        np.testing.assert_allclose(gv, rv, rtol=rtol, err_msg=f"{what}: row {targets[i]} values")
        if n:
            band = rv[-1] * (1 + 10 * rtol)  # entries tied with the k-th value may legitimately differ
            assert set(gc[gv > band * (1 + 10 * rtol)].tolist()) <= set(rc.tolist()), f"{what}: row {targets[i]} columns"
            assert set(rc[rv > band * (1 + 10 * rtol)].tolist()) <= set(gc.tolist()), f"{what}: row {targets[i]} columns"


def _oracle_rows(job, picks, k, what, rtol=1e-5):
    """Rows `picks` of the job against the ORACLE's kernel (oracle/spy_oracle.c, the restatement of s_plus.h:265-453 that is
    pinned against the compiled reference) on the job's own operands: A cut to the picked target rows, B to the rows those
    reference (every other row empty), the row / column vectors as the device pre-processing left them."""
    import torch
    import scipy.sparse as sp
    from oracle import oracle
    from parity import assert_topk_parity
    torch.cuda.synchronize()
    A, B, P, v = job.A, job.B, job.params, job.vectors
    targets = job.targets.cpu().numpy()
    ts = [int(targets[i]) for i in picks]
    aip = A.indptr.cpu().numpy().astype(np.int64)
    a_parts = [(A.indices[aip[t]:aip[t + 1]].cpu().numpy(), A.data[aip[t]:aip[t + 1]].cpu().numpy()) for t in ts]
    a_indptr = np.concatenate([[0], np.cumsum([len(x[0]) for x in a_parts])]).astype(np.int32)
    a_indices = np.concatenate([x[0] for x in a_parts]).astype(np.int32)
    a_data = np.concatenate([x[1] for x in a_parts]).astype(np.float32)
    used = torch.unique(torch.from_numpy(a_indices).to(B.indptr.device).long())
    starts, ends = B.indptr[used].long(), B.indptr[used + 1].long()
    lens = ends - starts
    pos = torch.arange(int(lens.sum()), device=used.device) - torch.repeat_interleave(torch.cumsum(lens, 0) - lens, lens)
    q = torch.repeat_interleave(starts, lens) + pos
    b_len = np.zeros(B.n_rows, dtype=np.int64)
    b_len[used.cpu().numpy()] = lens.cpu().numpy()
    b_indptr = np.concatenate([[0], np.cumsum(b_len)]).astype(np.int32)
    b = (B.data[q].cpu().numpy().astype(np.float32), B.indices[q].cpu().numpy().astype(np.int32), b_indptr)
    e = np.array([], dtype=np.float32)
    row_vec = lambda n: v[n][torch.tensor(ts, device=v[n].device)].cpu().numpy().astype(np.float32) if v.get(n) is not None else e
    col_vec = lambda n: v[n].cpu().numpy().astype(np.float32) if v.get(n) is not None else e
    zi = np.zeros(1, dtype=np.int32)
    rows, cols, vals = oracle.knn_kernel(
        np.arange(len(ts), dtype=np.int32), (a_data, a_indices, a_indptr), b,
        row_vec("Xt"), col_vec("Yt"), row_vec("Xc"), col_vec("Yc"), row_vec("Xd"), col_vec("Yd"),
        P["a1"], P["l1"], P["l2"], P["l3"], P["t1"], P["t2"], P["stabilized_shrink"], P["bayesian_shrink"], P["threshold"],
        k, job.n_cols, 0, zi, zi, 0, zi, zi)
    ref = oracle.slab_to_csr(rows, cols, vals, len(ts), job.n_cols)
    gc = job.out_cols.cpu().numpy().reshape(-1, k)[picks]
    gv = job.out_vals.cpu().numpy().reshape(-1, k)[picks]
    gn = job.out_counts.cpu().numpy()[picks]
    rr = np.concatenate([np.full(int(n), i) for i, n in enumerate(gn)])
    keep = np.concatenate([np.arange(k) < n for n in gn])
    got = sp.csr_array((gv.ravel()[keep], (rr, gc.ravel()[keep])), shape=(len(ts), job.n_cols))
    assert_topk_parity(ref, got, k=k, rtol=rtol, what=f"{what} oracle kernel rows")


def _job(matrix1, matrix2, k, target_rows, tuning=None, **kw):
    from similaripy_b200 import _engine
    job = _engine.prepare_job(matrix1, matrix2, k=k, target_rows=target_rows, verbose=False, device=0, tuning=tuning, **kw)
    job.run()
    return job


def test_cfg2_cosine_item_item_full_size():
    """configs[1]: cosine(URM.T, k=100), URM 1M x 200k d=1e-3, BM25-normalised -- the bench workload, all 200k rows."""
    import scipy.sparse as sp
    import similaripy_b200 as sim
    from oracle import oracle
    from parity import assert_topk_parity
    urm = sim.bm25(_gen(1_000_000, 200_000, 1e-3, 2), inplace=True)
    k = 100
    job = _job(urm.T, None, k, None, l2=1.0, c1=0.5, c2=0.5)
    counts = job.out_counts.cpu().numpy()
    assert counts.min() == k and counts.max() == k  # ~127k candidates per row
    _check_rows(job, [0, 1, 77_777, 199_999], k, "cfg2")
    # every row's best neighbour is the row itself with similarity 1
    cols = job.out_cols.view(-1, k)[:, 0].cpu().numpy()
    vals = job.out_vals.view(-1, k)[:, 0].cpu().numpy()
    np.testing.assert_array_equal(cols, np.arange(200_000))
    np.testing.assert_allclose(vals, 1.0, rtol=1e-5)
    # symmetry: s(i, j) == s(j, i) wherever both directions were kept
    res = sim.to_host(job.to_device_matrix())
    rt = res.T.tocsr()
    both = res.multiply(rt > 0)
    assert abs(both - rt.multiply(res > 0)).max() <= 1e-5
    # spot rows against the CPU oracle (host copies of the operands)
    s = urm.stored
    host = sp.csr_array((s.data.cpu().numpy(), s.indices.cpu().numpy(), s.indptr.cpu().numpy()), shape=urm.shape)
    rows = np.array([3, 4242, 123_456], dtype=np.int32)
    ref = oracle.similarity("cosine", host.T.tocsr(), host, k=k, target_rows=rows, format_output="csr", verbose=False)
    sub = res[rows].tocoo()
    chk = sp.csr_array((sub.data, (rows[sub.row], sub.col)), shape=res.shape)
    assert_topk_parity(ref, chk, k=k, rtol=1e-5, what="cfg2 oracle spot rows")


def test_cfg2_jaccard_binary_full_size_counts_with_integer_adds():
    """binary=True on the configs[1] URM (Jaccard item-item, k=100, 20 000 target rows): the stream engine counts the products
    with native integer adds (csrc/knn_stream_kernel.cuh, UNIT); against the float64 restatement, the oracle's kernel, and
    the float adds of the same engine (identical values: sums of ones are exact either way)."""
    import torch
    urm = _gen(1_000_000, 200_000, 1e-3, 2)
    k = 100
    rows = np.sort(np.random.default_rng(12).choice(200_000, size=20_000, replace=False)).astype(np.int32)
    kw = dict(l1=1.0, t1=1.0, t2=1.0, binary=True)
    job = _job(urm.T, None, k, rows, **kw)
    assert job.out_counts.cpu().numpy().min() == k
    _check_rows(job, [0, 9_999, 19_999], k, "cfg2 jaccard binary")
    _oracle_rows(job, [5, 10_000, 19_998], k, "cfg2 jaccard binary")
    if int(job.args.engine) == 2:
        assert int(job.args.unit_values) == 1
        flt = _job(urm.T, None, k, rows, tuning=dict(unit_values=False), **kw)
        assert int(flt.args.unit_values) == 0
        assert torch.equal(flt.out_counts, job.out_counts)
        assert torch.equal(torch.sort(flt.out_vals.view(-1, k), dim=1).values, torch.sort(job.out_vals.view(-1, k), dim=1).values)


def test_cfg3_s_plus_full_matrix_size():
    """configs[2]: s_plus(X, k=200, shrink=10), X 500k x 500k d=2e-3 (5e8 nnz, 10 column panels); 600 target rows."""
    x = _gen(500_000, 500_000, 2e-3, 3)
    rows = np.sort(np.random.default_rng(3).choice(500_000, size=600, replace=False)).astype(np.int32)
    kw = dict(l1=0.5, l2=0.5, l3=0.0, t1=1.0, t2=1.0, c1=0.5, c2=0.5, stabilized_shrink=10.0)
    k = 200
    job = _job(x, None, k, rows, **kw)
    assert int(job.args.n_panels) >= 9
    assert job.out_counts.cpu().numpy().min() == k
    _check_rows(job, [0, 299, 599], k, "cfg3")
    _oracle_rows(job, [1, 300, 598], k, "cfg3")
    # a different launch plan (2 CTAs per SM, narrower panels, other group width) selects the same neighbours
    job2 = _job(x, None, k, rows, tuning=dict(threads=512, panel_width=12_800, group=8), **kw)
    a = job.out_vals.view(-1, k)
    b = job2.out_vals.view(-1, k)
    assert float(((a - b).abs() / a.abs().clamp_min(1e-30)).max()) <= 1e-5
    same = (job.out_cols.view(-1, k) == job2.out_cols.view(-1, k)).float().mean()
    assert float(same) > 0.999  # apart from ties / last-ulp swaps


def test_cfg4_rp3beta_full_matrix_size():
    """configs[3]: rp3beta(URM.T, alpha=1, beta=0.6, k=100), URM 2M x 500k d=5e-4; 600 of the 500k target rows."""
    import similaripy_b200 as sim
    from similaripy_b200 import _engine
    urm = _gen(2_000_000, 500_000, 5e-4, 4)
    rows = np.sort(np.random.default_rng(4).choice(500_000, size=600, replace=False)).astype(np.int32)
    k = 100
    # the public function's pre-processing on the device (similarity.py:477-483), then the kernel-level job
    m1 = urm.T
    pop = _engine.axis_sum(urm, 0)                       # matrix2 = matrix1.T = URM; pop_m2 = column sums of URM
    m1n = sim.normalize(m1, norm="l1", axis=1)
    m2n = sim.normalize(urm, norm="l1", axis=1)
    job = _job(m1n, m2n, k, rows, weight_depop_matrix2=pop, p2=0.6, l3=1.0)
    assert job.out_counts.cpu().numpy().min() == k
    _check_rows(job, [0, 300, 599], k, "cfg4")
    _oracle_rows(job, [2, 301, 597], k, "cfg4")
    # and the public entry point gives the same rows
    res = sim.rp3beta(urm.T, alpha=1.0, beta=0.6, k=k, target_rows=rows, verbose=False, on_device=True)
    host = sim.to_host(res)
    vals = job.out_vals.view(-1, k).cpu().numpy()
    for i in (0, 300, 599):
        r = rows[i]
        np.testing.assert_allclose(np.sort(host.data[host.indptr[r]:host.indptr[r + 1]])[::-1], vals[i], rtol=1e-5)


def test_cfg5_recommend_with_filter_full_matrix_size():
    """configs[4]: dot_product(URM, S.T, k=100, filter_cols=URM), URM 5M x 200k d=1e-3 (1e9 nnz); 2000 target users.
    S is a synthetic 200k x 200k item model with ~100 neighbours per row."""
    import torch
    import similaripy_b200 as sim
    urm = _gen(5_000_000, 200_000, 1e-3, 5)
    s_t = _gen(200_000, 200_000, 5e-4, 55)  # plays S.T: rows = items seen, columns = items recommended
    rows = np.sort(np.random.default_rng(5).choice(5_000_000, size=2000, replace=False)).astype(np.int32)
    k = 100
    res = sim.dot_product(urm, s_t, k=k, target_rows=rows, filter_cols=urm, verbose=False, format_output="csr")
    assert res.shape == (5_000_000, 200_000)
    nz_rows = np.flatnonzero(np.diff(res.indptr))
    assert set(nz_rows.tolist()) <= set(rows.tolist())
    # nothing the user already has is recommended (s_plus.h:159-172)
    ip = urm.stored.indptr.cpu().numpy()
    ix = urm.stored.indices
    for r in rows[:200]:
        seen = set(ix[int(ip[r]):int(ip[r + 1])].cpu().numpy().tolist())
        assert not (seen & set(res.indices[res.indptr[r]:res.indptr[r + 1]].tolist()))
    # unfiltered kernel rows against the float64 restatement, then the filter removes exactly the seen items
    job = _job(urm, s_t, k, rows)
    _check_rows(job, [0, 999, 1999], k, "cfg5")
    _oracle_rows(job, [1, 1000, 1998], k, "cfg5")
    for i in (0, 999, 1999):
        r = int(rows[i])
        rc, rv = _row_reference(job, r, 20_000)
        seen = set(ix[int(ip[r]):int(ip[r + 1])].cpu().numpy().tolist())
        keep = [j for j, c in enumerate(rc) if c not in seen][:k]
        got = np.sort(res.data[res.indptr[r]:res.indptr[r + 1]])[::-1]
        np.testing.assert_allclose(got, rv[keep], rtol=1e-5)
    del torch
