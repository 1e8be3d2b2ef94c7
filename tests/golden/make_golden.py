"""Generate the golden fixtures tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference, compiled into oracle/_ref by oracle/build_ref.py):

    python tests/golden/make_golden.py

Every fixture stores the seeded inputs (CSR arrays), the call (function name + JSON kwargs) and what the
reference returned: for similarities the raw COO slab (rows, cols, values: n_targets*k entries in the
reference's heap order, num_threads=1 is irrelevant to the result) -- for normalizers the output data array.
The reference cannot travel to the GPU box, these vectors can.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_api  # noqa: E402


def rand_csr(n_rows, n_cols, density, seed, kind="float"):
    rng = np.random.default_rng(seed)
    m = sp.random_array((n_rows, n_cols), density=density, format="csr", dtype=np.float32, random_state=rng)
    if kind == "int":      # integer-valued: products and sums are exact in fp32
        m.data = np.floor(m.data * 5).astype(np.float32) + 1.0
    elif kind == "signed":  # mixed signs: exercises threshold < 0 and zero-sum candidates
        m.data = (m.data - 0.5).astype(np.float32)
    m.sort_indices()
    return m


def pack(m, prefix):
    return {f"{prefix}_data": m.data, f"{prefix}_indices": m.indices.astype(np.int32),
            f"{prefix}_indptr": m.indptr.astype(np.int32), f"{prefix}_shape": np.asarray(m.shape, dtype=np.int64)}


SIM_CASES = []


def sim_case(name, fn, m_kind="float", shape=(120, 90), density=0.08, seed=1, m2_shape=None, **kw):
    SIM_CASES.append(dict(name=name, fn=fn, m_kind=m_kind, shape=shape, density=density, seed=seed,
                          m2_shape=m2_shape, kw=kw))


# the nine public functions, top-k truncating
sim_case("dot_topk", "dot_product", k=10)
sim_case("cosine_topk", "cosine", k=10)
sim_case("asym_topk", "asymmetric_cosine", alpha=0.2, k=10)
sim_case("jaccard_topk", "jaccard", k=10)
sim_case("dice_topk", "dice", k=10)
sim_case("tversky_topk", "tversky", alpha=0.8, beta=0.4, k=10)
sim_case("p3alpha_topk", "p3alpha", alpha=0.8, k=10)
sim_case("rp3beta_topk", "rp3beta", alpha=0.8, beta=0.4, k=10)
sim_case("splus_topk", "s_plus", l1=0.5, l2=0.5, l3=1.0, t1=0.7, t2=0.3, c1=0.4, c2=0.6, alpha=1.0, beta1=0.2,
         beta2=0.6, pop1="sum", pop2="sum", k=10)
# full rows (k = n_cols: no truncation)
sim_case("cosine_full", "cosine", k=1000)
sim_case("splus_full", "s_plus", k=1000, shrink=3.0)
# shrink types
for st in ("stabilized", "bayesian", "additive"):
    sim_case(f"cosine_shrink_{st}", "cosine", k=10, shrink=10.0, shrink_type=st)
    sim_case(f"tversky_shrink_{st}", "tversky", alpha=0.5, beta=0.5, k=10, shrink=2.0, shrink_type=st)
# integer data: bit-exact
sim_case("dot_int", "dot_product", m_kind="int", k=10)
sim_case("jaccard_binary", "jaccard", k=10, binary=True)
sim_case("cosine_binary", "cosine", k=10, binary=True)
# row / column selectors
sim_case("cosine_target_rows", "cosine", k=10, target_rows=[3, 17, 5, 99, 100, 42])
sim_case("dot_filter_list", "dot_product", k=10, filter_cols=[0, 1, 2, 50, 119, 5000])
sim_case("dot_target_list", "dot_product", k=10, target_cols=[4, 8, 15, 16, 23, 42, 77, 118])
sim_case("dot_target_and_filter_list", "dot_product", k=10, target_cols=list(range(0, 120, 3)), filter_cols=[3, 6, 9])
sim_case("cosine_binary_filter_list", "cosine", k=10, binary=True, filter_cols=[1, 2, 3])  # the a11 quirk
sim_case("dot_filter_matrix", "dot_product", k=10, filter_cols="matrix:7")
sim_case("dot_target_matrix", "dot_product", k=10, target_cols="matrix:8")
sim_case("cosine_filter_and_target_matrix", "cosine", k=5, filter_cols="matrix:9", target_cols="matrix:10")
# threshold, a1, signed data
sim_case("cosine_threshold", "cosine", k=10, threshold=0.15)
sim_case("dot_signed_negative_threshold", "dot_product", m_kind="signed", k=10, threshold=-0.05)
sim_case("splus_alpha_pow", "s_plus", l1=0.0, l2=1.0, alpha=1.7, k=10)
# rectangular two-matrix call (README flow: URM x S^T with the URM as filter)
sim_case("dot_two_matrices", "dot_product", shape=(60, 90), m2_shape=(90, 70), k=8)
sim_case("dot_recommend_filter_urm", "dot_product", shape=(60, 90), m2_shape=(90, 90), k=8, filter_cols="matrix1",
         target_rows=[1, 2, 3, 59])
# the reference's blocked / popularity-permuted path
for bs in (None, 32, 64):
    sim_case(f"cosine_block_{bs}", "cosine", k=10, block_size=bs)
    sim_case(f"jaccard_binary_block_{bs}", "jaccard", k=10, binary=True, block_size=bs)


def resolve(spec, m1, out_shape, base_seed):
    if isinstance(spec, str) and spec.startswith("matrix:"):
        return rand_csr(out_shape[0], out_shape[1], 0.2, int(spec.split(":")[1]))
    if spec == "matrix1":
        return m1
    return spec


def make_similarity():
    for c in SIM_CASES:
        m1 = rand_csr(*c["shape"], c["density"], c["seed"], c["m_kind"])
        m2 = rand_csr(*c["m2_shape"], c["density"], c["seed"] + 100, c["m_kind"]) if c["m2_shape"] else None
        out_shape = (m1.shape[0], m2.shape[1] if m2 is not None else m1.shape[0])
        kw = dict(c["kw"])
        arrays = {}
        for sel in ("filter_cols", "target_cols"):
            if sel in kw:
                v = resolve(kw[sel], m1, out_shape, c["seed"])
                if sp.issparse(v):
                    arrays.update(pack(v.tocsr(), sel))
                    kw[sel] = "@matrix"
                    c.setdefault("mats", {})[sel] = v
        call_kw = {k: (c["mats"][k] if v == "@matrix" else v) for k, v in kw.items()} if "mats" in c else dict(kw)
        res = ref_api.similarity(c["fn"], m1.copy(), None if m2 is None else m2.copy(), format_output="coo",
                                 num_threads=1, **call_kw)
        arrays.update(pack(m1, "m1"))
        if m2 is not None:
            arrays.update(pack(m2, "m2"))
        np.savez_compressed(os.path.join(HERE, f"sim_{c['name']}.npz"), fn=c["fn"], kwargs=json.dumps(kw),
                            out_rows=res.row.astype(np.int32), out_cols=res.col.astype(np.int32),
                            out_vals=res.data.astype(np.float32), out_shape=np.asarray(res.shape, dtype=np.int64),
                            **arrays)
        print("sim", c["name"], res.shape, int((res.data != 0).sum()))


NORM_CASES = [("l1", "normalize", dict(norm="l1")), ("l2", "normalize", dict(norm="l2")), ("max", "normalize", dict(norm="max")),
              ("l2_axis0", "normalize", dict(norm="l2", axis=0)),
              ("bm25", "bm25", {}), ("bm25_axis0", "bm25", dict(axis=0)), ("bm25_k1b", "bm25", dict(k1=1.6, b=0.5)),
              ("bm25plus", "bm25plus", {}), ("bm25plus_delta", "bm25plus", dict(delta=0.5, logbase=2.0)),
              ("tfidf", "tfidf", {})]
for tf in ("binary", "raw", "sqrt", "freq", "log"):
    NORM_CASES.append((f"tfidf_tf_{tf}", "tfidf", dict(tf_mode=tf, idf_mode="smooth")))
for idf in ("unary", "base", "smooth", "prob", "bm25"):
    NORM_CASES.append((f"tfidf_idf_{idf}", "tfidf", dict(tf_mode="sqrt", idf_mode=idf, logbase=10.0)))
    NORM_CASES.append((f"bm25_idf_{idf}", "bm25", dict(idf_mode=idf, tf_mode="log")))


def make_normalization():
    for name, fn, kw in NORM_CASES:
        for dt in (np.float32, np.float64):
            for it in (np.int32, np.int64):
                if it == np.int64 and name not in ("l2", "bm25", "tfidf"):
                    continue
                m = rand_csr(150, 80, 0.07, 11)
                m.data = (m.data * 4).astype(dt)
                m = sp.csr_array((m.data, m.indices.astype(it), m.indptr.astype(it)), shape=m.shape)
                res = getattr(ref_api, fn)(m.copy(), **kw)
                tag = f"{name}_{np.dtype(dt).name}_{np.dtype(it).name}"
                np.savez_compressed(os.path.join(HERE, f"norm_{tag}.npz"), fn=fn, kwargs=json.dumps(kw),
                                    in_data=m.data, in_indices=m.indices, in_indptr=m.indptr,
                                    shape=np.asarray(m.shape, dtype=np.int64),
                                    out_data=res.data, out_indices=res.indices, out_indptr=res.indptr)
                print("norm", tag)


if __name__ == "__main__":
    assert ref_api.available(), "build oracle/_ref first: python oracle/build_ref.py"
    for f in os.listdir(HERE):
        if f.endswith(".npz"):
            os.remove(os.path.join(HERE, f))
    make_similarity()
    make_normalization()
