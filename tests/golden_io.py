"""Loader for the golden fixtures in tests/golden/ (generated from the unmodified reference by
tests/golden/make_golden.py).  Shared by the CPU (oracle) and GPU (CUDA path) golden tests."""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import scipy.sparse as sp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _csr(z, prefix):
    shape = tuple(int(x) for x in z[f"{prefix}_shape"])
    return sp.csr_array((z[f"{prefix}_data"], z[f"{prefix}_indices"], z[f"{prefix}_indptr"]), shape=shape)


def similarity_cases():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "sim_*.npz")))


def normalization_cases():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "norm_*.npz")))


def load_similarity(name):
    """-> dict(fn, kw, m1, m2, k, slab=(rows, cols, vals), shape, ref_csr)."""
    z = np.load(os.path.join(GOLDEN_DIR, f"sim_{name}.npz"), allow_pickle=False)
    kw = json.loads(str(z["kwargs"]))
    m1 = _csr(z, "m1")
    m2 = _csr(z, "m2") if "m2_data" in z.files else None
    for sel in ("filter_cols", "target_cols"):
        if kw.get(sel) == "@matrix":
            kw[sel] = _csr(z, sel)
    shape = tuple(int(x) for x in z["out_shape"])
    rows, cols, vals = z["out_rows"], z["out_cols"], z["out_vals"]
    n_targets = len(kw["target_rows"]) if kw.get("target_rows") is not None else shape[0]
    k = rows.shape[0] // max(n_targets, 1)
    ref = sp.coo_array((vals, (rows, cols)), shape=shape).tocsr()  # duplicates cannot occur: one slot per (row, col)
    ref.eliminate_zeros()
    return dict(fn=str(z["fn"]), kw=kw, m1=m1, m2=m2, k=k, slab=(rows, cols, vals), shape=shape, ref_csr=ref)


def load_normalization(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"norm_{name}.npz"), allow_pickle=False)
    shape = tuple(int(x) for x in z["shape"])
    m = sp.csr_array((z["in_data"], z["in_indices"], z["in_indptr"]), shape=shape)
    out = sp.csr_array((z["out_data"], z["out_indices"], z["out_indptr"]), shape=shape)
    return dict(fn=str(z["fn"]), kw=json.loads(str(z["kwargs"])), m=m, out=out)
