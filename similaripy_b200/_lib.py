"""ctypes binding of the C ABI declared in include/similaripy_b200.h.

The shared library is built in-tree by ``similaripy_b200.csrc.build`` (``__graft_entry__.build()``).
There is deliberately NO fallback: if the library is missing, or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SIMILARIPY_B200_LIB: alternative build of the same library (kernel experiments only)
LIB_PATH = os.environ.get("SIMILARIPY_B200_LIB") or os.path.join(_HERE, "libsimilaripy_b200.so")

SEL_NONE, SEL_ARRAY, SEL_MATRIX = 0, 1, 2
F32, F64 = 0, 1
I32, I64 = 0, 1
VAL_I32, VAL_I64 = 2, 3
TF_MODES = {"binary": 0, "raw": 1, "sqrt": 2, "freq": 3, "log": 4}
IDF_MODES = {"unary": 0, "base": 1, "smooth": 2, "prob": 3, "bm25": 4}

_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double


class KnnArgs(C.Structure):
    """Mirror of ``struct spy_knn_args``."""
    _fields_ = [
        ("n_targets", _i32), ("targets", _vp),
        ("a_rows", _i32), ("a_indptr", _vp), ("a_indices", _vp), ("a_data", _vp),
        ("b_rows", _i32), ("n_cols", _i32), ("b_indptr", _vp), ("b_indices", _vp), ("b_data", _vp),
        ("Xtversky", _vp), ("Ytversky", _vp), ("Xcosine", _vp), ("Ycosine", _vp), ("Xdepop", _vp), ("Ydepop", _vp),
        ("a1", _f32), ("l1", _f32), ("l2", _f32), ("l3", _f32), ("t1", _f32), ("t2", _f32),
        ("stabilized_shrink", _f32), ("bayesian_shrink", _f32), ("threshold", _f32),
        ("k", _i32),
        ("filter_mode", _i32), ("filter_indptr", _vp), ("filter_indices", _vp),
        ("target_mode", _i32), ("target_indptr", _vp), ("target_indices", _vp),
        ("out_rows", _vp), ("out_cols", _vp), ("out_values", _vp), ("out_counts", _vp),
        ("panel_width", _i32), ("b_split", _vp), ("split_stride", _i32), ("n_panels", _i32),
        ("threads", _i32), ("b_pairs", _vp), ("row_order", _vp),
        ("b_nnz", _i64), ("group", _i32),
        ("engine", _i32), ("b_chunk_indptr", _vp), ("b_chunks", _vp), ("toff", _vp), ("n_entries", _i64), ("aexp", _vp), ("a_nnz", _i64),
        ("unit_values", _i32),
    ]


# name -> (restype, argtypes); every symbol include/similaripy_b200.h declares
SIGNATURES = {
    "spy_abi_version": (C.c_int, []),
    "spy_device_count": (C.c_int, []),
    "spy_last_error": (C.c_char_p, []),
    "spy_launch_count": (_i64, [C.c_int]),
    "spy_knn_plan": (C.c_int, [C.POINTER(KnnArgs), C.c_int]),
    "spy_knn_pack_pairs_dev": (C.c_int, [_i64, _vp, _vp, _vp, _vp]),
    "spy_knn_scratch_bytes": (_i64, [C.POINTER(KnnArgs), C.c_int]),
    "spy_knn_build_split_dev": (C.c_int, [_i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "spy_knn_row_work_dev": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spy_knn_chunk_counts_dev": (C.c_int, [_i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "spy_knn_pad_chunks_dev": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32]),
    "spy_knn_row_lengths_dev": (C.c_int, [_i32, _vp, _vp, _vp, _vp]),
    "spy_knn_build_aexp_dev": (C.c_int, [C.POINTER(KnnArgs), _vp]),
    "spy_knn_topk_dev": (C.c_int, [C.POINTER(KnnArgs), _vp, _i64, _vp]),
    "spy_knn_topk_host": (C.c_int, [C.POINTER(KnnArgs), C.c_int]),
    "spy_knn_reforder_scratch_bytes": (_i64, [C.POINTER(KnnArgs)]),
    "spy_knn_topk_reforder_dev": (C.c_int, [C.POINTER(KnnArgs), _i32, _vp, _vp, _i64, _vp]),
    "spy_knn_topk_multi_host": (C.c_int, [C.POINTER(KnnArgs), _vp, _i32, _i32, _vp]),
    "spy_h2d_staged": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp]),
    "spy_normalize_rows_host": (C.c_int, [C.c_int, _i64, _vp, C.c_int, _vp, C.c_int, C.c_int]),
    "spy_tfidf_host": (C.c_int, [_i64, _i64, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, _f64, C.c_int]),
    "spy_bm25plus_host": (C.c_int, [_i64, _i64, _vp, C.c_int, _vp, _vp, C.c_int, _f64, _f64, _f64, C.c_int, C.c_int, _f64, C.c_int]),
    "spy_csr_row_sum_dev": (C.c_int, [_i32, _vp, _vp, C.c_int, _vp, _vp]),
    "spy_csr_col_sum_dev": (C.c_int, [_i64, _vp, _vp, C.c_int, _i32, _vp, _vp, _vp]),
    "spy_pow_shift_dev": (C.c_int, [_i64, _vp, C.c_int, _f32, _f32, _vp, _vp]),
    "spy_csr_col_count_dev": (C.c_int, [_i64, _vp, _i32, _vp, _vp]),
    "spy_scan_tmp_bytes": (_i64, [_i64]),
    "spy_exclusive_scan_i32_dev": (C.c_int, [_i64, _vp, _vp, _vp, _vp]),
    "spy_exclusive_scan_i64_dev": (C.c_int, [_i64, _vp, _vp, _vp, _vp]),
    "spy_csr_transpose_dev": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "spy_csr_sort_rows_dev": (C.c_int, [_i32, _vp, _vp, _vp, _vp]),
    "spy_csr_filter_count_dev": (C.c_int, [_i32, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "spy_csr_filter_compact_dev": (C.c_int, [_i32, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "spy_csr_wide_block_indptr_dev": (C.c_int, [_i64, _vp, _i64, _i64, _vp, _vp]),
    "spy_csr_indptr_add_dev": (C.c_int, [_i64, _vp, _vp, _vp]),
    "spy_slab_merge_dev": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spy_narrow_index_dev": (C.c_int, [_i64, _vp, _vp, _vp]),
    "spy_cast_values_dev": (C.c_int, [_i64, _vp, C.c_int, C.c_int, _vp, _vp]),
    "spy_slab_row_nnz_dev": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "spy_slab_compact_dev": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp]),
    "spy_slab_fill_rows_dev": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp]),
    "spy_normalize_rows_dev": (C.c_int, [C.c_int, _i64, _vp, C.c_int, _vp, C.c_int, _vp]),
    "spy_tfidf_scratch_bytes": (_i64, [_i64, _i64, C.c_int]),
    "spy_tfidf_dev": (C.c_int, [_i64, _i64, _vp, C.c_int, _vp, _vp, C.c_int, C.c_int, C.c_int, _f64, _vp, _vp]),
    "spy_bm25plus_dev": (C.c_int, [_i64, _i64, _vp, C.c_int, _vp, _vp, C.c_int, _f64, _f64, _f64,
                                   C.c_int, C.c_int, _f64, _vp, _vp]),
}

ABI_VERSION = 8
ENGINE_AUTO, ENGINE_FLAT, ENGINE_STREAM = 0, 1, 2
ERR_UNSUPPORTED = -4
ENGINES = {"auto": ENGINE_AUTO, "flat": ENGINE_FLAT, "stream": ENGINE_STREAM}
_lib = None


class SimilaripyB200Error(RuntimeError):
    """A C-ABI call returned a negative status."""


def load() -> C.CDLL:
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"similaripy_b200: CUDA library not found at {LIB_PATH}. Build it with "
            "`python -m similaripy_b200.csrc.build` (needs nvcc; cross-compiles sm_100a without a GPU). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.spy_abi_version() != ABI_VERSION:
        raise RuntimeError("similaripy_b200: ABI version mismatch between _lib.py and the shared library")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc < 0:
        msg = load().spy_last_error().decode("utf-8", "replace")
        raise SimilaripyB200Error(f"similaripy_b200 C-ABI call failed ({rc}): {msg}")


def device_count() -> int:
    return int(load().spy_device_count())


def launch_count(reset: bool = False) -> int:
    return int(load().spy_launch_count(1 if reset else 0))
