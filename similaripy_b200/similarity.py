"""Public similarity functions -- drop-in for ``similaripy.similarity`` (reference
similaripy/similarity.py:9-617).  Same names, keyword sets, defaults, return types and
exceptions; each function only picks the constants of the one fused kernel behind
``_engine.s_plus`` (the B200 replacement of ``cython_code.s_plus.s_plus``).

Three keyword-only extras exist on every function and default to the reference behaviour:
``device`` (CUDA device index, default: current device / $SIMILARIPY_B200_DEVICE),
``tuning`` (dict: threads / group / panel_width overrides for the launch plan) and
``on_device`` (return a ``DeviceMatrix`` that stays in HBM instead of a scipy matrix; inputs may be
``DeviceMatrix`` handles from ``similaripy_b200.to_device`` as well, so calls chain without PCIe trips).
"""
from __future__ import annotations

from typing import Literal, Optional, Union

import numpy as np
from scipy.sparse import sparray

from . import _engine as _sim
from .normalization import normalize as _normalize

_Rows = Optional[Union[list, np.ndarray]]
_Cols = Optional[Union[list, np.ndarray, sparray]]


def _shrink_values(shrink: float, shrink_type: str):
    """similarity.py:595-617 -> (stabilized, bayesian, additive)."""
    if shrink_type == "stabilized":
        return shrink, 0.0, 0.0
    if shrink_type == "bayesian":
        return 0.0, shrink, 0.0
    if shrink_type == "additive":
        return 0.0, 0.0, shrink
    raise ValueError("shrink_type must be one of 'stabilized', 'bayesian', or 'additive'")


__get_shrink_values__ = _shrink_values


def _call(matrix1, matrix2, shrink, shrink_type, common, **preset):
    stab, bayes, add = _shrink_values(shrink, shrink_type)
    return _sim.s_plus(matrix1, matrix2=matrix2, stabilized_shrink=stab, bayesian_shrink=bayes,
                       additive_shrink=add, **preset, **common)


def _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
            num_threads, block_size, device, tuning, on_device=False):
    return dict(k=k, threshold=threshold, binary=binary, target_rows=target_rows, target_cols=target_cols,
                filter_cols=filter_cols, verbose=verbose, format_output=format_output,
                num_threads=num_threads, block_size=block_size, device=device, tuning=tuning, on_device=on_device)


def dot_product(matrix1: sparray, matrix2: Optional[sparray] = None, k: int = 100, shrink: float = 0.0,
                shrink_type: Literal["stabilized", "bayesian", "additive"] = "stabilized", threshold: float = 0.0,
                binary: bool = False, target_rows: _Rows = None, target_cols: _Cols = None, filter_cols: _Cols = None,
                verbose: bool = True, format_output: Literal["csr", "coo"] = "coo", num_threads: int = 0,
                block_size: Optional[int] = 0, *, device=None, tuning=None, on_device=False) -> sparray:
    """Top-k dot product between rows of matrix1 and columns of matrix2 (similarity.py:9-64)."""
    return _call(matrix1, matrix2, shrink, shrink_type,
                 _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
                         num_threads, block_size, device, tuning, on_device))


def cosine(matrix1: sparray, matrix2: Optional[sparray] = None, k: int = 100, shrink: float = 0.0,
           shrink_type: Literal["stabilized", "bayesian", "additive"] = "stabilized", threshold: float = 0.0,
           binary: bool = False, target_rows: _Rows = None, target_cols: _Cols = None, filter_cols: _Cols = None,
           verbose: bool = True, format_output: Literal["csr", "coo"] = "coo", num_threads: int = 0,
           block_size: Optional[int] = 0, *, device=None, tuning=None, on_device=False) -> sparray:
    """Top-k cosine similarity (similarity.py:67-123): l2=1, c1=c2=0.5."""
    return _call(matrix1, matrix2, shrink, shrink_type,
                 _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
                         num_threads, block_size, device, tuning, on_device), l2=1, c1=0.5, c2=0.5)


def asymmetric_cosine(matrix1: sparray, matrix2: Optional[sparray] = None, alpha: float = 0.5, k: int = 100,
                      shrink: float = 0.0, shrink_type: Literal["stabilized", "bayesian", "additive"] = "stabilized",
                      threshold: float = 0.0, binary: bool = False, target_rows: _Rows = None,
                      target_cols: _Cols = None, filter_cols: _Cols = None, verbose: bool = True,
                      format_output: Literal["csr", "coo"] = "coo", num_threads: int = 0,
                      block_size: Optional[int] = 0, *, device=None, tuning=None, on_device=False) -> sparray:
    """Top-k asymmetric cosine (similarity.py:126-186): l2=1, c1=alpha, c2=1-alpha."""
    return _call(matrix1, matrix2, shrink, shrink_type,
                 _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
                         num_threads, block_size, device, tuning, on_device), l2=1, c1=alpha, c2=1 - alpha)


def tversky(matrix1: sparray, matrix2: Optional[sparray] = None, alpha: float = 1.0, beta: float = 1.0, k: int = 100,
            shrink: float = 0.0, shrink_type: Literal["stabilized", "bayesian", "additive"] = "stabilized",
            threshold: float = 0.0, binary: bool = False, target_rows: _Rows = None, target_cols: _Cols = None,
            filter_cols: _Cols = None, verbose: bool = True, format_output: Literal["csr", "coo"] = "coo",
            num_threads: int = 0, block_size: Optional[int] = 0, *, device=None, tuning=None, on_device=False) -> sparray:
    """Top-k Tversky similarity (similarity.py:189-249): l1=1, t1=alpha, t2=beta."""
    return _call(matrix1, matrix2, shrink, shrink_type,
                 _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
                         num_threads, block_size, device, tuning, on_device), l1=1, t1=alpha, t2=beta)


def jaccard(matrix1: sparray, matrix2: Optional[sparray] = None, k: int = 100, shrink: float = 0.0,
            shrink_type: Literal["stabilized", "bayesian", "additive"] = "stabilized", threshold: float = 0.0,
            binary: bool = False, target_rows: _Rows = None, target_cols: _Cols = None, filter_cols: _Cols = None,
            verbose: bool = True, format_output: Literal["csr", "coo"] = "coo", num_threads: int = 0,
            block_size: Optional[int] = 0, *, device=None, tuning=None, on_device=False) -> sparray:
    """Top-k Jaccard similarity (similarity.py:252-308): Tversky with t1=t2=1."""
    return _call(matrix1, matrix2, shrink, shrink_type,
                 _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
                         num_threads, block_size, device, tuning, on_device), l1=1, t1=1, t2=1)


def dice(matrix1: sparray, matrix2: Optional[sparray] = None, k: int = 100, shrink: float = 0.0,
         shrink_type: Literal["stabilized", "bayesian", "additive"] = "stabilized", threshold: float = 0.0,
         binary: bool = False, target_rows: _Rows = None, target_cols: _Cols = None, filter_cols: _Cols = None,
         verbose: bool = True, format_output: Literal["csr", "coo"] = "coo", num_threads: int = 0,
         block_size: Optional[int] = 0, *, device=None, tuning=None, on_device=False) -> sparray:
    """Top-k Dice similarity (similarity.py:311-367): Tversky with t1=t2=0.5."""
    return _call(matrix1, matrix2, shrink, shrink_type,
                 _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
                         num_threads, block_size, device, tuning, on_device), l1=1, t1=0.5, t2=0.5)


def _random_walk_inputs(matrix1, matrix2, alpha, device):
    """similarity.py:410-415 / 477-483: l1-normalise the rows of both operands, then data ** alpha."""
    if matrix2 is None:
        matrix2 = matrix1.T
    out = []
    for m in (matrix1, matrix2):
        m = _normalize(m, norm="l1", axis=1, inplace=False, device=device)
        if isinstance(m, _sim.DeviceMatrix):
            _sim.pow_values_(m, alpha)
        else:
            m.data = np.power(m.data, alpha)
        out.append(m)
    return out[0], out[1]


def p3alpha(matrix1: sparray, matrix2: Optional[sparray] = None, alpha: float = 1.0, k: int = 100,
            shrink: float = 0.0, shrink_type: Literal["stabilized", "bayesian", "additive"] = "stabilized",
            threshold: float = 0.0, binary: bool = False, target_rows: _Rows = None, target_cols: _Cols = None,
            filter_cols: _Cols = None, verbose: bool = True, format_output: Literal["csr", "coo"] = "coo",
            num_threads: int = 0, block_size: Optional[int] = 0, *, device=None, tuning=None, on_device=False) -> sparray:
    """Top-k P3alpha: normalised 3-step random walk (similarity.py:370-432)."""
    m1, m2 = _random_walk_inputs(matrix1, matrix2, alpha, device)
    return _call(m1, m2, shrink, shrink_type,
                 _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
                         num_threads, block_size, device, tuning, on_device))


def rp3beta(matrix1: sparray, matrix2: Optional[sparray] = None, alpha: float = 1.0, beta: float = 1.0, k: int = 100,
            shrink: float = 0.0, shrink_type: Literal["stabilized", "bayesian", "additive"] = "stabilized",
            threshold: float = 0.0, binary: bool = False, target_rows: _Rows = None, target_cols: _Cols = None,
            filter_cols: _Cols = None, verbose: bool = True, format_output: Literal["csr", "coo"] = "coo",
            num_threads: int = 0, block_size: Optional[int] = 0, *, device=None, tuning=None, on_device=False) -> sparray:
    """Top-k RP3beta: P3alpha with popularity penalisation (similarity.py:435-503)."""
    m2_in = matrix1.T if matrix2 is None else matrix2
    if isinstance(m2_in, _sim.DeviceMatrix):
        pop_m2 = _sim.axis_sum(m2_in, axis=0)
    else:
        pop_m2 = np.asarray(m2_in.sum(axis=0)).ravel()  # on the un-normalised matrix2, similarity.py:479
    m1, m2 = _random_walk_inputs(matrix1, m2_in, alpha, device)
    return _call(m1, m2, shrink, shrink_type,
                 _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
                         num_threads, block_size, device, tuning, on_device),
                 weight_depop_matrix2=pop_m2, p2=beta, l3=1)


def s_plus(matrix1: sparray, matrix2: Optional[sparray] = None, l1: float = 0.5, l2: float = 0.5, l3: float = 0.0,
           t1: float = 1.0, t2: float = 1.0, c1: float = 0.5, c2: float = 0.5,
           pop1: Optional[Union[Literal["none", "sum"], np.ndarray]] = "none",
           pop2: Optional[Union[Literal["none", "sum"], np.ndarray]] = "none",
           alpha: float = 1.0, beta1: float = 0.0, beta2: float = 0.0, k: int = 100, shrink: float = 0.0,
           shrink_type: Literal["stabilized", "bayesian", "additive"] = "stabilized", threshold: float = 0.0,
           binary: bool = False, target_rows: _Rows = None, target_cols: _Cols = None, filter_cols: _Cols = None,
           verbose: bool = True, format_output: Literal["csr", "coo"] = "coo", num_threads: int = 0,
           block_size: Optional[int] = 0, *, device=None, tuning=None, on_device=False) -> sparray:
    """Top-k S-Plus: Tversky + cosine + depopularisation under tunable weights (similarity.py:506-592)."""
    return _call(matrix1, matrix2, shrink, shrink_type,
                 _common(k, threshold, binary, target_rows, target_cols, filter_cols, verbose, format_output,
                         num_threads, block_size, device, tuning, on_device),
                 l1=l1, l2=l2, l3=l3, t1=t1, t2=t2, c1=c1, c2=c2, a1=alpha,
                 weight_depop_matrix1=pop1, weight_depop_matrix2=pop2, p1=beta1, p2=beta2)
