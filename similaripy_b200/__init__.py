"""similaripy_b200 -- B200-native (sm_100a) drop-in for the sparse-KNN similarity hot path of
bogliosimone/similaripy: the same public functions as ``similaripy/__init__.py:8-36``, computed by
hand-written CUDA kernels behind a C ABI (include/similaripy_b200.h).  No CPU fallback."""
__version__ = "0.1.0"

from . import normalization, sharded, similarity  # noqa: F401
from .normalization import bm25, bm25plus, normalize, tfidf
from .similarity import (asymmetric_cosine, cosine, dice, dot_product, jaccard, p3alpha, rp3beta, s_plus, tversky)
from ._engine import DeviceMatrix, to_device, to_host
from . import cython_code  # noqa: F401  (compat shim: sim.cython_code.utils.get_num_threads)

__all__ = [
    "__version__", "normalize", "bm25", "bm25plus", "tfidf", "dot_product", "cosine", "asymmetric_cosine",
    "jaccard", "dice", "tversky", "p3alpha", "rp3beta", "s_plus",
    "DeviceMatrix", "to_device", "to_host", "sharded",
]
