"""``similaripy.cython_code.utils`` look-alike (reference utils.pyx:18-40)."""
import os

import numpy as np

from .. import _lib


def get_num_threads() -> int:
    """The reference returns omp_get_max_threads(); the unit of parallelism here is the GPU.
    Returns the number of visible CUDA devices when there is one, else the host core count."""
    try:
        n = _lib.device_count()
    except Exception:
        n = 0
    return n if n >= 1 else (os.cpu_count() or 1)


def get_index_dtype(maxval: int):
    return np.int32 if maxval <= np.iinfo(np.int32).max else np.int64
