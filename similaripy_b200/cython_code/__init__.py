"""Compatibility shim for callers that reach into ``similaripy.cython_code`` (the reference's tests do,
tests/test_similarity.py:384-390).  Nothing here is Cython; the names map onto the CUDA library."""
from . import utils  # noqa: F401
