"""GPU: the reference's OWN unit tests (tests/reference_fixtures/, unmodified copies of the upstream
tests/test_similarity.py:289-617 and tests/test_normalization.py:12-96) run against similaripy_b200 with
``import similaripy`` resolved to this package -- the drop-in claim of SURVEY.md 8b, checked literally."""
import contextlib
import importlib.util
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = os.path.join(HERE, "reference_fixtures")
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("spy_engine")]


@contextlib.contextmanager
def _aliased():
    """`import similaripy` (also inside the fixtures' test bodies, e.g. test_example_code) resolves to similaripy_b200."""
    import similaripy_b200
    import similaripy_b200.cython_code
    import similaripy_b200.cython_code.utils
    import similaripy_b200.normalization
    import similaripy_b200.similarity
    alias = {"similaripy": similaripy_b200, "similaripy.normalization": similaripy_b200.normalization,
             "similaripy.similarity": similaripy_b200.similarity, "similaripy.cython_code": similaripy_b200.cython_code,
             "similaripy.cython_code.utils": similaripy_b200.cython_code.utils}
    saved = {k: sys.modules.get(k) for k in alias}
    sys.modules.update(alias)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _load(name):
    with _aliased():
        spec = importlib.util.spec_from_file_location(f"_ref_fixture_{name}", os.path.join(FIXTURES, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    return mod


def _cases():
    out = []
    for name in ("ref_test_similarity", "ref_test_normalization"):
        text = open(os.path.join(FIXTURES, name + ".py")).read()
        for line in text.splitlines():
            if line.startswith("def test_"):
                out.append((name, line[4:line.index("(")]))
    return out


_MODULES = {}


@pytest.mark.parametrize("module,test", _cases(), ids=lambda v: v)
def test_reference_test(module, test):
    if module not in _MODULES:
        _MODULES[module] = _load(module)
    with _aliased():
        getattr(_MODULES[module], test)()


def test_all_sixteen_reference_tests_are_covered():
    assert len(_cases()) == 16
