"""CPU: the drop-in boundary.  The C-ABI library loads and exports every symbol include/similaripy_b200.h
declares; the ctypes struct mirrors the C struct; the product never touches the oracle; without a GPU the
product fails loudly instead of falling back; the reference's Python-side validation errors are reproduced
before any device work."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "similaripy_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spy_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from similaripy_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(declared)
    assert lib.spy_abi_version() == _lib.ABI_VERSION == 8


def test_struct_layout_matches_c(tmp_path):
    """sizeof / offsetof of spy_knn_args as gcc sees the header == the ctypes mirror."""
    from similaripy_b200 import _lib
    fields = [f[0] for f in _lib.KnnArgs._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){",
             'printf("%zu\\n", sizeof(spy_knn_args));']
    lines += [f'printf("%zu\\n", offsetof(spy_knn_args, {f}));' for f in fields]
    lines += ["return 0;}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(_lib.KnnArgs)
    for name, off in zip(fields, out[1:]):
        assert getattr(_lib.KnnArgs, name).offset == off, name


def test_no_compute_needed_entry_points():
    from similaripy_b200 import _lib
    lib = _lib.load()
    assert lib.spy_scan_tmp_bytes(10_000) > 0
    assert lib.spy_tfidf_scratch_bytes(100, 50, _lib.F32) > 0
    a = _lib.KnnArgs()
    a.k, a.n_cols, a.n_targets = 100, 200_000, 1000
    assert lib.spy_knn_plan(ctypes.byref(a), -1) == 0  # device -1: plan with B200 defaults, no CUDA call
    assert a.n_panels >= 4 and a.panel_width % 128 == 0 and a.panel_width * a.n_panels >= a.n_cols
    a.k = 0
    assert lib.spy_knn_plan(ctypes.byref(a), -1) < 0
    assert b"k must be" in lib.spy_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "similaripy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "spy_oracle" not in text, f
    code = "import sys; import similaripy_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def _no_gpu():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a host without CUDA devices")
def test_fails_loudly_without_gpu():
    import similaripy_b200 as sim
    m = sp.random_array((30, 20), density=0.2, format="csr", dtype=np.float32, random_state=np.random.default_rng(0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sim.cosine(m, k=5, verbose=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sim.normalize(m)


def test_validation_errors_match_the_reference_conventions():
    """s_plus_utils.pyx:19-125 / similarity.py:615-616 / normalization.py:76-113: types and ordering."""
    import similaripy_b200 as sim
    m = sp.random_array((30, 20), density=0.2, format="csr", dtype=np.float32, random_state=np.random.default_rng(0))
    with pytest.raises(TypeError):
        sim.cosine(np.zeros((3, 3)), k=5)
    with pytest.raises(ValueError):
        sim.cosine(m, m, k=5)                       # 30x20 @ 30x20: incompatible
    with pytest.raises(ValueError):
        sim.cosine(m, k=0)
    with pytest.raises(ValueError):
        sim.cosine(m, k=5, shrink_type="nope")
    with pytest.raises(ValueError):
        sim.cosine(m, k=5, target_rows=list(range(31)))
    with pytest.raises(TypeError):
        sim.cosine(m, k=5, filter_cols="abc")
    with pytest.raises(ValueError):
        sim.cosine(m, k=5, filter_cols=sp.csr_array(np.ones((3, 3), dtype=np.float32)))
    with pytest.raises(TypeError):
        sim.cosine(m, k=5, verbose=1)
    with pytest.raises(ValueError):
        sim.cosine(m, k=5, verbose=False, format_output="dense")
    with pytest.raises(ValueError):
        sim.s_plus(m, k=5, pop1="bogus", verbose=False)
    with pytest.raises(ValueError):
        sim.normalize(m, norm="l3")
    with pytest.raises(ValueError):
        sim.normalize(m, axis=2)
    with pytest.raises(TypeError):
        sim.normalize(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        sim.tfidf(m, tf_mode="x")
    with pytest.raises(ValueError):
        sim.bm25(m, idf_mode="x")
