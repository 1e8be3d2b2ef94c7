"""Build the UNMODIFIED reference (bogliosimone/similaripy @ /root/reference) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under similaripy_b200/ may import this.

Recipe: for each Cython module of the reference
(similaripy/cython_code/{s_plus,s_plus_utils,normalization,utils}.pyx) run
``cython --cplus`` and ``g++ -O3 -ffast-math -std=c++17 -fopenmp -fPIC -shared``
(the reference's own flags, CMakeLists.txt:79-82,121-125).  The sources are
read where they lie under /root/reference; generated C++ goes to a temporary
directory and ONLY the compiled extension modules are written to
``oracle/_ref/similaripy_ref/cython_code/`` (git-ignored, but shipped to the GPU
box by gpurun).  No reference source file is copied into the repository.

On a machine without /root/reference (the GPU box) this script is a no-op and
the prebuilt extension modules are used as they are.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SIMILARIPY_REFERENCE", "/root/reference")
OUT_PKG = os.path.join(HERE, "_ref", "similaripy_ref")
MODULES = ("utils", "s_plus_utils", "normalization", "s_plus")


def ref_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "similaripy", "cython_code"))


def built() -> bool:
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(OUT_PKG, "cython_code", m + suffix)) for m in MODULES)


def build(force: bool = False, verbose: bool = True) -> bool:
    """Returns True when oracle/_ref is usable afterwards."""
    if built() and not force:
        return True
    if not ref_available():
        return built()
    import numpy as np

    src = os.path.join(REF_ROOT, "similaripy", "cython_code")
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    py_inc = sysconfig.get_paths()["include"]
    np_inc = np.get_include()
    out_dir = os.path.join(OUT_PKG, "cython_code")
    os.makedirs(out_dir, exist_ok=True)
    # package markers are generated (empty), not copied
    for d in (OUT_PKG, out_dir):
        open(os.path.join(d, "__init__.py"), "w").close()
    with tempfile.TemporaryDirectory(prefix="spy_refbuild_") as tmp:
        for m in MODULES:
            cpp = os.path.join(tmp, m + ".cpp")
            cmd = [sys.executable, "-m", "cython", os.path.join(src, m + ".pyx"), "--cplus", "-3",
                   "--output-file", cpp, "-I", src]
            if verbose:
                print("[oracle/_ref]", " ".join(cmd))
            subprocess.check_call(cmd, cwd=tmp)
            so = os.path.join(tmp, m + suffix)
            cmd = ["g++", "-O3", "-ffast-math", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-w",
                   "-I", py_inc, "-I", np_inc, "-I", src, cpp, "-o", so]
            if verbose:
                print("[oracle/_ref]", " ".join(cmd))
            subprocess.check_call(cmd, cwd=tmp)
            shutil.copy2(so, os.path.join(out_dir, m + suffix))
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built:", ok)
    sys.exit(0 if ok else 1)
