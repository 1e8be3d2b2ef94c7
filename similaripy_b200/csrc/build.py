"""Compile the CUDA sources of similaripy_b200 into one shared library for sm_100a.

    python -m similaripy_b200.csrc.build [--force]

nvcc cross-compiles without a GPU.  The library is written next to the package
(``similaripy_b200/libsimilaripy_b200.so``): it is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libsimilaripy_b200.so")
SOURCES = ("api.cu", "knn_kernel.cu", "knn_stream.cu", "knn_stream_sparse.cu", "knn_inst_g4.cu", "knn_inst_g8.cu", "knn_inst_g16.cu", "knn_inst_g32.cu",
           "csr_ops.cu", "normalize.cu", "host_api.cu", "knn_reforder.cu")
HEADERS = ("common.cuh", "knn_kernel.cuh", "knn_stream_kernel.cuh", "knn_stream_impl.inc", "knn_inst.inc", os.path.join("..", "..", "include", "similaripy_b200.h"))
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--use_fast_math=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the similaripy_b200 CUDA library cannot be built")


STAMP = LIB + ".stamp"  # hash of the sources the library was built from (mtimes do not survive the gpurun snapshot)


def source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for d in [os.path.join(HERE, s) for s in SOURCES + HEADERS]:
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False, extra_flags=(), out_path: str | None = None) -> str:
    """Default build: the product library.  extra_flags / out: experimental variants (e.g. -DSPY_UNROLL=4)."""
    variant = out_path is not None
    if not variant and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + list(extra_flags)
    objs = []
    build_dir = os.path.join(HERE, "build" if not variant else "build_" + os.path.basename(out_path))
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-Xptxas", "-v", "-c", os.path.join(HERE, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(build_dir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    target = out_path if variant else LIB
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target, *objs, "-lcudart"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    if not variant:
        with open(STAMP, "w") as f:
            f.write(source_hash())
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
