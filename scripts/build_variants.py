#!/usr/bin/env python
"""Build experimental variants of the hot kernel next to the product library (similaripy_b200/libspy_<name>.so, git-ignored,
shipped to the GPU box by gpurun), each from the compile-time knobs of knn_kernel.cuh:

    python scripts/build_variants.py noprefetch=-DSPY_PREFETCH=0 nospec=-DSPY_SPECULATE=0 unroll4=-DSPY_UNROLL=4,-DSPY_DRAIN_BATCH=4

then compare them on a B200, cheapest first:

    scripts/variant_probe 1000000 200000 200 100 50000 similaripy_b200/libsimilaripy_b200.so similaripy_b200/libspy_*.so
    SPY_LIB_TESTS=1 SPY_LIBS="$(ls similaripy_b200/libspy_*.so | tr '\\n' ' ')" bash scripts/gpu_iter.sh <tag>
"""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from similaripy_b200.csrc import build  # noqa: E402

DEFAULTS = {"noprefetch": ["-DSPY_PREFETCH=0"], "nospec": ["-DSPY_SPECULATE=0"], "nosharpen": ["-DSPY_SHARPEN=0"]}


def main(argv):
    variants = dict(a.split("=", 1) for a in argv) if argv else None
    variants = {k: v.split(",") for k, v in variants.items()} if variants else DEFAULTS
    for name, flags in variants.items():
        lib = f"libspy_{name}.so"
        out = build.build(extra_flags=flags, out_path=os.path.join(build.PKG, lib))
        log = open(os.path.join(build.HERE, "build_" + lib, "ptxas.log")).read()
        spills = re.findall(r"knn_flat_kernelILi1024ELi2ELb1ELi8E\w+\n\s+(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)", log)
        print(out, " ".join(flags), "| cosine kernel (1024 threads, 8 lanes):", spills[0] if spills else "?")


if __name__ == "__main__":
    main(sys.argv[1:])
