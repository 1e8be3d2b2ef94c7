#!/usr/bin/env python
"""Build the experimental variants of the hot kernel next to the product library (similaripy_b200/libspy_*.so, git-ignored,
shipped to the GPU box by gpurun).  Measure them with
    SPY_LIB_TESTS=1 SPY_LIBS="$(ls similaripy_b200/libspy_*.so | tr '\\n' ' ')" bash scripts/gpu_iter.sh <tag>
"""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from similaripy_b200.csrc import build

TWO = ["-DSPY_TWOSEG=1"]
VARIANTS = {
    "libspy_early.so": ["-DSPY_EARLY_GATHER=1"],
    "libspy_twoseg.so": TWO + ["-DSPY_TWOSEG_INLINE=__forceinline__"],
    "libspy_twoseg_early.so": TWO + ["-DSPY_TWOSEG_INLINE=__forceinline__", "-DSPY_EARLY_GATHER=1"],
    "libspy_twoseg_ni.so": TWO + ["-DSPY_TWOSEG_INLINE=__noinline__"],
    "libspy_twoseg_ni_early.so": TWO + ["-DSPY_TWOSEG_INLINE=__noinline__", "-DSPY_EARLY_GATHER=1"],
}
only = sys.argv[1:]
for name, flags in VARIANTS.items():
    if only and name not in only:
        continue
    out = build.build(extra_flags=flags, out_path=os.path.join(build.PKG, name))
    log = open(os.path.join(build.HERE, "build_" + name, "ptxas.log")).read()
    spills = re.findall(r"knn_flat_kernelILi1024ELi2ELb1ELi8E\w+\n\s+(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)", log)
    print(out, " ".join(flags), "| cosine kernel:", spills[0] if spills else "?")
