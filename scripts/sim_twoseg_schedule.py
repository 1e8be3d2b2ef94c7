#!/usr/bin/env python
"""CPU model of the entry schedule of accumulate_chunk_two (knn_kernel.cuh, -DSPY_TWOSEG=1): which lane holds which
entry's bounds when, through the claims, the far fetch by lane role and the shuffles.  Checks that every staged entry
0..n-1 is expanded exactly once and with ITS OWN bounds, for every group width and many interleavings of the warps.
(A desk check for a kernel variant that could not be run when it was written; not part of the test suite.)"""
import random
import sys


def run(n, NT, G, seed):
    rng = random.Random(seed)
    GROUPS, GPW, NW = NT // G, 32 // G, NT // 32
    bounds = lambda i: (1000 + 7 * i, 1000 + 7 * i + 3) if i < n else (0, 0)   # what fetch(i) returns
    counter = [4 * GROUPS]
    done = {}

    def warp(w):
        lanes = range(32)
        gl = [l % G for l in lanes]; gw = [l // G for l in lanes]
        role = [g & 3 for g in gl]; gbase = [l - l % G for l in lanes]
        base0 = w * 2 * GPW; base1 = base0 + 2 * GROUPS
        sA = [bounds(base0 + gw[l]) for l in lanes]
        sB = [bounds(base0 + GPW + gw[l]) for l in lanes]
        f = [bounds(base1 + (role[l] & 1) * GPW + gw[l]) for l in lanes]
        while base0 < n:
            base2 = n
            if base1 < n:
                base2 = counter[0]; counter[0] += 2 * GPW
            yield  # other warps may claim in between
            for l in lanes:
                if role[l] >= 2:
                    f[l] = bounds(base2 + (role[l] & 1) * GPW + gw[l])
            for l in lanes:
                iA = base0 + gw[l]; iB = iA + GPW
                for i, b in ((iA, sA[l]), (iB, sB[l])):
                    if i < n:
                        assert b == bounds(i), (w, l, i, b)
                        if gl[l] == 0:
                            assert i not in done, ("twice", i)
                            done[i] = True
                    else:
                        assert b == (0, 0) or True
            nsA = [f[gbase[l]] for l in lanes]
            nsB = [f[gbase[l] + 1] for l in lanes]
            nf = [f[l + 2] if (l % 4) + 2 < 4 else f[l] for l in lanes]   # __shfl_down_sync(.., 2, 4)
            sA, sB, f = nsA, nsB, nf
            base0, base1 = base1, base2

    gens = [warp(w) for w in range(NW)]
    live = list(gens)
    while live:
        g = rng.choice(live)
        try:
            next(g)
        except StopIteration:
            live.remove(g)
    assert sorted(done) == list(range(n)), (n, NT, G, sorted(set(range(n)) - set(done))[:5])


if __name__ == "__main__":
    cases = 0
    for NT in (512, 768, 1024):
        for G in (4, 8, 16, 32):
            for n in (0, 1, 3, 31, 64, 127, 128, 129, 255, 256, 257, 500, 511, 512, 513, 1000, 1279, 1280, 1600):
                for seed in range(3):
                    run(n, NT, G, seed); cases += 1
    print("ok:", cases, "cases")
