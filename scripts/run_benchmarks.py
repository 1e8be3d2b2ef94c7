#!/usr/bin/env python
"""Benchmark-harness parity: the upstream report of tests/benchmarks/run_benchmarks.py:336-378 (metadata /
config / datasets / results with computation_time, std_time, throughput, nnz, avg_neighbors, rounds, all_times;
metric definitions of tests/benchmarks/benchmark.py:184-189: wall clock around the PUBLIC call on a host scipy matrix,
throughput = n_items / time) produced by similaripy_b200 -- and, with --with-reference, by the compiled reference on the
same box -- on a MovieLens-32M-SHAPED synthetic URM: 200 948 users x 84 432 items, 32 M interactions
(tests/benchmarks/README.md:194-205; the real file needs a download), power-law user activity and item popularity.

    python scripts/run_benchmarks.py [--similarities cosine rp3beta ...] [--rounds 3] [--k 100] [--with-reference]
                                     [--output gpurun_out/benchmark_movielens32m_shaped.json]
"""
import argparse, json, os, platform, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scipy.sparse as sp


def movielens_shaped(n_users=200_948, n_items=84_432, nnz=32_000_204, seed=32):
    """Zipf-like activity (exponent 0.8) and popularity (exponent 1.0): ratings in {0.5, 1.0, ..., 5.0}."""
    rng = np.random.default_rng(seed)
    pu = 1.0 / np.arange(1, n_users + 1) ** 0.8
    pi = 1.0 / np.arange(1, n_items + 1) ** 1.0
    draw = int(nnz * 1.35)  # duplicates are merged below
    u = rng.choice(n_users, size=draw, p=pu / pu.sum()).astype(np.int64)
    i = rng.choice(n_items, size=draw, p=pi / pi.sum()).astype(np.int64)
    key = np.unique(rng.permutation(n_users)[u] * n_items + rng.permutation(n_items)[i])
    if key.shape[0] > nnz:
        key = np.sort(rng.choice(key, size=nnz, replace=False))
    r = key // n_items
    data = (rng.integers(1, 11, size=key.shape[0]) * 0.5).astype(np.float32)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=n_users), out=indptr[1:])
    return sp.csr_array((data, (key - r * n_items).astype(np.int32), indptr.astype(np.int32)), shape=(n_users, n_items))


def run(fn, item_matrix, k, shrink, threshold, rounds, extra):
    times, res = [], None
    for _ in range(rounds):
        t0 = time.perf_counter()
        res = fn(item_matrix, k=k, shrink=shrink, threshold=threshold, verbose=False, num_threads=0, block_size=0, **extra)
        times.append(time.perf_counter() - t0)
    n_items = res.shape[0]
    mean = float(np.mean(times))
    return {"computation_time": round(mean, 4), "std_time": round(float(np.std(times)), 4), "throughput": round(n_items / mean, 1),
            "nnz": int(res.nnz), "avg_neighbors": round(res.nnz / n_items, 1), "rounds": rounds, "all_times": [round(t, 4) for t in times]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--similarities", nargs="+", default=["cosine", "dot_product", "rp3beta"])
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--shrink", type=float, default=0.0)
    ap.add_argument("--threshold", type=float, default=0.0)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--with-reference", action="store_true")
    ap.add_argument("--output", default=os.path.join(ROOT, "gpurun_out", "benchmark_movielens32m_shaped.json"))
    args = ap.parse_args()
    import similaripy_b200 as sim
    urm = movielens_shaped()
    item_matrix = urm.T  # item-item, benchmark.py:161
    extras = {"rp3beta": dict(alpha=1.0, beta=0.6), "p3alpha": dict(alpha=1.0), "asymmetric_cosine": dict(alpha=0.5),
              "tversky": dict(alpha=1.0, beta=1.0)}
    key = "movielens-shaped:32m"
    report = {"metadata": {"platform": platform.platform(), "python": platform.python_version(), "cpu_count": os.cpu_count(),
                           "implementation": "similaripy_b200 (B200, sm_100a)", "note": "synthetic MovieLens-32M-shaped URM (power law); "
                           "the first round includes one-off CUDA context / library start-up"},
              "config": {"datasets": [["movielens-shaped", "32m"]], "similarities": args.similarities, "k": args.k, "shrink": args.shrink,
                         "threshold": args.threshold, "num_threads": 0, "block_size": "default", "rounds": args.rounds},
              "datasets": {key: {"shape": list(urm.shape), "nnz": int(urm.nnz), "density": round(urm.nnz / (urm.shape[0] * urm.shape[1]), 8)}},
              "results": {key: {}}}
    sim.cosine(item_matrix[:64], item_matrix[:64].T, k=4, verbose=False)  # start-up outside the first measured round
    for name in args.similarities:
        report["results"][key][name] = run(getattr(sim, name), item_matrix, args.k, args.shrink, args.threshold, args.rounds, extras.get(name, {}))
        print(name, report["results"][key][name], flush=True)
    if args.with_reference:
        from oracle import ref_api
        if ref_api.available():
            report["reference_results"] = {key: {}}
            for name in args.similarities:
                fn = lambda m, **kw: ref_api.similarity(name, m, **{k: v for k, v in kw.items() if k != "verbose"}, format_output="csr")
                report["reference_results"][key][name] = run(fn, item_matrix.tocsr(), args.k, args.shrink, args.threshold, 1, extras.get(name, {}))
                print("reference", name, report["reference_results"][key][name], flush=True)
    os.makedirs(os.path.dirname(args.output), exist_ok=True)
    json.dump(report, open(args.output, "w"), indent=2)


if __name__ == "__main__":
    main()
