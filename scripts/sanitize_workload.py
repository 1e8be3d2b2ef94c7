"""Workload for compute-sanitizer (memcheck / racecheck): dense candidate sets (buffer overflow, speculative bound, several
panels, both CTA shapes), matrix filter / target selectors, normalizers.
    compute-sanitizer --tool memcheck  python scripts/sanitize_workload.py
    compute-sanitizer --tool racecheck python scripts/sanitize_workload.py
Round 1 on B200: 0 errors, 0 hazards.  Round 2 adds the stream kernel (both builds, dense + sparse hand-overs) and the block path."""
import numpy as np, scipy.sparse as sp, sys
sys.path.insert(0, '/root/repo')
import similaripy_b200 as sim
rng = np.random.default_rng(5)
# dense candidates: overflow + speculation + several panels, plus a matrix filter and a matrix target
urm = sp.random_array((800, 14000), density=0.03, format="csr", dtype=np.float32, random_state=rng)
a = urm.T.tocsr()
rows = list(range(0, 14000, 700))
for name, kw in (("cosine", {}), ("rp3beta", dict(alpha=0.9, beta=0.4)), ("jaccard", {}), ("dot_product", {})):
    for tuning in (None, dict(panel_width=2048), dict(threads=512, panel_width=1024, group=4)):
        r = getattr(sim, name)(a, urm, k=50, target_rows=rows, verbose=False, format_output="csr", tuning=tuning, **kw)
print("dense ok", r.nnz)
u2 = sp.random_array((300, 500), density=0.05, format="csr", dtype=np.float32, random_state=rng)
s2 = sp.random_array((500, 500), density=0.06, format="csr", dtype=np.float32, random_state=rng)
r = sim.dot_product(u2, s2, k=12, filter_cols=u2, verbose=False, format_output="csr")
tm = sp.random_array((300, 500), density=0.3, format="csr", dtype=np.float32, random_state=rng)
r = sim.dot_product(u2, s2, k=12, target_cols=tm, verbose=False, format_output="coo")
w = sim.bm25(u2); w = sim.tfidf(u2); w = sim.normalize(u2, "l1")
print("selectors + normalizers ok")
# stream kernel, both builds: dense snapshots and double-buffered sparse hand-overs interleaved
sys.path.insert(0, '/root/repo/tests')
from test_stream_handover_gpu import _operands
from similaripy_b200 import _engine
a2, b2 = _operands(41)
for dw in (8, 16):
    for name in ("dot_product", "cosine"):
        r = getattr(sim, name)(a2[:48], b2, k=40, verbose=False, format_output="csr", tuning=dict(engine="stream", drain_warps=dw, panel_width=10240))
print("stream hand-overs ok", r.nnz)
# matrices beyond int32 stored entries: the block path forced on a small matrix (row blocks x column blocks, slab merge)
_engine.WIDE_NNZ_LIMIT = 3000
m = sp.random_array((500, 400), density=0.05, format="csr", dtype=np.float32, random_state=rng)
r = sim.cosine(m, k=20, verbose=False, format_output="csr")
r = sim.cosine(m.T, k=20, verbose=False, format_output="csr")
print("wide blocks ok", r.nnz)
# passes of more than 32 staged blocks (target rows of 1025..1056 entries) and rows that need a second pass
n_u = 3000
b3 = sp.random_array((n_u, 6000), density=20 / 6000, format="csr", dtype=np.float32, random_state=rng)
rows3, cols3 = [], []
for r, n in enumerate((1030, 1056, 1057, 1100, 40, 2200)):
    c = rng.choice(n_u, size=n, replace=False)
    rows3 += [r] * n; cols3 += c.tolist()
a3 = sp.csr_array((rng.random(len(rows3)).astype(np.float32) + 0.1, (np.asarray(rows3), np.asarray(cols3))), shape=(6, n_u))
for dw in (8, 16):
    r = sim.cosine(a3, b3, k=30, verbose=False, format_output="csr", tuning=dict(engine="stream", drain_warps=dw, panel_width=2048))
print("long passes ok", r.nnz)
