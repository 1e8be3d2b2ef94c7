"""CPU: bench.py's synthetic generator builds the same bits with numpy (reference arm) and with torch (GPU arm; run
on the CPU device here), and the planner resolves the kernel generations as documented."""
import ctypes

import numpy as np
import pytest

import bench


@pytest.mark.parametrize("shape", [(5000, 3000, 0.01, 7), (300, 50, 0.2, 2), (1000, 2000, 0.025, 0)])
def test_host_and_device_generators_are_bit_identical(shape):
    import torch
    n_rows, n_cols, density, seed = shape
    m = bench.gen_urm_host(n_rows, n_cols, density, seed)
    ip, ix, dv = bench.gen_urm_device(n_rows, n_cols, density, seed, torch.device("cpu"))
    assert np.array_equal(m.indptr, ip.numpy()) and np.array_equal(m.indices, ix.numpy())
    assert np.array_equal(m.data, dv.numpy()) and m.data.dtype == np.float32
    assert m.data.min() > 0.0 and m.data.max() <= 1.0
    # rows: distinct ascending columns, nnz close to n_rows * n_cols * density
    for r in (0, n_rows // 2, n_rows - 1):
        row = m.indices[m.indptr[r]: m.indptr[r + 1]]
        assert np.all(np.diff(row) > 0)
    assert abs(m.nnz / (n_rows * n_cols * density) - 1.0) < 0.1


def test_engine_resolution_in_the_planner():
    from similaripy_b200 import _lib
    lib = _lib.load()

    def plan(**kw):
        a = _lib.KnnArgs()
        a.k, a.n_cols, a.n_targets, a.a1 = 100, 200_000, 1000, 1.0
        for k, v in kw.items():
            setattr(a, k, v)
        return lib.spy_knn_plan(ctypes.byref(a), -1), a

    rc, a = plan(engine=_lib.ENGINE_STREAM)
    assert rc == 0 and a.engine == _lib.ENGINE_STREAM and a.panel_width % 2048 == 0 and a.threads == 1024
    assert a.panel_width * a.n_panels >= 200_000 and a.panel_width <= 65536
    rc, a = plan(engine=_lib.ENGINE_FLAT)
    assert rc == 0 and a.engine == _lib.ENGINE_FLAT and a.panel_width % 128 == 0
    rc, a = plan()  # auto resolves to one of the two
    assert rc == 0 and a.engine in (_lib.ENGINE_FLAT, _lib.ENGINE_STREAM)
    # what the stream engine does not cover is refused when asked for explicitly ...
    for kw in (dict(target_mode=_lib.SEL_MATRIX), dict(a1=0.5), dict(bayesian_shrink=1.0), dict(k=600), dict(threads=512)):
        rc, _ = plan(engine=_lib.ENGINE_STREAM, **kw)
        assert rc == _lib.ERR_UNSUPPORTED, kw
        rc, a = plan(**kw)  # ... and takes the flat engine on its own
        assert rc == 0 and a.engine == _lib.ENGINE_FLAT, kw
    # the build of the stream kernel follows the products per (row, panel) when the planner knows the operand sizes:
    # configs[4] (200 x 100 products per row, 5 panels) takes 16 drain warps, configs[1] (2e5 products per row) 8
    rc, a = plan(a_rows=5_000_000, a_nnz=1_000_000_000, b_rows=200_000, b_nnz=20_000_000)
    assert rc == 0 and a.engine == _lib.ENGINE_STREAM and a.group == 16
    rc, a = plan(a_rows=200_000, a_nnz=200_000_000, b_rows=1_000_000, b_nnz=200_000_000)
    assert rc == 0 and a.engine == _lib.ENGINE_STREAM and a.group == 8
    w8 = a.panel_width
    # ... an explicit request wins, and a plan that is fed back (engine and group as returned) resolves to itself
    rc, a = plan(engine=_lib.ENGINE_STREAM, group=16, a_rows=200_000, a_nnz=200_000_000, b_rows=1_000_000, b_nnz=200_000_000)
    assert rc == 0 and a.group == 16 and a.panel_width == w8  # (both builds hold a 40960-column panel)
    rc, a = plan(engine=_lib.ENGINE_STREAM, group=8, a_rows=5_000_000, a_nnz=1_000_000_000, b_rows=200_000, b_nnz=20_000_000)
    assert rc == 0 and a.group == 8
    rc, a = plan(group=16)  # the flat engine's lanes per segment do not select a build when the engine is not asked for
    assert rc == 0 and (a.engine == _lib.ENGINE_FLAT or a.group in (8, 16))
    rc, _ = plan(engine=7)
    assert rc < 0
    rc, _ = plan(threads=768)  # experiment builds only
    assert rc < 0 and b"threads must be" in lib.spy_last_error()


def test_clock_sampler_counts_only_the_timed_region():
    """bench.ClockSampler: nvidia-smi is started before the warm-up steps, mark() opens the timed region; a region shorter
    than the sampling period falls back to the last samples of the warm-up steps and says so."""
    line = lambda mhz, slow="Not Active": f"0, {mhz}, 1965, 700.0, 0x0, {slow}, Not Active, Not Active, Active"
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    s.lines = [line(1200), line(1500)]            # while nvidia-smi came up / warm-up
    s.mark()
    s.lines += [line(1965), line(1950), line(1965)]
    out = s.stop()
    assert out["sm_mhz"] == 1965.0 and out["samples"] == 3 and out["reasons"] == ["sw_power_cap"] and "note" not in out
    s2 = bench.ClockSampler(0)
    s2.proc = s.proc
    s2.lines = [line(1800), line(1965, "Active")]
    s2.mark()                                     # no sample inside the region
    out2 = s2.stop()
    assert out2["samples"] == 2 and "hw_slowdown" in out2["reasons"] and "shorter than the sampling period" in out2["note"]
    s3 = bench.ClockSampler(0)
    s3.proc = s.proc
    assert s3.stop()["reasons"] == ["no samples"]
