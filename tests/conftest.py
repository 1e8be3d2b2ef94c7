import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- both kernel generations: modules that list the `spy_engine` fixture run every test with the flat engine and with
#      the stream engine preferred, in both of its builds (8 and 16 drain warps); configurations the stream engine does
#      not cover fall back to the flat one ----
def pytest_generate_tests(metafunc):
    if "spy_engine" in metafunc.fixturenames:
        metafunc.parametrize("spy_engine", ["flat", "stream8", "stream16"], indirect=True)


@pytest.fixture
def spy_engine(request):
    from similaripy_b200 import _engine
    saved = dict(_engine.DEFAULT_TUNING)
    if request.param == "flat":
        _engine.DEFAULT_TUNING["engine_prefer"] = "flat"
    else:
        _engine.DEFAULT_TUNING["engine_prefer"] = "stream"
        _engine.DEFAULT_TUNING["drain_warps"] = int(request.param[len("stream"):])
    yield request.param
    _engine.DEFAULT_TUNING.clear()
    _engine.DEFAULT_TUNING.update(saved)
