#!/usr/bin/env python
"""Hot-kernel numbers of BASELINE.json configs[2..4] (bench.other_configs: full-size operands, sampled target rows) for
BOTH kernel generations -- which one the planner should prefer per configuration.
usage: python scripts/bench_engines.py [flat stream]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from similaripy_b200 import _engine

peak, _ = bench.load_peak()
for eng in (sys.argv[1:] or ["flat", "stream"]):
    _engine.DEFAULT_TUNING.clear()
    _engine.DEFAULT_TUNING["engine_prefer"] = eng
    for line in bench.other_configs(0, peak, 1.0):
        line["engine_preferred"] = eng
        print(json.dumps(line), flush=True)
