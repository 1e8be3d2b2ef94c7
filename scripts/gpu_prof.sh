#!/bin/bash
# ncu full capture of the hot kernel on the cfg2 bench (first launch), plus the launch list
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_flat -c 1 -o gpurun_out/prof_knn_v2 -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_v2.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_v2.log
