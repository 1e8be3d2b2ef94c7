#!/bin/bash
# first GPU pass: smoke, parity tests, small + full bench, ncu launch list + full capture of the hot kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --scale 0.1 --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full rc=$?"
cat gpurun_out/bench_full.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_panel -s 1 -c 1 -o gpurun_out/prof_knn \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
