#!/bin/bash
# full evidence pass for one kernel version: bash scripts/gpu_round.sh <tag>
#   smoke, GPU parity tests, full bench line (value/e2e/roofline/cpu_baseline), reference arm,
#   ncu launch list of the bench command, ncu --set full capture of the hot kernel
TAG=${1:-vX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu_$TAG.txt; nproc >> gpurun_out/gpu_$TAG.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
if [ -z "$SKIP_REF" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --ref-budget-s 90 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_$TAG.json
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_ -c 1 -o gpurun_out/prof_knn_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
# DRAM traffic of the hot kernel on the other configurations (configs[4]: B is L2 resident): one launch each
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:knn_stream --csv \
    --log-file gpurun_out/configs_traffic_$TAG.csv python scripts/bench_configs.py cfg3 cfg4 cfg5 > gpurun_out/configs_traffic_$TAG.log 2>&1; echo "ncu configs rc=$?"
ls -la gpurun_out | tail -20
