#!/bin/bash
# iteration pass: parity tests + bench (+ optional ncu full capture of the hot kernel): bash scripts/gpu_iter.sh [tag] [prof]
# SPY_VARIANTS="threads=512 ..."  extra bench runs with SPY_TUNING=<variant>;  SPY_LIBS="path.so ..." extra runs with other builds
# (built here with similaripy_b200.csrc.build.build(extra_flags=[...], out_path=...)); SPY_LIB_TESTS=1: parity tests on each of them first
TAG=${1:-iter}; PROF=${2:-}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_$TAG.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); r = j['roofline']
    print(sys.argv[2], 'value', j['value'], 'ms/step', j['ms_per_step'], 'kernel_ms', r['kernel_ms'], 'frac', r['frac'], 'Gprod/s', r['gproducts_per_s'],
          'e2e', j['e2e'] and j['e2e']['value'], r['plan'])
except Exception as e:
    print(sys.argv[2], 'FAILED', e)
PY
}
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
show gpurun_out/bench_$TAG.json main
for variant in $SPY_VARIANTS; do
  SPY_TUNING="$variant" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > "gpurun_out/bench_${TAG}_${variant}.json" 2>> gpurun_out/bench_$TAG.err
  show "gpurun_out/bench_${TAG}_${variant}.json" "$variant"
done
for lib in $SPY_LIBS; do
  b=$(basename $lib .so)
  if [ -n "$SPY_LIB_TESTS" ]; then  # parity of the variant build first: a fast wrong kernel is not a result
    SIMILARIPY_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_similarity_gpu.py tests/test_configs_gpu.py -m gpu -x -q > "gpurun_out/pytest_gpu_${TAG}_${b}.log" 2>&1
    echo "$b pytest rc=$?"; tail -2 "gpurun_out/pytest_gpu_${TAG}_${b}.log"
  fi
  SIMILARIPY_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > "gpurun_out/bench_${TAG}_${b}.json" 2>> gpurun_out/bench_$TAG.err
  show "gpurun_out/bench_${TAG}_${b}.json" "$b"
done
if [ -n "$PROF" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_flat -c 1 -o gpurun_out/prof_knn_$TAG -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
