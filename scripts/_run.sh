mkdir -p gpurun_out
L=similaripy_b200
echo "== probe cfg2-shaped" > gpurun_out/probe_q.txt
SPY_PROBE_PHASES=1 timeout 600 scripts/variant_probe 1000000 200000 200 100 50000 $L/libspy_head.so@2 $L/libsimilaripy_b200.so@2 $L/libspy_nohint.so@2 $L/libspy_hintstress.so@2 $L/libspy_kstiming.so@2 >> gpurun_out/probe_q.txt 2>&1
echo "== probe short rows (cfg4-like)" >> gpurun_out/probe_q.txt
SPY_PROBE_PHASES=1 timeout 600 scripts/variant_probe 200000 200000 140 100 100000 $L/libsimilaripy_b200.so@1 $L/libsimilaripy_b200.so@2 $L/libspy_nohint.so@2 $L/libspy_kstiming.so@2 >> gpurun_out/probe_q.txt 2>&1
cat gpurun_out/probe_q.txt | grep -v "^    "
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_q.log
SIMILARIPY_B200_LIB=$PWD/$L/libspy_hintstress.so timeout 600 python -m pytest tests/test_similarity_gpu.py tests/test_configs_gpu.py tests/test_golden_gpu.py -m gpu -x -q > gpurun_out/pytest_q_hintstress.log 2>&1; echo "hintstress pytest rc=$?"; tail -3 gpurun_out/pytest_q_hintstress.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; echo "bench rc=$?"
SIMILARIPY_B200_LIB=$PWD/$L/libspy_nohint.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/bench_q_nohint.json 2>> gpurun_out/bench_q.err
python - <<'PY'
import json
for f in ['bench_q','bench_q_nohint']:
    j=json.load(open(f'gpurun_out/{f}.json')); r=j['roofline']; print(f, j['ms_per_step'], r['kernel_ms'], r['frac'], r['gproducts_per_s'])
PY
