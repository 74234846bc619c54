set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/r2e_pytest_parity.log 2>&1
echo "pytest parity rc $?"; tail -6 gpurun_out/r2e_pytest_parity.log
for spec in "lzp 16 2 zipf_text" "lzp 64 2 source_corpus" "lzp 16 2 log_stream" "lzp 256 1 zipf_text" "bwt 16 3 zipf_text" "bwt 64 2 source_corpus" "unbwt 16 2 zipf_text"; do
  timeout 300 python tools/stage_driver.py $spec 2>&1 | tail -3
done > gpurun_out/r2e_stage_times.log 2>&1
cat gpurun_out/r2e_stage_times.log
( time timeout 900 python bench.py --steps 1 --warmup 1 --no-headline ) > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print('e2e', d['e2e']['value'], 'value', d['value'], 'sa', d['roofline_sa_radix']['frac'], d['roofline_sa_radix']['ms'], 'stages', d['stage_ms_per_step'])
PY
tail -3 gpurun_out/r2e_bench.err
