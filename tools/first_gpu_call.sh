#!/usr/bin/env bash
# One GPU call that answers every open question of DESIGN.md 6c (run from the repo root on a B200 box):
#   /usr/local/graft/bin/gpurun --timeout 720 -- 'bash tools/first_gpu_call.sh'
# 1. parity + time of every entropy / LZP kernel variant (ctypes only, starts in seconds)
# 2. clock64 phase breakdown of the decoders and encoders (needs tools/variants/lib_cmprof.so, built by
#    `nvcc ... -DBZ_CM_PROFILE`)
# 3. the GPU parity tests of the variants that are still opt-in (stage tests only: fast)
# Every step has its own timeout; logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 150 python tools/eval_variants.py --mib 1 --reps 2 --out gpurun_out/eval_next.json > gpurun_out/eval_next.log 2>&1
echo "eval rc $?"; tail -60 gpurun_out/eval_next.log
if [ -f tools/variants/lib_cmprof.so ]; then
  timeout 90 python tools/cm_prof2.py 1 > gpurun_out/cm_prof_next.log 2>&1
  echo "prof rc $?"; cat gpurun_out/cm_prof_next.log
fi
BZ3_B200_TEST_NEW=1 timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_cm or stage_lzp" > gpurun_out/pytest_new_variants.log 2>&1
echo "pytest(new variants) rc $?"; tail -4 gpurun_out/pytest_new_variants.log
# 4. host-side changes made after the last GPU call: workspace pool under 48 blocks in flight, container front end
timeout 300 python -m pytest tests/test_gpu_stream.py -m gpu -x -q > gpurun_out/pytest_stream.log 2>&1
echo "pytest(stream / many blocks) rc $?"; tail -4 gpurun_out/pytest_stream.log
# (separately, ~3 GPU-minutes: what 128 blocks in flight buy)
#   python bench.py --workload zipf2g_b16 --steps 1 --warmup 3 > gpurun_out/bench_zipf2g_b16.json
