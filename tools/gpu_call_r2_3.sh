#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export BZ3_B200_AUTOSELECT=0
./tools/variants/ubench_walk > gpurun_out/r2c_ubench_walk.log 2>&1; cat gpurun_out/r2c_ubench_walk.log
timeout 240 python tools/eval_variants.py --mib 1 --reps 2 --enc 0,10 --dec 0 --out gpurun_out/r2c_eval_1MiB.json > gpurun_out/r2c_eval_1MiB.log 2>&1
echo "eval1 rc $?"; grep "enc v" gpurun_out/r2c_eval_1MiB.log
BZ3_PROF_DEC=0 BZ3_PROF_ENC=0,10 timeout 120 python tools/cm_prof2.py 1 > gpurun_out/r2c_cm_prof.log 2>&1
echo "prof rc $?"; cat gpurun_out/r2c_cm_prof.log
