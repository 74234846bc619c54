#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export BZ3_B200_AUTOSELECT=0
timeout 240 python tools/eval_variants.py --mib 1 --reps 2 --enc 0 --dec 0,10 --out gpurun_out/r2b_eval_1MiB.json > gpurun_out/r2b_eval_1MiB.log 2>&1
echo "eval1 rc $?"; grep -v "^  lzp" gpurun_out/r2b_eval_1MiB.log | tail -20
timeout 200 python tools/eval_variants.py --mib 16 --reps 1 --sets zipf_text --enc 0 --dec 0,10 --out gpurun_out/r2b_eval_16MiB.json > gpurun_out/r2b_eval_16MiB.log 2>&1
echo "eval16 rc $?"; grep "dec v" gpurun_out/r2b_eval_16MiB.log
BZ3_PROF_DEC=0,10 BZ3_PROF_ENC=0 timeout 120 python tools/cm_prof2.py 1 > gpurun_out/r2b_cm_prof.log 2>&1
echo "prof rc $?"; cat gpurun_out/r2b_cm_prof.log
