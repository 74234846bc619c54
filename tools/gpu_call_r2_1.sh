#!/usr/bin/env bash
# round 2, GPU call 1: time every shipped/candidate kernel on real data, then the blocks-in-flight curve
set -u
mkdir -p gpurun_out
export BZ3_B200_AUTOSELECT=0
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_1_smi.log 2>&1; nproc >> gpurun_out/r2_1_smi.log; free -g >> gpurun_out/r2_1_smi.log
timeout 240 python tools/eval_variants.py --mib 1 --reps 2 --out gpurun_out/r2_eval_1MiB.json > gpurun_out/r2_eval_1MiB.log 2>&1
echo "eval1 rc $?"; tail -70 gpurun_out/r2_eval_1MiB.log
timeout 300 python tools/eval_variants.py --mib 16 --reps 1 --sets zipf_text --enc 0,6 --dec 0,8,9 --out gpurun_out/r2_eval_16MiB.json > gpurun_out/r2_eval_16MiB.log 2>&1
echo "eval16 rc $?"; tail -20 gpurun_out/r2_eval_16MiB.log
timeout 120 python tools/cm_prof2.py 1 > gpurun_out/r2_cm_prof.log 2>&1
echo "prof rc $?"; cat gpurun_out/r2_cm_prof.log
timeout 300 python tools/inflight_curve.py --mib 4 --ks 1,6,37,74,148,296 --out gpurun_out/r2_inflight_zipf4m.json > gpurun_out/r2_inflight_zipf4m.log 2>&1
echo "inflight rc $?"; cat gpurun_out/r2_inflight_zipf4m.log
