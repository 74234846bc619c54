set -u
mkdir -p gpurun_out
L=gpurun_out/r3g_final_sanity.log; : > $L
( timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2 ) | tee -a $L
( timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q -k "not 511mib and not 256mib and not 64mib and not 32mib" 2>&1 | tail -3 ) | tee -a $L
