set -u
mkdir -p gpurun_out
BZ3_B200_ARENAS=4 timeout 45 python -m pytest tests/test_gpu_large.py -x -q -k "many_blocks or batch_of_16mib" 2>&1 | tail -3 | tee gpurun_out/r3h_arenas4.log
