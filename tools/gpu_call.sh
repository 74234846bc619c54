set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_ncu_bwt64.csv python tools/stage_driver.py bwt 64 1 > /dev/null 2>&1
grep "rs_onesweep_kernel<unsigned long" gpurun_out/r2i_ncu_bwt64.csv | awk -F'","' '{print $9, $NF}' | head -9 | tr '\n' ' '; echo
timeout 300 python tools/stage_driver.py bwt 64 3 | tail -2
timeout 300 python tools/stage_driver.py bwt 16 4 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bwt or block" 2>&1 | tail -2
