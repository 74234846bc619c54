set -u
mkdir -p gpurun_out
for tma in 1 0; do
  BZ3_B200_RS_TMA=$tma timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_ncu_bwt64_tma$tma.csv python tools/stage_driver.py bwt 64 1 > /dev/null 2>&1
  echo "tma=$tma"; grep "rs_onesweep_kernel<unsigned long" gpurun_out/r2h_ncu_bwt64_tma$tma.csv | awk -F'","' '{print $9, $NF}' | head -9 | tr '\n' ' '; echo
  BZ3_B200_RS_TMA=$tma timeout 300 python tools/stage_driver.py bwt 64 3 | tail -2
  BZ3_B200_RS_TMA=$tma timeout 300 python tools/stage_driver.py bwt 16 4 | tail -2
done 2>&1 | tee gpurun_out/r2h_tma_ab.log
for spec in "lzp 64 1 source_corpus" "lzp 16 1 log_stream" "lzp 256 1 source_corpus"; do
  timeout 300 python tools/stage_driver.py $spec 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2h_lzp.log
