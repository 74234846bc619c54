set -u
mkdir -p gpurun_out
L=gpurun_out/r3f_ab.log; : > $L
( timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream.py -x -q 2>&1 | tail -3 ) | tee -a $L
for c in zipf_text source_corpus; do
  echo "== decoder, row address carried in the fast tier $c" >> $L
  timeout 120 python tools/stage_driver.py cm_dec_bwt 4 2 $c 2>&1 | tail -1 >> $L
  echo "== decoder, previous build $c" >> $L
  BZ3_B200_LIB=tools/variants/lib_prev.so timeout 120 python tools/stage_driver.py cm_dec_bwt 4 2 $c 2>&1 | tail -1 >> $L
done
echo "== lzp 64 MiB zipf_text: unrolled warp-sum when all warps are active" >> $L
timeout 200 python tools/stage_driver.py lzp 64 2 zipf_text 2>&1 | tail -1 >> $L
cat $L
