set -u
mkdir -p gpurun_out
L=gpurun_out/r3a_ab.log; : > $L
for c in zipf_text source_corpus mixed; do
  echo "== decoder, tier B rolled (small code) $c" >> $L
  timeout 120 python tools/stage_driver.py cm_dec_bwt 4 2 $c 2>&1 | tail -1 >> $L
  echo "== decoder, previous build (tier B unrolled) $c" >> $L
  BZ3_B200_LIB=tools/variants/lib_prev.so timeout 120 python tools/stage_driver.py cm_dec_bwt 4 2 $c 2>&1 | tail -1 >> $L
done
echo "== lzp 64 MiB source corpus, one round trip for the tail compares" >> $L
timeout 200 python tools/stage_driver.py lzp 64 2 source_corpus 2>&1 | tail -1 >> $L
echo "== lzp 64 MiB source corpus, previous build" >> $L
BZ3_B200_LIB=tools/variants/lib_prev.so timeout 200 python tools/stage_driver.py lzp 64 2 source_corpus 2>&1 | tail -1 >> $L
timeout 200 python tools/cm_prof2.py 1 >> $L 2>&1
cat $L
