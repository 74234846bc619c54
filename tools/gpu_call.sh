set -u
mkdir -p gpurun_out
for mt in 1 2; do
  echo "model thread $mt"
  for g in zipf_text source_corpus mixed; do
    BZ3_B200_CM_MT=$mt timeout 300 python tools/stage_driver.py cm_dec_bwt 4 2 $g | tail -1
  done
done 2>&1 | tee gpurun_out/r2m_cm_mt_ab.log
