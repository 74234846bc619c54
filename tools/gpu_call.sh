set -u
mkdir -p gpurun_out
for r in 0 1; do
  echo "decoder exact tier rolled=$r"
  for g in zipf_text source_corpus mixed; do
    BZ3_B200_CM_ROLL=$r timeout 300 python tools/stage_driver.py cm_dec_bwt 4 2 $g | tail -1
  done
done 2>&1 | tee gpurun_out/r2v_cm_dec_roll_ab.log
timeout 300 python tools/stage_driver.py cm_enc_bwt 4 2 zipf_text | tail -1
