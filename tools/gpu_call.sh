set -u
mkdir -p gpurun_out
for f in 0 1; do
  echo "decoder exact tier $f"
  for g in zipf_text source_corpus mixed; do
    BZ3_B200_CM_DEC_EXACT=$f timeout 300 python tools/stage_driver.py cm_dec_bwt 4 2 $g | tail -1
  done
done 2>&1 | tee gpurun_out/r2s_cm_dec_exact_ab.log
BZ3_B200_CM_DEC_EXACT=1 BZ3_B200_CM_ENC_FAST=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stage_cm or block or hostile" 2>&1 | tail -2
