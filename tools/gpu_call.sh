set -u
mkdir -p gpurun_out
for sp in 0 1; do
  echo "spread $sp"
  for g in zipf_text source_corpus mixed; do
    BZ3_B200_CM_SPREAD=$sp timeout 300 python tools/stage_driver.py cm_dec_bwt 4 2 $g | tail -1
  done
done 2>&1 | tee gpurun_out/r2n_cm_spread_ab.log
