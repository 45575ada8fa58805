mkdir -p gpurun_out
for skew in 0 100; do
  echo "== CARC_S3F_SKEW=$skew"
  CARC_S3F_SKEW=$skew timeout 600 python scripts/matvec_paths.py --paths 3 --sizes 4:16,6:16,8:16 --out gpurun_out/r2_skew2_$skew.md > gpurun_out/r2_skew2_$skew.log 2>&1
  cat gpurun_out/r2_skew2_$skew.md
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage3f_kernel -c 1 -f -o gpurun_out/r2_stage3f_D8_chi8 python scripts/matvec_paths.py --paths 3 --sizes 8:8 --steps 1 > gpurun_out/r2_ncu_D8.log 2>&1
tail -3 gpurun_out/r2_ncu_D8.log
ls -la gpurun_out/*.ncu-rep | tail -3
