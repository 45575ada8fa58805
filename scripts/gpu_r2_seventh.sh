mkdir -p gpurun_out
for np in 0 1; do
  echo "== CARC_S3F_NP=$np"
  CARC_S3F_NP=$np CARC_S3F_SKEW=0 timeout 600 python scripts/matvec_paths.py --paths 3 --sizes 3:9,4:16,5:16,6:16,7:16,8:16 --out gpurun_out/r2_np_$np.md > gpurun_out/r2_np_$np.log 2>&1
  cat gpurun_out/r2_np_$np.md; tail -2 gpurun_out/r2_np_$np.log
done
CARC_S3F_SKEW=0 timeout 600 python scripts/matvec_paths.py --paths 3,2 --sizes 9:8,10:8,11:6,12:6 --out gpurun_out/r2_large_D.md > gpurun_out/r2_large_D.log 2>&1
cat gpurun_out/r2_large_D.md; tail -3 gpurun_out/r2_large_D.log
(timeout 600 python -m pytest tests/test_gpu_core.py -q 2>&1 | tail -8) > gpurun_out/r2_pytest_core.txt; cat gpurun_out/r2_pytest_core.txt
