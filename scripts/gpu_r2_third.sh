# Round 2, third GPU call: new parity tests (all), start-up skew experiment on the folded fused kernel.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_runs.py tests/test_gpu_compression_lossy.py tests/test_gpu_multi.py -q) > gpurun_out/r2_pytest_new.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_new.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_new.txt | tail -30
for skew in 0 50 100 150; do
  echo "== CARC_S3F_SKEW=$skew"
  CARC_S3F_SKEW=$skew timeout 600 python scripts/matvec_paths.py --paths 3 --sizes 3:9,5:16,6:16,7:16,8:16 --out gpurun_out/r2_skew_$skew.md > gpurun_out/r2_skew_$skew.log 2>&1
  cat gpurun_out/r2_skew_$skew.md
done
