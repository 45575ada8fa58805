mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -5
for ws in 1 0; do
  echo "== CARC_S3F_WS=$ws"
  CARC_S3F_WS=$ws timeout 600 python scripts/matvec_paths.py --paths 3 --sizes 3:9,4:16,5:16,6:16,7:16,8:16,9:8,10:8,11:6 --out gpurun_out/r2_ws_$ws.md > gpurun_out/r2_ws_$ws.log 2>&1
  cat gpurun_out/r2_ws_$ws.md | sort -u; tail -3 gpurun_out/r2_ws_$ws.log
done
