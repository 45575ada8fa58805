mkdir -p gpurun_out
for ws in 1 0; do
  echo "== CARC_S3F_WS=$ws"
  CARC_S3F_WS=$ws timeout 300 python scripts/matvec_paths.py --paths 3 --sizes 8:8,8:16 --out gpurun_out/r2_ws_$ws.md > gpurun_out/r2_ws_$ws.log 2>&1
  cat gpurun_out/r2_ws_$ws.md; tail -3 gpurun_out/r2_ws_$ws.log
done
timeout 600 python -m pytest tests/test_gpu_core.py tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -5
