mkdir -p gpurun_out
for pb in 1 2 3; do
  echo "== CARC_S3F_PB=$pb"
  CARC_S3F_PB=$pb timeout 600 python scripts/matvec_paths.py --paths 3 --sizes 3:9,4:16,5:16,6:16,7:16,8:16 --out gpurun_out/r2_ws_pb$pb.md > gpurun_out/r2_ws_pb$pb.log 2>&1
  cat gpurun_out/r2_ws_pb$pb.md | sort -u; tail -3 gpurun_out/r2_ws_pb$pb.log
done
echo "== D=12 fused vs unfused"
timeout 600 python scripts/matvec_paths.py --paths 2,3 --sizes 12:6 --out gpurun_out/r2_ws_d12.md > gpurun_out/r2_ws_d12.log 2>&1
cat gpurun_out/r2_ws_d12.md | sort -u; tail -3 gpurun_out/r2_ws_d12.log
