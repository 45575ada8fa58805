mkdir -p gpurun_out
echo "== default row blocks"
timeout 600 python scripts/matvec_paths.py --paths 3 --sizes 5:16,6:16,7:16,8:16,9:8,10:8,11:6 --out gpurun_out/r2_pb_default.md > gpurun_out/r2_pb_default.log 2>&1
cat gpurun_out/r2_pb_default.md; tail -2 gpurun_out/r2_pb_default.log
for pb in 1 2 3; do
echo "== CARC_S3F_PB=$pb"
CARC_S3F_PB=$pb timeout 600 python scripts/matvec_paths.py --paths 3 --sizes 6:16,8:16,9:8,10:8 --out gpurun_out/r2_pb_$pb.md > gpurun_out/r2_pb_$pb.log 2>&1
cat gpurun_out/r2_pb_$pb.md; tail -2 gpurun_out/r2_pb_$pb.log
done
