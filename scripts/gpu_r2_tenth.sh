mkdir -p gpurun_out
echo "== default"
timeout 600 python scripts/matvec_paths.py --paths 3 --sizes 3:6,4:8,5:8,6:8,7:8,8:8 --out gpurun_out/r2_small_default.md > gpurun_out/r2_small_default.log 2>&1
cat gpurun_out/r2_small_default.md; tail -2 gpurun_out/r2_small_default.log
echo "== CARC_S3F_PB=1"
CARC_S3F_PB=1 timeout 600 python scripts/matvec_paths.py --paths 3 --sizes 6:8 --out gpurun_out/r2_small_pb1.md > gpurun_out/r2_small_pb1.log 2>&1
cat gpurun_out/r2_small_pb1.md
python scripts/sweep_bench.py --sizes 6x8 --cpu-max-D 0 2>&1 | tail -2
CARC_S3F_PB=1 python scripts/sweep_bench.py --sizes 6x8 --cpu-max-D 0 2>&1 | tail -2
