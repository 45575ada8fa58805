# Round 2, second GPU call: the new parity tests, fused / unfused matvec numbers over D = 2..12, per-warp wait profile.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_runs.py tests/test_gpu_compression_lossy.py tests/test_gpu_multi.py -q -x) > gpurun_out/r2_pytest_new.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_new.txt
tail -30 gpurun_out/r2_pytest_new.txt
timeout 900 python scripts/matvec_paths.py --sizes 2:4,3:9,4:8,4:16,5:16,6:16,7:16,8:16,9:8,10:8,12:6 --out gpurun_out/r2_matvec_paths.md > gpurun_out/r2_matvec_paths.log 2>&1
cat gpurun_out/r2_matvec_paths.md
(cd carcassonne_b200/csrc && rm -f stage3f.o && make EXTRA=-DS3F_PROFILE -j8 > /dev/null 2>&1)
timeout 600 python scripts/s3f_waits.py 8:16,8:8,6:16,7:16,5:16 > gpurun_out/r2_s3f_waits.txt 2>&1
cat gpurun_out/r2_s3f_waits.txt
