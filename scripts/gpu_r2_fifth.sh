mkdir -p gpurun_out
timeout 900 python scripts/matvec_paths.py --paths 3,0 --sizes 2:4,3:9,4:8,4:16,5:16,6:16,7:16,8:16 --out gpurun_out/r2_matvec_na.md > gpurun_out/r2_matvec_na.log 2>&1
cat gpurun_out/r2_matvec_na.md; tail -3 gpurun_out/r2_matvec_na.log
(timeout 600 python -m pytest tests/test_gpu_core.py tests/test_gpu_large.py -q 2>&1 | tail -5) > gpurun_out/r2_pytest_core.txt; cat gpurun_out/r2_pytest_core.txt
