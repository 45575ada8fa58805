mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.txt)
tail -4 gpurun_out/pytest_gpu.txt
timeout 400 python scripts/hbm_kernels.py --out gpurun_out/hbm_kernels.md 2>&1 | tail -10
