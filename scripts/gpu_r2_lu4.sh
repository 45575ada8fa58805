mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q -x) > gpurun_out/r2c_pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2c_pytest_gpu.txt
python scripts/lu_bench.py 2592 8192 2>&1 | tail -2
python scripts/sweep_bench.py --sizes 3x6 --cpu-max-D 0 > /dev/null 2>&1
python scripts/sweep_bench.py --sizes 3x6,4x8,6x8,8x8 --cpu-max-D 0 2>&1 | cut -c1-330 | tee gpurun_out/r2c_sweep_small.txt
