timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_system.py tests/test_gpu_runs.py -m gpu -x -q 2>&1 | tail -3
python scripts/sweep_bench.py --sizes 3x6 --cpu-max-D 0 > /dev/null 2>&1
python scripts/sweep_bench.py --sizes 3x6,4x8,6x8,8x8 --cpu-max-D 0 2>&1 | cut -c1-260
