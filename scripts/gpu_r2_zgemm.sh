CARC_ZGEMM_WS=0 python scripts/zgemm_bench.py 2>&1 | tail -9
python scripts/zgemm_bench.py 2>&1 | tail -9
timeout 900 python -m pytest tests/test_gpu_core.py tests/test_gpu_linalg.py tests/test_gpu_system.py -m gpu -x -q 2>&1 | tail -3
python scripts/lu_bench.py 2592 8192 2>&1 | tail -2
