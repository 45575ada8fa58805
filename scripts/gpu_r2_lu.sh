CARC_LU_CLUSTER=0 python scripts/lu_bench.py 2592 8192 2>&1 | tail -2
python scripts/lu_bench.py 100 700 2592 4500 8192 2>&1 | tail -5
for ob in 128 512; do echo "OB=$ob"; CARC_LU_OUTER=$ob python scripts/lu_bench.py 2592 8192 2>&1 | tail -2; done
LU_BENCH_SCIPY_MAX=0 python scripts/lu_bench.py 13122 20000 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_system.py -m gpu -x -q -k "lu or cholesky or relax" 2>&1 | tail -2
