CUDA_MODULE_LOADING=LAZY python scripts/first_launch.py 2>&1 | tail -12
CUDA_MODULE_LOADING=EAGER python scripts/first_launch.py 2>&1 | tail -12
for i in 1 2; do python scripts/sweep_bench.py --sizes 3x6,4x8 --cpu-max-D 0 2>&1 | cut -c1-200; done
