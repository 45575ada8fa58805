mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_system.py tests/test_gpu_runs.py tests/test_gpu_1d.py -m gpu -x -q 2>&1 | tail -4
python scripts/sweep_bench.py --sizes 3x6 --cpu-max-D 0 > /dev/null 2>&1
for i in 1 2; do python scripts/sweep_bench.py --sizes 3x6,4x8 --cpu-max-D 0 2>&1 | cut -c1-230; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2b_launches_sweep_D3.csv python scripts/sweep_bench.py --sizes 3x6 --cpu-max-D 0 > gpurun_out/r2b_sweep_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2b_launches_sweep_D3.csv gpurun_out/r2b_sweep_launches_D3_chi6.txt | head -14
