mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.txt
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_gpu.txt | cut -c1-220 | tail -12
python scripts/sweep_bench.py --sizes 3x6,4x8 --cpu-max-D 0 > /dev/null 2>&1; python scripts/sweep_bench.py --sizes 3x6,4x8,6x8 --cpu-max-D 0 > gpurun_out/r2_sweep_small.txt 2>&1
cut -c1-330 gpurun_out/r2_sweep_small.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_sweep_D3.csv python scripts/sweep_bench.py --sizes 3x6 --cpu-max-D 0 > gpurun_out/r2_sweep_ncu.log 2>&1
tail -2 gpurun_out/r2_sweep_ncu.log | cut -c1-200
