# single-GPU: whole GPU suite, sweep timings at the small sizes, default bench line
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.txt
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_gpu.txt | cut -c1-220 | tail -30
python scripts/sweep_bench.py --sizes 3x6,4x8 --cpu-max-D 0 > gpurun_out/r2_sweep_small.txt 2>&1; python scripts/sweep_bench.py --sizes 3x6,4x8 --cpu-max-D 0 >> gpurun_out/r2_sweep_small.txt 2>&1
cut -c1-420 gpurun_out/r2_sweep_small.txt
