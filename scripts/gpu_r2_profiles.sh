# Round-2 evidence run (single GPU): whole GPU suite, small-size sweep timings, ncu launch lists and full captures.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.txt
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_gpu.txt | cut -c1-220 | tail -12
python scripts/sweep_bench.py --sizes 3x6,4x8 --cpu-max-D 0 > /dev/null 2>&1; python scripts/sweep_bench.py --sizes 3x6,4x8,6x8 --cpu-max-D 0 > gpurun_out/r2_sweep_small.txt 2>&1
cut -c1-330 gpurun_out/r2_sweep_small.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-sweep --no-heisenberg > gpurun_out/r2_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage3f_kernel -c 1 -f -o gpurun_out/r2_stage3f_D8_chi16 python scripts/matvec_paths.py --paths 3 --sizes 8:16 --steps 1 > gpurun_out/r2_ncu_D8_chi16.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:stage3f_kernel -c 2 -f -o gpurun_out/r2_stage3f_D6_chi16 python scripts/matvec_paths.py --paths 3 --sizes 6:16 --steps 1 > gpurun_out/r2_ncu_D6_chi16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2_launches_sweep_D3.csv python scripts/sweep_bench.py --sizes 3x6 --cpu-max-D 0 > gpurun_out/r2_sweep_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_*.csv | tail -6
