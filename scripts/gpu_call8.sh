mkdir -p gpurun_out
timeout 300 python scripts/sweep_bench.py --sizes 3x6,4x8 --cpu-max-D 0 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/launches_sweep_D3.csv python scripts/sweep_bench.py --sizes 3x6 --cpu-max-D 0 > gpurun_out/sweep_ncu.log 2>&1
tail -2 gpurun_out/sweep_ncu.log | cut -c1-300
