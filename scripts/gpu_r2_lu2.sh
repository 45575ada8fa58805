mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_lu_launches.csv python scripts/lu_bench.py 8192 > gpurun_out/r2_lu_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2_lu_launches.csv gpurun_out/r2_lu_launches_n8192.txt | head -14
