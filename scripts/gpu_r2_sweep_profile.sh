mkdir -p gpurun_out
for sz in 8x8 6x8; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r2d_launches_sweep_$sz.csv python scripts/sweep_bench.py --sizes $sz --cpu-max-D 0 > gpurun_out/r2d_sweep_ncu_$sz.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2d_launches_sweep_$sz.csv gpurun_out/r2d_sweep_launches_$sz.txt | head -24
rm -f gpurun_out/r2d_launches_sweep_$sz.csv
done
