mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lu_panel_cluster_kernel --launch-skip 5 -c 1 -f -o gpurun_out/r2_lu_panel python scripts/lu_bench.py 8192 > gpurun_out/r2_lu_ncu2.log 2>&1
ncu -i gpurun_out/r2_lu_panel.ncu-rep --page source --csv > gpurun_out/r2_lu_panel_source.csv 2>/dev/null
python scripts/ncu_summary.py report gpurun_out/r2_lu_panel.ncu-rep gpurun_out/r2_lu_panel_full.txt | head -40
rm -f gpurun_out/r2_lu_panel.ncu-rep
