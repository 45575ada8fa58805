mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage3f_kernel -c 1 -f -o gpurun_out/r2b_stage3f_D8_chi8 python scripts/matvec_paths.py --paths 3 --sizes 8:8 --steps 1 > gpurun_out/r2b_ncu_D8.log 2>&1
tail -2 gpurun_out/r2b_ncu_D8.log
