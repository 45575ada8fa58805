# ncu evidence for the headline workload: launch list of the default bench command + one full capture of the top kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-sweep > gpurun_out/launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage3f_kernel -c 1 -f -o gpurun_out/r1_stage3f_D8_chi16 python bench.py --no-cpu --no-sweep --steps 1 --warmup 1 > gpurun_out/ncu_D8.log 2>&1
tail -2 gpurun_out/ncu_D8.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage3f_kernel -c 1 -f -o gpurun_out/r1_stage3f_D6_chi16 python bench.py --D 6 --no-cpu --no-sweep --steps 1 --warmup 1 > gpurun_out/ncu_D6.log 2>&1
tail -2 gpurun_out/ncu_D6.log
timeout 600 python scripts/matvec_paths.py --sizes 2:4,3:9,4:8,4:16,5:16,6:16,7:16,8:16 --out gpurun_out/matvec_paths.md > gpurun_out/matvec_paths.log 2>&1
cat gpurun_out/matvec_paths.md
