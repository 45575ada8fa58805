mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-sweep --no-heisenberg --no-parity > gpurun_out/r2_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage3f_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2_stage3f_D8_chi16 python scripts/matvec_paths.py --paths 3 --sizes 8:16 --steps 1 > gpurun_out/r2_ncu_D8_chi16.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:stage3f_kernel --launch-skip 2 -c 2 -f -o gpurun_out/r2_stage3f_D6_chi16 python scripts/matvec_paths.py --paths 3 --sizes 6:16 --steps 1 > gpurun_out/r2_ncu_D6_chi16.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:stage3f_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2_stage3f_D4_chi16 python scripts/matvec_paths.py --paths 3 --sizes 4:16 --steps 1 > gpurun_out/r2_ncu_D4_chi16.log 2>&1
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err
cut -c1-200 gpurun_out/r2_bench_final_n1.json; tail -4 gpurun_out/r2_bench_final_n1.err
(time timeout 400 python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/r2_bench_final_ref.json 2> gpurun_out/r2_bench_final_ref.err
cut -c1-300 gpurun_out/r2_bench_final_ref.json; tail -4 gpurun_out/r2_bench_final_ref.err
timeout 600 python scripts/matvec_paths.py --paths 1,3,2,0 --sizes 2:4,3:9,4:8,4:16,5:16,6:16,7:16,8:16,9:8,10:8,11:6,12:6 --out gpurun_out/r2_matvec_paths.md > gpurun_out/r2_matvec_paths.log 2>&1
cat gpurun_out/r2_matvec_paths.md | tail -50
