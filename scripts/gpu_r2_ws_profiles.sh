# Evidence run for the warp-specialised stage3f kernels (single GPU): launch list and full ncu captures; the D = 6 / D = 4
# reports are summarised on the box (gpurun_out/ comes back only below 64 MiB), the D = 8 report comes back whole.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ws_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-sweep --no-heisenberg --no-parity > gpurun_out/r2ws_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage3f_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2ws_stage3f_D8_chi16 python scripts/matvec_paths.py --paths 3 --sizes 8:16 --steps 1 > gpurun_out/r2ws_ncu_D8_chi16.log 2>&1
for sz in 7:16 6:16 5:16 4:16; do
  n=${sz%%:*}
  timeout 600 ncu --set full --clock-control none -k regex:stage3f_kernel --launch-skip 2 -c 2 -f -o /tmp/r2ws_D$n python scripts/matvec_paths.py --paths 3 --sizes $sz --steps 1 > gpurun_out/r2ws_ncu_D${n}_chi16.log 2>&1
  python scripts/ncu_summary.py report /tmp/r2ws_D$n.ncu-rep gpurun_out/r2ws_stage3f_D${n}_chi16_full.txt > /dev/null 2>&1
done
python scripts/sweep_bench.py --sizes 3x6,4x8 --cpu-max-D 0 > /dev/null 2>&1; python scripts/sweep_bench.py --sizes 3x6,4x8,6x8,8x8 --cpu-max-D 0 > gpurun_out/r2ws_sweep_small.txt 2>&1
cut -c1-330 gpurun_out/r2ws_sweep_small.txt
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r2ws_bench_n1.json 2> gpurun_out/r2ws_bench_n1.err
ls -la gpurun_out/r2ws*
