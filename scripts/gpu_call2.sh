mkdir -p gpurun_out
for cfg in "5 12 3 stage3f_kernel" "5 12 1 stage3_kernel" "7 12 3 stage3f_kernel" "6 12 1 stage3_kernel"; do
  set -- $cfg
  (timeout 300 ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 -f -o gpurun_out/r1_D$1_chi$2_path$3 python bench.py --D $1 --chi $2 --no-cpu --no-sweep --steps 1 --warmup 1 --path $3 > gpurun_out/ncu_D$1_p$3.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_D$1_p$3.log)
  tail -2 gpurun_out/ncu_D$1_p$3.log
done
