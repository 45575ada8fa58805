mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_core.py tests/test_gpu_large.py -x -q -k stage3 > gpurun_out/pytest_stage3.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_stage3.txt)
tail -3 gpurun_out/pytest_stage3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stage3f_kernel -c 1 -f -o gpurun_out/r1_stage3f_D8_chi16 python bench.py --no-cpu --no-sweep --steps 1 --warmup 1 > gpurun_out/ncu_D8.log 2>&1
tail -2 gpurun_out/ncu_D8.log
timeout 300 python scripts/matvec_paths.py --sizes 5:16,6:16,7:16,8:16 2>&1 | grep "| [13] | [13] |"
