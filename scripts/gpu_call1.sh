mkdir -p gpurun_out
(timeout 200 python scripts/dmma_rate.py > gpurun_out/dmma_rate.txt 2>&1; echo "rc=$?" >> gpurun_out/dmma_rate.txt)
(timeout 600 python -m pytest tests/test_gpu_core.py tests/test_gpu_large.py -x -q > gpurun_out/pytest_stage3.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_stage3.txt)
tail -5 gpurun_out/pytest_stage3.txt
(timeout 420 python scripts/matvec_paths.py --out gpurun_out/matvec_paths.md > gpurun_out/matvec_paths.log 2>&1; echo "rc=$?" >> gpurun_out/matvec_paths.log)
cat gpurun_out/matvec_paths.log | tail -25
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:stage3f_kernel -c 1 -f -o gpurun_out/r1_stage3f_D6_chi12 python bench.py --D 6 --chi 12 --no-cpu --no-sweep --steps 1 --warmup 1 --path 3 > gpurun_out/ncu_D6.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_D6.log)
tail -3 gpurun_out/ncu_D6.log
cat gpurun_out/dmma_rate.txt
