mkdir -p gpurun_out
N=${N:-2}
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py) > gpurun_out/r2_multi_gpu_check_n$N.txt 2>&1
grep -v "Warn\|warn\|^\s*$\|\*\*\*\|OMP_NUM" gpurun_out/r2_multi_gpu_check_n$N.txt | tail -22 | cut -c1-250
