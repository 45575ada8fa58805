mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_core.py tests/test_gpu_large.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py) > gpurun_out/r2_multi_gpu_check_n2.txt 2>&1
grep -v "Warn\|warn\|^\s*$\|\*\*\*\|OMP_NUM" gpurun_out/r2_multi_gpu_check_n2.txt | tail -4 | cut -c1-200
python scripts/matvec_paths.py --paths 3 --sizes 3:9,4:8,4:16,6:8,6:16,8:8,8:16 --out /tmp/p.md > /tmp/p.log 2>&1; sort -u /tmp/p.md | grep -v "^| D\|^|--"
