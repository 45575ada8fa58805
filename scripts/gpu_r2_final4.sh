mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2h_pytest_gpu.txt 2>&1; grep -E "passed|failed" gpurun_out/r2h_pytest_gpu.txt | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py) > gpurun_out/r2_multi_gpu_check_n2.txt 2>&1
grep -v "Warn\|warn\|^\s*$\|\*\*\*\|OMP_NUM" gpurun_out/r2_multi_gpu_check_n2.txt | tail -3 | cut -c1-200
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; cut -c1-200 gpurun_out/r2h_bench_n1.json; tail -4 gpurun_out/r2h_bench_n1.err
