mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2_pytest_gpu_n2.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu_n2.txt
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_gpu_n2.txt | cut -c1-220 | tail -12
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --capacity) > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?" >> gpurun_out/r2_bench_n2.err
cut -c1-200 gpurun_out/r2_bench_n2.json; tail -4 gpurun_out/r2_bench_n2.err
(time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 scripts/sweep_couplings.py --J-grid 0.05:0.7:8) > gpurun_out/r2_sweep_couplings_n2.json 2> gpurun_out/r2_sweep_couplings_n2.err
cut -c1-600 gpurun_out/r2_sweep_couplings_n2.json
(time timeout 300 python scripts/sweep_couplings.py --J-grid 0.05:0.7:8) > gpurun_out/r2_sweep_couplings_n1.json 2> gpurun_out/r2_sweep_couplings_n1.err
cut -c1-300 gpurun_out/r2_sweep_couplings_n1.json
