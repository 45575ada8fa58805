mkdir -p gpurun_out
N=${N:-8}
(time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?" >> gpurun_out/r2_bench_n$N.err
cut -c1-300 gpurun_out/r2_bench_n$N.json; tail -6 gpurun_out/r2_bench_n$N.err
(time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/multi_gpu_check.py) > gpurun_out/r2_multi_gpu_check_n$N.txt 2>&1
tail -8 gpurun_out/r2_multi_gpu_check_n$N.txt
(time timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 scripts/sweep_couplings.py --J-grid 0.05:0.8:16) > gpurun_out/r2_sweep_couplings_n$N.json 2> gpurun_out/r2_sweep_couplings_n$N.err
cut -c1-300 gpurun_out/r2_sweep_couplings_n$N.json
