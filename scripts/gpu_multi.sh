mkdir -p gpurun_out
N=${N:-2}
(timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.txt)
tail -3 gpurun_out/pytest_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-400; tail -2 gpurun_out/bench_n$N.err
