# Round-2 correctness + first numbers on N GPUs of one box (N = 2 by default): GPU test suite (incl. the torchrun
# multi-GPU checks), the default bench line at N = 1 and the sharded one at N.
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2_gpus.txt
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.txt
tail -5 gpurun_out/r2_pytest_gpu.txt
(time timeout 600 python bench.py --steps 10 --warmup 3) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?" >> gpurun_out/r2_bench_n1.err
cut -c1-600 gpurun_out/r2_bench_n1.json; tail -8 gpurun_out/r2_bench_n1.err
(time timeout 300 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cut -c1-300 gpurun_out/r2_bench_ref.json
if [ "$N" -gt 1 ]; then
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?" >> gpurun_out/r2_bench_n$N.err
cut -c1-600 gpurun_out/r2_bench_n$N.json; tail -8 gpurun_out/r2_bench_n$N.err
fi
