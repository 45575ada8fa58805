mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2g_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r2g_pytest_gpu.txt | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; cut -c1-200 gpurun_out/r2g_bench_n1.json; tail -4 gpurun_out/r2g_bench_n1.err
