mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.txt ) 2> gpurun_out/pytest_time.txt
tail -4 gpurun_out/pytest_gpu.txt; cat gpurun_out/pytest_time.txt
( time timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?" >> gpurun_out/bench.err ) 2> gpurun_out/bench_time.txt
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_time.txt
