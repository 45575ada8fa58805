mkdir -p gpurun_out
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/dmma_loops scripts/dmma_loops.cu && /tmp/dmma_loops > gpurun_out/r2_dmma_loops.txt 2>&1
cat gpurun_out/r2_dmma_loops.txt
timeout 300 python scripts/trace_runs.py > gpurun_out/r2_trace_runs.txt 2>&1; cat gpurun_out/r2_trace_runs.txt | tail -30
(time timeout 900 python -m pytest tests/test_gpu_runs.py tests/test_gpu_compression_lossy.py tests/test_gpu_multi.py -q) > gpurun_out/r2_pytest_new.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_new.txt
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_new.txt | cut -c1-200 | tail -30
