mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.txt
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_gpu.txt | cut -c1-220 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
