mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_two_site.py tests/test_gpu_system.py -q -k "compression or two_site or edge" 2>&1 | tail -40) > gpurun_out/r2_pytest_opcomp.txt; grep -E "passed|failed|FAILED|^E  " gpurun_out/r2_pytest_opcomp.txt | cut -c1-250 | tail -30
