mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_core.py tests/test_gpu_large.py -x -q -k stage3 > gpurun_out/pytest_stage3.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_stage3.txt)
tail -3 gpurun_out/pytest_stage3.txt
timeout 300 python scripts/matvec_paths.py --sizes ${SIZES:-3:9,5:16,6:16,7:16,8:16} 2>&1 | grep "| [13] | [13] |"
