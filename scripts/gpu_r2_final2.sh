# Final single-GPU evidence run of round 2: GPU suite, smoke, bench line, path table, solver micro-benchmarks.
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/r2f_pytest_gpu.txt 2>&1; tail -4 gpurun_out/r2f_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
(time timeout 600 python bench.py --steps 20 --warmup 5) > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; cut -c1-200 gpurun_out/r2f_bench_n1.json; tail -4 gpurun_out/r2f_bench_n1.err
timeout 600 python scripts/matvec_paths.py --paths 1,3,2,0 --sizes 2:4,3:9,4:8,4:16,5:16,6:16,7:16,8:16,9:8,10:8,11:6,12:6 --out gpurun_out/r2f_matvec_paths.md > gpurun_out/r2f_matvec_paths.log 2>&1; tail -3 gpurun_out/r2f_matvec_paths.log
python scripts/lu_bench.py 700 2592 4500 8192 > gpurun_out/r2f_lu_bench.txt 2>&1; LU_BENCH_SCIPY_MAX=0 python scripts/lu_bench.py 13122 20000 >> gpurun_out/r2f_lu_bench.txt 2>&1; cat gpurun_out/r2f_lu_bench.txt
python scripts/solve_bench.py > gpurun_out/r2f_solve_bench.txt 2>&1; cat gpurun_out/r2f_solve_bench.txt
python scripts/zgemm_bench.py > gpurun_out/r2f_zgemm_bench.txt 2>&1; cat gpurun_out/r2f_zgemm_bench.txt
