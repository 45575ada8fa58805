mkdir -p gpurun_out
cat > /tmp/z.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from carcassonne_b200.data import gemm, _empty
M = N = K = 4096
A = _empty((M, K)); B = _empty((K, N)); C = _empty((M, N))
torch.view_as_real(A).normal_(); torch.view_as_real(B).normal_()
for _ in range(4): gemm(0, 0, M, N, K, A, K, B, N, C)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none -k regex:zgemm_kernel --launch-skip 2 -c 1 -f -o /tmp/zg python /tmp/z.py > gpurun_out/r2_zgemm_ncu.log 2>&1
python scripts/ncu_summary.py report /tmp/zg.ncu-rep gpurun_out/r2_zgemm_ws_4096_full.txt | head -34
