"""zgemm_kernel rate at a few shapes (CARC_ZGEMM_WS=0 runs the symmetric kernel) next to torch.matmul (cuBLAS ZGEMM)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from carcassonne_b200.data import gemm, _empty

def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3

print("CARC_ZGEMM_WS =", os.environ.get("CARC_ZGEMM_WS", "1 (default)"))
for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 512), (8192, 8192, 256), (8192, 8192, 64), (4096, 4096, 64), (512, 512, 262144), (8192, 192, 64), (256, 8192, 64)]:
    A = _empty((M, K)); B = _empty((K, N)); C = _empty((M, N))
    torch.view_as_real(A).normal_(); torch.view_as_real(B).normal_()
    dt = t(lambda: gemm(0, 0, M, N, K, A, K, B, N, C))
    dt2 = t(lambda: torch.matmul(A, B, out=C))
    ref = torch.matmul(A, B)
    gemm(0, 0, M, N, K, A, K, B, N, C)
    err = float((C - ref).abs().max() / ref.abs().max())
    print("zgemm %6d x %6d x %6d: %8.3f ms %5.1f TF/s | cuBLAS %8.3f ms %5.1f TF/s | max rel diff %.1e" % (
        M, N, K, dt * 1e3, 8 * M * N * K / dt / 1e12, dt2 * 1e3, 8 * M * N * K / dt2 / 1e12, err), flush=True)
    del A, B, C, ref
