"""Timing of the pieces of minimizeExpectation at one size: environment build, normalization matrix, LU, solves, matvecs."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carcassonne_b200 import synthetic
from carcassonne_b200.utils import LUFactors
from carcassonne_b200.data import DeviceData

def t(fn, n=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n, r

for chi, D in [(8, 4), (8, 6), (8, 8)]:
    s = synthetic.device_system(chi, D)
    for d in range(4):
        s.contractTowards(d)
        for c in range(4):
            for d2 in range(2):
                s.compressCornerStateTowards(c, d2, chi)
    dt, (H, N) = t(s.formExpectationAndNormalizationMultipliers)
    print(f"chi={chi} D={D}: form multipliers {dt*1e3:.1f} ms, terms {H.device_operator.num_terms} groups {H.device_operator.num_groups}")
    dt, M = t(N.formMatrix); n = M.shape[0]
    print(f"   N matrix n={n}: {dt*1e3:.1f} ms")
    dt, lu = t(lambda: LUFactors(M)); print(f"   LU factor: {dt*1e3:.1f} ms  ({8/3*n**3/dt/1e12:.2f} TFLOP/s)")
    v = s.state_center_data
    b = v.ravel()
    dt, _ = t(lambda: lu.solve(b), 5); print(f"   LU solve: {dt*1e3:.2f} ms")
    dt, _ = t(lambda: H(v), 5); print(f"   H matvec: {dt*1e3:.2f} ms ({8*H.cost_of_multiply/dt/1e12:.1f} TF/s ref-equivalent)")
    dt, _ = t(lambda: N(v), 5); print(f"   N matvec: {dt*1e3:.2f} ms")
    dt, _ = t(lambda: s.minimizeExpectation()); print(f"   minimizeExpectation: {dt*1e3:.1f} ms")
    dt, _ = t(lambda: s.contractTowards(0)); print(f"   contractTowards: {dt*1e3:.1f} ms")
    dt, _ = t(lambda: s.compressCornerStateTowards(0, 1, chi)); print(f"   compress (grown bond): {dt*1e3:.1f} ms")
    # raw GEMM rate
    for (M_, N_, K_) in [(4096, 4096, 4096), (4096, 4096, 64), (512, 512, 262144)]:
        from carcassonne_b200.data import gemm, _empty
        A_ = _empty((M_, K_)); B_ = _empty((K_, N_)); C_ = _empty((M_, N_))
        torch.view_as_real(A_).normal_(); torch.view_as_real(B_).normal_()
        dt, _ = t(lambda: gemm(0, 0, M_, N_, K_, A_, K_, B_, N_, C_), 3)
        print(f"   zgemm {M_}x{N_}x{K_}: {dt*1e3:.2f} ms {8*M_*N_*K_/dt/1e12:.1f} TF/s")
    del s, H, N, M, lu
    torch.cuda.empty_cache()
