import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carcassonne_b200 import synthetic
from carcassonne_b200.utils import LUFactors

def dist(label, fn, n=40):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    print(f"{label}: min {ts[0]:.3f} med {ts[len(ts)//2]:.3f} p90 {ts[int(len(ts)*0.9)]:.3f} max {ts[-1]:.3f} ms", flush=True)

chi, D = 8, 6
s = synthetic.device_system(chi, D)
for d in range(4):
    s.contractTowards(d)
    for c in range(4):
        for d2 in range(2):
            s.compressCornerStateTowards(c, d2, chi)
H, N = s.formExpectationAndNormalizationMultipliers()
M = N.formMatrix()
lu = LUFactors(M)
v = s.state_center_data
b = v.ravel()
dist("H matvec", lambda: H(v))
dist("N matvec", lambda: N(v))
dist("LU solve (wavefront)", lambda: lu.solve(b))
dist("LU solve (reference)", lambda: lu.solve_reference(b))
dist("LU factor", lambda: LUFactors(M), 10)
dist("minimizeExpectation", lambda: s.minimizeExpectation(), 8)
for i in range(10):
    st = {}
    torch.cuda.synchronize(); t0 = time.perf_counter(); s.minimizeExpectation(statistics=st); torch.cuda.synchronize()
    print("minimize %d: %.1f ms" % (i, (time.perf_counter() - t0) * 1e3), {k: st[k] for k in ("counted", "multiplications", "normalization", "gmres_iterations")}, "ritz %.12g" % st["ritz_value"].real, flush=True)
