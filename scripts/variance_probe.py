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

chi, D = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 6)
s = synthetic.device_system(chi, D)
for d in range(4):
    s.contractTowards(d)
    for c in range(4):
        for d2 in range(2):
            s.compressCornerStateTowards(c, d2, chi)
from carcassonne_b200.tensors._2d import sparse as _sp
def build_env():
    _sp.environment_cache.clear()
    return s.formExpectationAndNormalizationMultipliers()
dist("environment build (stage 1 + 2 + plan, cache cleared)", build_env, 5)
H, N = s.formExpectationAndNormalizationMultipliers()
print("terms", len(H.terms), "groups", H.device_operator.num_groups, "dimension", H.shape[0], flush=True)
dist("N formMatrix", lambda: N.formMatrix(), 5)
M = N.formMatrix()
lu = LUFactors(M)
v = s.state_center_data
b = v.ravel()
dist("H matvec", lambda: H(v))
dist("N matvec", lambda: N(v))
dist("LU solve (wavefront)", lambda: lu.solve(b))
dist("LU solve (reference)", lambda: lu.solve_reference(b))
dist("LU factor", lambda: LUFactors(M), 10)
print("cholesky path:", LUFactors(M, try_cholesky=True).method)
dist("Cholesky factor (LU form)", lambda: LUFactors(M, try_cholesky=True), 10)

# phase breakdown of minimizeExpectation: environment + plan, N matrix, factorisation, relax
from carcassonne_b200 import utils as _u
phase = {}
def timed(name, fn):
    def wrapper(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); phase[name] = phase.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return r
    return wrapper
_orig_init = _u.LUFactors.__init__
_u.LUFactors.__init__ = timed("factor", _orig_init)
_orig_form = _u.Multiplier.formMatrix if hasattr(_u.Multiplier, "formMatrix") else None
s.formExpectationAndNormalizationMultipliers = timed("environment", s.formExpectationAndNormalizationMultipliers)
from carcassonne_b200 import _lib
_orig_relax = _lib.lib.carc_relax
class _Lib:
    def __getattr__(self, k):
        return timed("relax", _orig_relax) if k == "carc_relax" else getattr(_lib.lib, k)
import carcassonne_b200._lib as _L
_L.lib = _Lib()
for i in range(8):
    st = {}
    phase.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter(); s.minimizeExpectation(statistics=st); torch.cuda.synchronize()
    print("minimize %d: %.1f ms" % (i, (time.perf_counter() - t0) * 1e3), {k: st[k] for k in ("multiplications", "normalization")},
          {k: round(v, 1) for k, v in phase.items()}, flush=True)
