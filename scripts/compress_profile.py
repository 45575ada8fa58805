import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carcassonne_b200 import synthetic, compression, _lib
from carcassonne_b200.data import DeviceData, gemm_hermitian, gemm, _empty
from carcassonne_b200.sparse import Identity

def t(label, fn, n=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); print(f"   {label}: {(time.perf_counter()-t0)/n*1e3:.1f} ms", flush=True); return r

chi, D = int(sys.argv[1]), int(sys.argv[2])
s = synthetic.device_system(chi, D)
s.contractTowards(0)
I = Identity()
corner_id = 0
cj = t("join corner", lambda: s.corners[corner_id][I].join((0, 1, 2), 3, 4, 5))
sj = t("join side", lambda: s.sides[corner_id][I].join(0, 1, 2, (3, 4, 5, 6, 7)))
print("L", cj.shape, "R", sj.shape)
form = t("GramForm (LL0, RR0, T)", lambda: compression._GramForm(cj, sj))
old = cj.shape[1]; o2 = old*old; r = sj.shape[3]
RR0 = _empty((o2, o2))
t("  RR0 hermitian gemm alone", lambda: gemm_hermitian(_lib.OP_J, _lib.OP_T, o2, r, sj._t, r, sj._t, r, RR0))
t("  RR0 full gemm", lambda: gemm(_lib.OP_J, _lib.OP_T, o2, o2, r, sj._t, r, sj._t, r, RR0))
np.random.seed(0)
c = DeviceData.newRandom(old, chi).unitize()
g, rhs = t("normal_equations (one ALS round)", lambda: form.normal_equations(c))
x = t("solve", lambda: compression._solve_normal_equations(g, rhs, 1e-10))
t("unitize", lambda: x.split(old, chi).unitize())
comp = t("computeProductCompressor total", lambda: compression.computeProductCompressor(cj, sj, chi))
t("project corner", lambda: s._project(s.corners[corner_id], 3, comp, False))
t("project side", lambda: s._project(s.sides[corner_id], 0, comp, True))
t("contractTowards(1)", lambda: s.contractTowards(1))
