"""carc_lu_solve_blocks (permutation + two wavefront substitutions) timing and residual at a few sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carcassonne_b200.data import DeviceData
from carcassonne_b200.utils import LUFactors

rng = np.random.default_rng(7)
for n in [int(a) for a in sys.argv[1:]] or [162, 512, 2592, 8192]:
    M = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    lu = LUFactors(DeviceData.fromArray(M))
    db = DeviceData.fromArray(b)
    x = lu.solve(db).toArray()
    xr = lu.solve_reference(db).toArray()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): lu.solve(db)
    e1.record(); torch.cuda.synchronize()
    print("n=%d: solve %.3f ms, residual %.2e, vs plain substitution %.2e" % (
        n, e0.elapsed_time(e1) / 20, np.linalg.norm(M @ x - b) / np.linalg.norm(b), np.linalg.norm(x - xr) / np.linalg.norm(xr)), flush=True)
