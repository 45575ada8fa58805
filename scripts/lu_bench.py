"""carc_lu_factor timing and check against scipy at a few sizes (pivots, reconstruction)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.linalg
from carcassonne_b200.data import DeviceData
from carcassonne_b200.utils import LUFactors

sizes = [int(a) for a in sys.argv[1:]] or [100, 700, 2592, 8192]
rng = np.random.default_rng(5)
for n in sizes:
    M = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    d = DeviceData.fromArray(M)
    lu = LUFactors(d)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lu = LUFactors(d)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    dt = min(ts)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x = lu.solve(DeviceData.fromArray(b)).toArray()
    res = np.linalg.norm(M @ x - b) / np.linalg.norm(b)
    msg = ""
    if n <= int(os.environ.get("LU_BENCH_SCIPY_MAX", "5000")):
        lu_ref, piv_ref = scipy.linalg.lu_factor(M)
        msg = " pivots equal scipy: %s, max |LU - scipy| %.2e" % (
            bool(np.array_equal(lu.piv.cpu().numpy(), piv_ref)), float(np.abs(lu.lu.toArray() - lu_ref).max()))
    print("n=%d: LU %.2f ms (%.2f TFLOP/s), solve residual %.2e%s" % (n, dt * 1e3, 8 / 3 * n ** 3 / dt / 1e12, res, msg), flush=True)
