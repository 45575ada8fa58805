"""Wall-clock cost of the FIRST apply of a stage-3 operator per shape (lazy module loading, cudaFuncSetAttribute) next to
the steady-state cost, in a fresh process.  Run with CUDA_MODULE_LOADING=EAGER / LAZY to compare."""
import sys, time, os, torch
t00 = time.perf_counter()
sys.path.insert(0, '.')
from carcassonne_b200.data import DeviceData as DD
from carcassonne_b200.operator import Stage3Operator
torch.cuda.init(); torch.zeros(1, device='cuda'); torch.cuda.synchronize()
print("CUDA_MODULE_LOADING", os.environ.get("CUDA_MODULE_LOADING"), "import+init %.3f s" % (time.perf_counter() - t00), flush=True)
def run(D, X):
    A = DD(torch.randn(X, D, D, D, D, dtype=torch.complex128, device='cuda'))
    B = DD(torch.randn(X, D, D, D, D, dtype=torch.complex128, device='cuda'))
    v = DD(torch.randn(D, D, D, D, 2, dtype=torch.complex128, device='cuda'))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    op = Stage3Operator(v.shape)
    op.add_term(A, B, None)
    op.finalize()
    out = torch.empty_like(v._t)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ts = []
    for _ in range(4):
        t = time.perf_counter(); op.apply_raw(v._t, out); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    print("D=%d X=%d path %d: build %.1f ms, applies %s ms" % (D, X, op.path, 1e3 * (t1 - t0), ["%.2f" % (1e3 * t) for t in ts]), flush=True)
for D, X in ((3, 324), (3, 400), (4, 1024), (5, 1600), (6, 2304), (7, 3136), (8, 4096), (3, 324)):
    run(D, X)
