import sys, time, numpy as np, torch, ctypes as C
sys.path.insert(0, '.')
from carcassonne_b200.data import DeviceData as DD
from carcassonne_b200.operator import Stage3Operator
from carcassonne_b200 import _lib
tf = C.c_double(); _lib.check(_lib.lib.carc_dmma_peak(4000, C.byref(tf), None)); print("DMMA peak TF/s", tf.value)
def run(D, X, terms=1, path=1):
    P = D*D
    A = [DD(torch.randn(X, D, D, D, D, dtype=torch.complex128, device='cuda')) for _ in range(terms)]
    B = [DD(torch.randn(X, D, D, D, D, dtype=torch.complex128, device='cuda')) for _ in range(terms)]
    v = DD(torch.randn(D, D, D, D, 2, dtype=torch.complex128, device='cuda'))
    op = Stage3Operator(v.shape)
    for a, b in zip(A, B): op.add_term(a, b, None)
    op.finalize().set_path(path)
    out = torch.empty_like(v._t)
    for _ in range(3): op.apply_raw(v._t, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n): op.apply_raw(v._t, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/n
    flops = 8*op.cost_of_multiply
    byts = terms*2*X*P*P*16
    print(f"D={D} X={X} T={terms} path={path}: {ms:.3f} ms  {flops/ms/1e9:.1f} TFLOP/s  {byts/ms/1e6:.1f} GB/s algorithmic", flush=True)
for path in (1, 2):
    run(8, 4096, 1, path); run(8, 16384, 2, path); run(4, 65536, 1, path); run(6, 16384, 1, path); run(2, 65536*4, 1, path); run(3, 65536, 1, path)
