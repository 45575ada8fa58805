import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from carcassonne_b200 import _lib
from carcassonne_b200.data import gemm, gemm_hermitian, _empty
o2, r = 4096, 65536
R = _empty((o2, r)); torch.view_as_real(R).normal_()
G = _empty((o2, o2))
for _ in range(2):
    gemm(_lib.OP_J, _lib.OP_T, o2, o2, r, R, r, R, r, G)
A = _empty((4096, 4096)); torch.view_as_real(A).normal_()
for _ in range(2):
    gemm(_lib.OP_N, _lib.OP_N, 4096, 4096, 4096, A, 4096, A, 4096, G)
torch.cuda.synchronize()
