"""Parity at benchmark scale through size-independent properties (the oracle cannot run there): linearity of the
expectation matvec, additivity over X slabs (what the multi-GPU sharding relies on), agreement of the fused kernel
with the unfused DMMA-GEMM path, and oracle agreement on a slab small enough for the CPU."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tfim_like_terms():
    Z = np.diag([1.0, -1.0]).astype(complex)
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    return [(1, 0, None), (0, 1, None), (0, 0, -Z), (4, 0, X), (5, 0, -X), (0, 4, -X), (0, 5, X), (3, 2, None), (2, 3, None)]


def _build(D, X, seed=0):
    from carcassonne_b200.data import DeviceData
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)

    def rnd(*shape):
        t = torch.empty(shape, dtype=torch.complex128, device="cuda")
        torch.view_as_real(t).normal_(generator=gen)
        return t

    scale = 1.0 / (D * D * np.sqrt(X))
    A = [DeviceData(rnd(X, D, D, D, D).mul_(scale)) for _ in range(6)]
    B = [DeviceData(rnd(X, D, D, D, D).mul_(scale)) for _ in range(6)]
    return A, B, rnd


def _operator(A, B, D, lo=None, hi=None, path=0):
    from carcassonne_b200.data import DeviceData
    from carcassonne_b200.operator import Stage3Operator
    op = Stage3Operator((D, D, D, D, 2))
    for a, b, o in _tfim_like_terms():
        ta = A[a] if lo is None else DeviceData(A[a]._t[lo:hi])
        tb = B[b] if lo is None else DeviceData(B[b]._t[lo:hi])
        op.add_term(ta, tb, o)
    return op.finalize().set_path(path)


def _relerr(a, b):
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("D,chi", [(8, 12), (6, 16), (4, 24), (7, 12), (5, 16)])
def test_matvec_properties_at_scale(D, chi):
    from carcassonne_b200.data import DeviceData
    X = chi ** 4
    need = 12 * 16 * X * D ** 4 * 1.3
    if torch.cuda.mem_get_info()[0] < need:
        pytest.skip("not enough free device memory")
    A, B, rnd = _build(D, X)
    op = _operator(A, B, D)
    v1, v2 = rnd(D, D, D, D, 2), rnd(D, D, D, D, 2)
    a, b = 0.7 - 0.3j, -1.1 + 0.4j
    h1, h2 = op(DeviceData(v1))._t, op(DeviceData(v2))._t
    h12 = op(DeviceData(a * v1 + b * v2))._t
    assert _relerr(h12, a * h1 + b * h2) < 1e-12                       # linearity
    assert torch.equal(op(DeviceData(v1))._t, h1)                      # deterministic
    # additivity over X slabs (ragged split): the identity the multi-GPU sharding relies on
    cut = X // 3 + 1
    parts = _operator(A, B, D, 0, cut)(DeviceData(v1))._t + _operator(A, B, D, cut, X)(DeviceData(v1))._t
    assert _relerr(parts, h1) < 1e-12
    # fused kernel vs the unfused DMMA-GEMM path, and vs the CPU oracle, on a slab the oracle can handle
    lo, hi = X // 2, X // 2 + 48
    fused = _operator(A, B, D, lo, hi, path=1)(DeviceData(v1))._t
    unfused = _operator(A, B, D, lo, hi, path=2)(DeviceData(v1))._t
    folded = _operator(A, B, D, lo, hi, path=3)(DeviceData(v1))._t
    assert _relerr(fused, unfused) < 1e-12
    assert _relerr(folded, unfused) < 1e-12
    # both tilings of the fused kernel on the whole environment (whichever one the automatic choice did not take)
    assert _relerr(_operator(A, B, D, path=1)(DeviceData(v1))._t, h1) < 1e-12
    assert _relerr(_operator(A, B, D, path=3)(DeviceData(v1))._t, h1) < 1e-12
    from oracle import dense
    vh = v1.cpu().numpy()
    ref = sum(dense.stage3_multiply_joined(A[x]._t[lo:hi].cpu().numpy(), B[y]._t[lo:hi].cpu().numpy(), vh, o)
              for x, y, o in _tfim_like_terms())
    assert np.linalg.norm(fused.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-12
