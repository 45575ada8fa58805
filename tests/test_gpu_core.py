"""GPU parity tests of the CUDA core (permute, DMMA ZGEMM, tensordot, fused stage-3 matvec) against the CPU
oracle / NumPy on the same seeded inputs.  All calls go through the C ABI of libcarc_b200.so."""
import itertools

import numpy as np
import pytest

from golden_io import load, relerr

pytestmark = pytest.mark.gpu

MATVEC_TOL = 1e-12   # north_star: matvec output relative norm error <= 1e-12


def crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


@pytest.fixture(scope="module")
def dd():
    from carcassonne_b200.data import DeviceData
    return DeviceData


@pytest.mark.parametrize("shape,perm", [
    ((3, 4, 5), (2, 0, 1)),
    ((2, 3, 1, 4, 2), (4, 2, 0, 3, 1)),
    ((7,), (0,)),
    ((33, 65), (1, 0)),
    ((5, 40, 3, 36), (2, 3, 0, 1)),
    ((2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2), (11, 0, 9, 2, 7, 4, 5, 6, 3, 8, 1, 10)),
    ((4, 4, 1, 4, 4, 1, 3, 3), (1, 0, 2, 4, 3, 5, 7, 6)),
    # trailing axes rearranged among themselves (block path): Hermitian symmetrisation at D = 8, small state-bond swaps
    ((6, 5, 1, 4, 3, 1, 8, 8), (1, 0, 2, 4, 3, 5, 7, 6)),
    ((37, 3, 4, 5), (0, 3, 2, 1)),
    ((3, 7, 2, 3, 2, 3), (1, 0, 4, 5, 2, 3)),
    ((9, 11, 32, 32), (1, 0, 3, 2)),
    ((5, 2, 2, 2, 2, 2, 2), (0, 6, 5, 4, 3, 2, 1)),
])
def test_permute(dd, shape, perm):
    rng = np.random.default_rng(0)
    a = crand(rng, *shape)
    out = dd.fromArray(a).transpose(*perm).toArray()
    assert np.array_equal(out, a.transpose(perm))
    out = dd.fromArray(a)._permuted(perm, conj=1).toArray()
    assert np.array_equal(out, a.transpose(perm).conj())


def test_join_matches_reference_semantics(dd):
    rng = np.random.default_rng(1)
    a = crand(rng, 2, 3, 4, 5)
    out = dd.fromArray(a).join((2, 0), 3, 1).toArray()
    assert np.array_equal(out, a.transpose(2, 0, 3, 1).reshape(8, 5, 3))


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 7, 3), (128, 64, 16), (130, 70, 37), (64, 300, 129), (257, 33, 8), (100, 70, 4100)])
@pytest.mark.parametrize("opA,opB", list(itertools.product(range(4), range(4))))
def test_zgemm(dd, M, N, K, opA, opB):
    import torch
    from carcassonne_b200 import _lib
    from carcassonne_b200.data import gemm
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    a = crand(rng, M, K)
    b = crand(rng, K, N)
    c0 = crand(rng, M, N)
    a_st = {0: a, 1: a.T, 2: a.T.conj(), 3: a.conj()}[opA]   # what is stored so that op(stored) == a
    b_st = {0: b, 1: b.T, 2: b.T.conj(), 3: b.conj()}[opB]
    A = torch.from_numpy(np.ascontiguousarray(a_st)).cuda()
    B = torch.from_numpy(np.ascontiguousarray(b_st)).cuda()
    Cb = torch.from_numpy(c0.copy()).cuda()
    alpha, beta = 0.7 - 0.2j, -0.3 + 0.5j
    gemm(opA, opB, M, N, K, A, a_st.shape[1], B, b_st.shape[1], Cb, alpha=alpha, beta=beta)
    ref = alpha * (a @ b) + beta * c0
    assert relerr(Cb.cpu().numpy(), ref) < (1e-14 if K < 1000 else 1e-13)
    assert _lib.lib.carc_version() >= 100


@pytest.mark.parametrize("n,k", [(5, 7), (64, 100), (200, 33), (300, 1000), (130, 5000)])
def test_zgemm_hermitian(dd, n, k):
    """Gram matrices: only the upper tiles are computed, the lower triangle is their exact conjugate mirror."""
    import torch
    from carcassonne_b200 import _lib
    from carcassonne_b200.data import gemm_hermitian
    rng = np.random.default_rng(n + k)
    a = crand(rng, k, n)
    A = torch.from_numpy(a).cuda()
    G = torch.empty(n, n, dtype=torch.complex128, device="cuda")
    gemm_hermitian(_lib.OP_C, _lib.OP_N, n, k, A, n, A, n, G)           # a^H a
    g = G.cpu().numpy()
    assert relerr(g, a.conj().T @ a) < 1e-13
    off = ~np.eye(n, dtype=bool)
    if k < 1000:                                          # (long-K small outputs take the split-K path instead)
        assert np.array_equal(g[off], g.conj().T[off])    # the lower triangle is the exact mirror
    assert np.max(np.abs(np.diag(g).imag)) < 1e-13 * np.max(np.abs(g))
    r = crand(rng, n, k)
    R = torch.from_numpy(r).cuda()
    gemm_hermitian(_lib.OP_J, _lib.OP_T, n, k, R, k, R, k, G)           # conj(r) r^T
    assert relerr(G.cpu().numpy(), r.conj() @ r.T) < 1e-13


def test_zgemm_output_map_and_kmap(dd):
    import torch
    from carcassonne_b200 import _lib
    from carcassonne_b200.data import gemm
    rng = np.random.default_rng(5)
    X, P, Q, S, d, R = 5, 6, 7, 9, 2, 4
    A = crand(rng, X, P, Q)
    w = crand(rng, Q, S, d)
    Bm = crand(rng, X, R, S)
    At, wt, Bt = (torch.from_numpy(x).cuda() for x in (A, w, Bm))
    T = torch.empty(P, d, X, S, dtype=torch.complex128, device="cuda")
    gemm(_lib.OP_N, _lib.OP_N, X * P, S * d, Q, At, Q, wt, S * d, T,
         out_map=(P, S, d * X * S, d, 1, X * S))
    ref_T = np.einsum("xpq,qSs->psxS", A, w)
    assert relerr(T.cpu().numpy(), ref_T) < 1e-14
    out = torch.zeros(P, R, d, dtype=torch.complex128, device="cuda")
    gemm(_lib.OP_N, _lib.OP_T, P * d, R, X * S, T, X * S, Bt, S, out, beta=1.0,
         out_map=(d, R * d, 1, R, 0, d), k_map=(0, 0, S, R * S))
    ref = np.einsum("psxS,xrS->prs", ref_T, Bm)
    assert relerr(out.cpu().numpy(), ref) < 1e-14


@pytest.mark.parametrize("sa,sb,axa,axb", [
    ((3, 4, 5), (5, 4, 2), (1, 2), (1, 0)),
    ((6, 2, 3), (2, 3, 7), (1, 2), (0, 1)),
    ((2, 3, 4), (4, 3, 5, 2), (0, 1), (3, 1)),
    ((4, 4), (4, 4), (), ()),
    ((3, 1, 2), (2, 3, 1), (0, 1, 2), (1, 2, 0)),
])
def test_contract_with(dd, sa, sb, axa, axb):
    rng = np.random.default_rng(2)
    a, b = crand(rng, *sa), crand(rng, *sb)
    out = dd.fromArray(a).contractWith(dd.fromArray(b), axa, axb)
    ref = np.tensordot(a, b, (axa, axb))
    assert out.shape == ref.shape
    assert relerr(out.toArray(), ref) < 1e-14


def test_absorb_matrix_at(dd):
    g = load("data_ops")
    for n in range(4):
        t, axis, m = g["na%d_in" % n], int(g["na%d_axis" % n]), g["am%d_m" % n]
        out = dd.fromArray(t).absorbMatrixAt(axis, dd.fromArray(m))
        assert relerr(out.toArray(), g["am%d_out" % n]) < 1e-14


@pytest.mark.parametrize("shape,axis,rows", [
    ((37, 40), 0, 1), ((37, 40), 0, 5), ((13, 300), 0, 8), ((3, 11, 70), 1, 9), ((2, 64, 5, 33), 1, 16),
    ((5, 4, 257), 1, 3), ((6, 2, 1000), 0, 4), ((1, 130, 32), 1, 12),
])
def test_absorb_short_matrix_streaming_kernel(dd, shape, axis, rows):
    """absorbMatrixAt with <= 16 rows against a long tensor goes through carc_mode_product (the compressor projections
    of reference system/_2d.py:199-228); ragged k (not a multiple of 4) and post (not a multiple of the block)."""
    rng = np.random.default_rng(11)
    t = crand(rng, *shape)
    m = crand(rng, rows, shape[axis])
    out = dd.fromArray(t).absorbMatrixAt(axis, dd.fromArray(m)).toArray()
    ref = np.moveaxis(np.tensordot(m, t, (1, axis)), 0, axis)
    assert out.shape == ref.shape
    assert relerr(out, ref) < 1e-14


def test_mode_product_rejects_tall_matrix():
    import torch
    from carcassonne_b200 import _lib
    x = torch.zeros(64 * 64, dtype=torch.complex128, device="cuda")
    rc = _lib.lib.carc_mode_product(x.data_ptr(), x.data_ptr(), x.data_ptr(), 17, 4, 1, 64, None)
    assert rc == _lib.ERR_UNSUPPORTED


def test_elementwise_and_reductions(dd):
    rng = np.random.default_rng(3)
    a, b = crand(rng, 37, 5, 3), crand(rng, 37, 5, 3)
    A, B = dd.fromArray(a), dd.fromArray(b)
    assert relerr((A + B).toArray(), a + b) < 1e-15
    assert relerr((A - B).toArray(), a - b) < 1e-15
    assert relerr((A * B).toArray(), a * b) < 1e-15
    assert relerr((A * (0.5 - 2j)).toArray(), a * (0.5 - 2j)) < 1e-15
    assert relerr((-A).toArray(), -a) == 0
    assert relerr(A.conj().toArray(), a.conj()) == 0
    assert abs(A.norm() - np.linalg.norm(a)) < 1e-13 * np.linalg.norm(a)
    assert abs(A.contractWithAlongAll(B) - np.sum(a * b)) < 1e-12
    assert abs(A.conj().contractWithAlongAll(B) - np.vdot(a, b)) < 1e-12
    Cc = A.copy()
    Cc += B
    assert relerr(Cc.toArray(), a + b) < 1e-15
    assert not A.hasNaN()
    a2 = a.copy()
    a2[3, 2, 1] = np.nan
    assert dd.fromArray(a2).hasNaN()


def _stage3_case(rng, chi2, D, d=2, terms=1, dims=None):
    from oracle import dense
    d0, d1, d2, d3 = dims if dims else (D, D, D, D)
    s2_0 = [crand(rng, chi2, chi2, d0, d1, d0, d1) for _ in range(terms)]
    s2_1 = [crand(rng, chi2, chi2, d2, d3, d2, d3) for _ in range(terms)]
    ops = [None if t % 2 == 0 else crand(rng, d, d) for t in range(terms)]
    v = crand(rng, d0, d1, d2, d3, d)
    ref = sum(dense.stage3_multiply(a, b, v, o) for a, b, o in zip(s2_0, s2_1, ops))
    return s2_0, s2_1, ops, v, ref


@pytest.mark.parametrize("chi2,D,terms", [(2, 2, 1), (4, 2, 3), (3, 3, 2), (4, 4, 2), (9, 3, 4), (5, 5, 1), (4, 6, 2),
                                          (3, 7, 1), (6, 8, 2)])
@pytest.mark.parametrize("path", [1, 2, 3])
def test_stage3_operator(dd, chi2, D, terms, path):
    from carcassonne_b200.operator import Stage3Operator, prejoin_halves
    rng = np.random.default_rng(chi2 * 100 + D * 10 + terms)
    s2_0, s2_1, ops, v, ref = _stage3_case(rng, chi2, D, terms=terms)
    op = Stage3Operator(v.shape)
    for a, b, o in zip(s2_0, s2_1, ops):
        A, B = prejoin_halves(dd.fromArray(a), dd.fromArray(b))
        op.add_term(A, B, o)
    op.finalize().set_path(path)
    out = op(dd.fromArray(v)).toArray()
    assert relerr(out, ref) < MATVEC_TOL
    # linearity (size-independent property): op(2v + i v) == (2 + i) op(v)
    out2 = op(dd.fromArray((2 + 1j) * v)).toArray()
    assert relerr(out2, (2 + 1j) * ref) < MATVEC_TOL


@pytest.mark.parametrize("chi2,dims,terms", [(2, (9, 9, 9, 9), 2), (2, (10, 10, 10, 10), 3), (2, (11, 11, 11, 11), 1),
                                             (2, (12, 12, 12, 12), 2), (3, (9, 8, 5, 11), 2), (2, (3, 30, 2, 40), 2),
                                             (3, (12, 12, 4, 4), 2), (2, (5, 5, 10, 13), 3)])
@pytest.mark.parametrize("path", [0, 2, 3])
def test_stage3_large_bond_dimensions(dd, chi2, dims, terms, path):
    """Outputs beyond the 64 x 64 a single launch of the fused kernel covers (D = 9 .. 12, ragged bonds): computed in
    row / column blocks of the output; the reference's recipe (tensors/_2d/dense.py:115-160) has no such limit.  The
    automatic choice must be the fused kernel."""
    from carcassonne_b200.operator import Stage3Operator, prejoin_halves
    rng = np.random.default_rng(chi2 * 1000 + sum(dims) * 10 + terms)
    s2_0, s2_1, ops, v, ref = _stage3_case(rng, chi2, None, terms=terms, dims=dims)
    op = Stage3Operator(v.shape)
    for a, b, o in zip(s2_0, s2_1, ops):
        A, B = prejoin_halves(dd.fromArray(a), dd.fromArray(b))
        op.add_term(A, B, o)
    op.finalize().set_path(path)
    if path == 0:      # fused up to two column blocks of the output (R <= 128); beyond, the unfused GEMMs are faster
        assert op.path == (3 if dims[2] * dims[3] <= 128 else 2)
    out = op(dd.fromArray(v)).toArray()
    assert relerr(out, ref) < MATVEC_TOL
    out2 = op(dd.fromArray((2 + 1j) * v)).toArray()
    assert relerr(out2, (2 + 1j) * ref) < MATVEC_TOL


@pytest.mark.parametrize("dims,d", [((2, 3, 3, 2), 2), ((1, 1, 1, 1), 2), ((3, 2, 2, 4), 3), ((2, 2, 2, 2), 1),
                                     ((4, 4, 2, 2), 2), ((2, 2, 8, 8), 2), ((3, 5, 7, 1), 2), ((7, 7, 5, 3), 2),
                                     ((1, 5, 6, 6), 2)])
@pytest.mark.parametrize("path", [0, 3])
def test_stage3_ragged_shapes(dd, dims, d, path):
    from carcassonne_b200.operator import Stage3Operator, prejoin_halves
    if path == 3 and d != 2:
        pytest.skip("the folded tiling is specific to d = 2")
    rng = np.random.default_rng(sum(dims) + d)
    s2_0, s2_1, ops, v, ref = _stage3_case(rng, 3, None, d=d, terms=2, dims=dims)
    op = Stage3Operator(v.shape)
    for a, b, o in zip(s2_0, s2_1, ops):
        A, B = prejoin_halves(dd.fromArray(a), dd.fromArray(b))
        op.add_term(A, B, o)
    op.finalize().set_path(path)
    out = op(dd.fromArray(v)).toArray()
    assert relerr(out, ref) < MATVEC_TOL


def test_stage3_folded_rejects_other_physical_dimensions(dd):
    """Forcing the folded tiling on a shape outside its envelope is an error, not a silent fallback."""
    from carcassonne_b200 import _lib
    from carcassonne_b200.operator import Stage3Operator, prejoin_halves
    rng = np.random.default_rng(5)
    s2_0, s2_1, ops, v, ref = _stage3_case(rng, 2, 2, d=3, terms=1)
    op = Stage3Operator(v.shape)
    A, B = prejoin_halves(dd.fromArray(s2_0[0]), dd.fromArray(s2_1[0]))
    op.add_term(A, B, None).finalize().set_path(3)
    with pytest.raises(Exception):
        op(dd.fromArray(v))
    op.set_path(0)
    assert relerr(op(dd.fromArray(v)).toArray(), ref) < MATVEC_TOL


@pytest.mark.parametrize("chi2,D", [(3, 2), (4, 4), (5, 8), (2, 6), (4, 3), (3, 5), (3, 7)])
@pytest.mark.parametrize("path", [0, 1, 2, 3])
def test_stage3_shared_tensors(dd, chi2, D, path):
    """The TFIM term structure of bench.py: 9 terms over 6 + 6 tensors, several sharing the same half-1 tensor (their
    first products are summed before one second product) and the same half-0 tensor."""
    from carcassonne_b200.operator import Stage3Operator, prejoin_halves
    from oracle import dense
    rng = np.random.default_rng(chi2 * 10 + D)
    s2_0 = [crand(rng, chi2, chi2, D, D, D, D) for _ in range(6)]
    s2_1 = [crand(rng, chi2, chi2, D, D, D, D) for _ in range(6)]
    Z = np.diag([1.0, -1.0]).astype(complex)
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    G = crand(rng, 2, 2)
    terms = [(1, 0, None), (0, 1, None), (0, 0, -Z), (4, 0, X), (5, 0, -0.7 * X), (0, 4, -0.7 * X), (0, 5, G),
             (3, 2, None), (2, 3, None)]
    v = crand(rng, D, D, D, D, 2)
    ref = sum(dense.stage3_multiply(s2_0[a], s2_1[b], v, o) for a, b, o in terms)
    halves_0 = [dd.fromArray(x).join((0, 1), 4, 5, 2, 3) for x in s2_0]
    halves_1 = [dd.fromArray(x).join((1, 0), 4, 5, 2, 3) for x in s2_1]
    op = Stage3Operator(v.shape)
    for a, b, o in terms:
        op.add_term(halves_0[a], halves_1[b], o)
    op.finalize().set_path(path)
    assert op.num_terms == 9 and op.num_groups == 4          # B_I star (4 terms), A_I star (3), two single terms
    # 7 first + 6 second products per x instead of the reference's 9 + 9
    per_product = 8.0 * chi2 * chi2 * D ** 6 * 2
    assert abs(op.executed_flops - 13 * per_product) < 1e-6 * op.executed_flops
    out = op(dd.fromArray(v)).toArray()
    assert relerr(out, ref) < MATVEC_TOL


def test_stage3_golden(dd):
    """The reference's own multiplier output (tests/golden/dense_recipes.npz)."""
    from carcassonne_b200.operator import Stage3Operator, prejoin_halves
    g = load("dense_recipes")
    A, B = prejoin_halves(dd.fromArray(g["st3_s2_0"]), dd.fromArray(g["st3_s2_1"]))
    v = dd.fromArray(g["st3_v"])
    op = Stage3Operator(v.shape).add_term(A, B, None).finalize()
    assert relerr(op(v).toArray(), g["st3_norm_out"]) < MATVEC_TOL
    assert op.cost_of_multiply == int(g["st3_norm_cost"][0])
    op = Stage3Operator(v.shape).add_term(A, B, g["st3_op"]).finalize()
    assert relerr(op(v).toArray(), g["st3_dense_out"]) < MATVEC_TOL
    assert op.cost_of_multiply == int(g["st3_dense_cost"][0])


def test_stage3_host_entry_point(dd):
    """carc_stage3_matvec_host: the end-to-end C-ABI call with host buffers."""
    import ctypes as C
    from carcassonne_b200 import _lib
    from oracle import dense
    rng = np.random.default_rng(9)
    s2_0, s2_1, ops, v, ref = _stage3_case(rng, 4, 4, terms=2)
    halves = [dense.stage3_prejoin(a, b) for a, b in zip(s2_0, s2_1)]
    nt = len(halves)
    A = [np.ascontiguousarray(h[0]) for h in halves]
    B = [np.ascontiguousarray(h[1]) for h in halves]
    Ap = (C.c_void_p * nt)(*[a.ctypes.data for a in A])
    Bp = (C.c_void_p * nt)(*[b.ctypes.data for b in B])
    X = (C.c_int64 * nt)(*[a.shape[0] for a in A])
    opsc = [None if o is None else np.ascontiguousarray(o) for o in ops]
    Op = (C.POINTER(C.c_double) * nt)(*[
        C.cast(None, C.POINTER(C.c_double)) if o is None else o.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))
        for o in opsc])
    out = np.empty_like(v)
    _lib.check(_lib.lib.carc_stage3_matvec_host(nt, Ap, Bp, X, Op, 16, 16, 16, 16, 2, v.ctypes.data, out.ctypes.data,
                                                None))
    assert relerr(out, ref) < MATVEC_TOL


def test_dmma_peak_runs():
    import ctypes as C
    from carcassonne_b200 import _lib
    tf = C.c_double()
    _lib.check(_lib.lib.carc_dmma_peak(2000, C.byref(tf), None))
    print("DMMA peak TFLOP/s:", tf.value)
    assert tf.value > 1.0
