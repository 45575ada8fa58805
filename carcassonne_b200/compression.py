"""State-bond compression: ``computeProductCompressor`` (reference carcassonne/compression.py:26-45) on device.

Find the isometry c [new, old] such that projecting the shared (state, state*) bond pair between L and R with
c (x) c* preserves the product L.R: alternating least squares from a random start, each round a least-squares
problem min_x |A(c) x - b| followed by a polar projection (``unitize``).

Device formulation: the (l r) x (old new) matrix A(c) of the reference's generated ``formMatrix`` is produced by
three small mode products and one GEMM, the normal equations A^H A x = A^H b are formed in Gram form by two more
GEMMs (an (old new)^2 matrix -- a few hundred on a side) and solved directly with the device LU after a relative
shift of 1e-10 on the diagonal.  The reference hands them to GMRES(20) at rtol 1e-5, which needs thousands of
matvecs on these ill-conditioned systems (kappa ~ 1e5 at chi = 8, D = 4) and stops 5 % away from the least-squares
solution; the shifted direct solve returns the minimum-norm least-squares solution to ~1e-10, deterministically, in
a few milliseconds.  (Device GMRES and CG exist behind carc_gmres / carc_cg and are parity-tested.)  The polar factor
comes from the device Jacobi SVD.
"""
import ctypes as C

from . import _lib
from ._lib import lib, check
from .data import DeviceData, _empty, _ptr, _stream, gemm, gemm_hermitian, gemm_scatter
from .utils import LUFactors, SolverDidNotConverge, _DenseOperator


# True: run the Gram-form ALS round by round from Python (the pre-round-2 path, kept as the cross-check of
# carc_product_compressor in tests/test_gpu_system.py)
_PYTHON_ALS = False


def _gmres_dense(matrix, rhs, rtol=1e-5, restart=20, maxiter=None):
    """x = matrix^-1 rhs by GMRES from x0 = 0 (scipy.sparse.linalg.gmres defaults)."""
    n = matrix.shape[0]
    op = _DenseOperator(matrix)
    x = _empty((n,))
    iters = C.c_int(0)
    resid = C.c_double(0.0)
    rc = lib.carc_gmres(op._handle, _ptr(rhs._t), _ptr(x), float(rtol), int(restart),
                        int(maxiter if maxiter is not None else max(200, n)), C.byref(iters), C.byref(resid), _stream())
    op.close()
    if rc == _lib.ERR_NO_CONVERGENCE:
        raise SolverDidNotConverge(lib.carc_last_error().decode("utf-8", "replace"))
    check(rc)
    return DeviceData(x)


def _cg_dense(matrix, rhs, rtol=1e-10, maxiter=None):
    """Minimum-norm solution of the Hermitian PSD system matrix x = rhs by device CG from x0 = 0."""
    n = matrix.shape[0]
    op = _DenseOperator(matrix)
    x = _empty((n,))
    iters = C.c_int(0)
    resid = C.c_double(0.0)
    rc = lib.carc_cg(op._handle, _ptr(rhs._t), _ptr(x), float(rtol), int(maxiter if maxiter is not None else 20 * n),
                     C.byref(iters), C.byref(resid), _stream())
    op.close()
    check(rc)
    return DeviceData(x), iters.value, resid.value


def formProductCompressorMatrix(L, c, R):
    """A[(l r), (i n)] = sum L[l,i,j,p] conj(c)[j,m] conj(c)[k,n] c[q,m] R[k,q,p,r] with c of shape [old, new]
    (the generated contractor of reference compression.py:11-25, rows [L0 R3], columns [L1 c^H_0])."""
    l, old, _, p = L.shape
    r = R.shape[3]
    new = c.shape[1]
    # Pc[j, q] = sum_m conj(c)[j, m] c[q, m]
    Pc = _empty((old, old))
    gemm(_lib.OP_J, _lib.OP_T, old, old, new, c._t, new, c._t, new, Pc)
    # Lc[l, i, q, p] = sum_j L[l, i, j, p] Pc[j, q]
    Lc = L.absorbMatrixAt(2, DeviceData(Pc).transpose())
    # Rc[q, p, r, n] = sum_k conj(c)[k, n] R[k, q, p, r]      (a small tensor: new x old x p x r)
    Rc = _empty((old * p * r, new))
    gemm(_lib.OP_T, _lib.OP_J, old * p * r, new, old, R._t, old * p * r, c._t, new, Rc)
    # A[l, r, i, n] = sum_{q, p} Lc[(l i), (q p)] Rc[(q p), (r n)], written straight into the [(l r), (i n)] layout
    A = _empty((l * r, old * new))
    gemm_scatter(_lib.OP_N, _lib.OP_N, l * old, r * new, old * p, Lc._t, old * p, Rc, r * new, A,
                 ((l, r * old * new), (old, new)), ((r, old * new), (new, 1)))
    return DeviceData(A)


def _solve_normal_equations(gram, rhs, regularization):
    """x = (G + eps mean(diag G) I)^-1 rhs by the device LU.  G = A^H A is singular whenever the product is exactly
    compressible; the shift keeps the factorisation well defined, leaves the solution in range(G) untouched to
    O(eps) and returns (numerically) zero along the null space -- the minimum-norm least-squares solution the
    reference's GMRES from x0 = 0 approximates to 1e-5."""
    n = gram.shape[0]
    eye = DeviceData.newIdentity(n)
    trace = eye.contractWithAlongAll(gram)                   # sum_i G[i, i] (reduction kernel)
    gram._axpby(regularization * float(trace.real) / n, eye, 1.0)
    return LUFactors(gram).solve(rhs)


def factored_side_gram(side, end):
    """Gram matrix of an enlarged side over everything but the state-bond pair at one end, built from the factors
    of the center -> side absorption that produced it (``side' = side (x) E`` summed over the (g h) legs, recorded by
    tensors/_2d/dense._absorb_center), or None when the tensor carries no such record.

        end = 0:  G[(k q),(k' q')] = sum_rest conj(side'[k, q, rest]) side'[k', q', rest]    (axes 0, 1: the R of
                  compressCornerStateTowardsRight, reference system/_2d.py:214-228)
        end = 1:  the same over axes 3, 4 (the L of compressCornerStateTowardsLeft, system/_2d.py:199-213)

    With S2 = sum conj(side) side over the far end and F = sum conj(E) E over the far and outward center legs,
    G = sum_{g h g' h'} S2[.. g h .. g' h'] F[g h .. g' h' ..]: chi^4 D^8 multiply-adds in one GEMM instead of the
    chi^6 D^10 of the direct sum over the enlarged tensor (32 x fewer at chi = D = 8), and the enlarged tensor is not
    read at all."""
    record = getattr(side, "_factors", None)
    if record is None or record[0] != "center_into_side":
        return None
    _, small, E, dims = record
    s = small.shape
    g, h, nl, ml, nr, mr, no, mo = dims
    E8 = E.split(g, h, nl, ml, nr, mr, no, mo)
    if end == 0:
        Sm = small.join((0, 1, 6, 7), (2, 3, 4, 5))
        Em = E8.join((0, 1, 2, 3), (4, 5, 6, 7))
        a, b, va, vb = s[0], s[1], nl, ml
    else:
        Sm = small.join((3, 4, 6, 7), (0, 1, 2, 5))
        Em = E8.join((0, 1, 4, 5), (2, 3, 6, 7))
        a, b, va, vb = s[3], s[4], nr, mr
    ns, ks = Sm.shape
    S2 = _empty((ns, ns))
    gemm_hermitian(_lib.OP_J, _lib.OP_T, ns, ks, Sm._t, ks, Sm._t, ks, S2)     # [(a b g h), (a' b' g' h')]
    ne, ke = Em.shape
    F = _empty((ne, ne))
    gemm_hermitian(_lib.OP_J, _lib.OP_T, ne, ke, Em._t, ke, Em._t, ke, F)      # [(g h va vb), (g' h' va' vb')]
    S2g = DeviceData(S2).split(a, b, g, h, a, b, g, h).join((0, 1, 4, 5), (2, 3, 6, 7))
    Fg = DeviceData(F).split(g, h, va, vb, g, h, va, vb).join((0, 1, 4, 5), (2, 3, 6, 7))
    G = S2g.contractWith(Fg, (1,), (0,)).split(a, b, a, b, va, vb, va, vb)      # [a, b, a', b', va, vb, va', vb']
    return G.join((0, 4, 1, 5), (2, 6, 3, 7))                                   # [((a va) (b vb)), ((a' va') (b' vb'))]


class _GramForm:
    """Normal equations of the ALS without ever forming A (A is (l r) x (old new), 137 GB
    at chi = D = 8).  With LL0 = L^H L and RR0 = R R^H over the outer legs (computed ONCE per compression: two
    Hermitian DMMA GEMMs with K = l and K = r) and T = LL0 . RR0^T,

        A^H A [(i n),(i' n')] = sum_{j j'} LL0[i,j,i',j'] W[j,j',n,n'],   W = sum_{q q'} conj(Pc[j,q]) Pc[j',q'] RRc[n,q,n',q']
        A^H b [(i n)]         = sum_{j q k} conj(Pc[j,q]) c[k,n] T[i,j,k,q]
        RRc = c^T RR0 conj(c),   Pc = conj(c) c^T

    so an ALS round touches only old^4-sized objects through a handful of small GEMMs (the big factor LL0 enters one
    (old^2 x old^2) x (old^2 x new^2) product) instead of O(l r old^2 new^2) work.  Operator bond 1 (Identity
    tensors) only."""

    def __init__(self, L, R, left_gram=None, right_gram=None):
        l, old, _, _ = L.shape
        r = R.shape[3]
        self.old = old
        o2 = old * old
        if left_gram is not None:
            LL0 = left_gram._t
        else:
            LL0 = _empty((o2, o2))
            gemm_hermitian(_lib.OP_C, _lib.OP_N, o2, l, L._t, o2, L._t, o2, LL0)   # sum_l conj(L[l,(ij)]) L[l,(i'j')]
        if right_gram is not None:
            RR0 = right_gram._t
        else:
            RR0 = _empty((o2, o2))
            gemm_hermitian(_lib.OP_J, _lib.OP_T, o2, r, R._t, r, R._t, r, RR0)     # sum_r conj(R[(kq),r]) R[(k'q'),r]
        T = _empty((o2, o2))
        gemm(_lib.OP_N, _lib.OP_T, o2, o2, o2, LL0, o2, RR0, o2, T)                 # T[(ij),(kq)]
        # LL0 regrouped once to [(i i'), (j j')] so that every round is a single GEMM against W
        self.LLg = DeviceData(LL0).split(old, old, old, old).join((0, 2), (1, 3))
        self.RR0 = DeviceData(RR0).split(old, old, old, old)
        self.T = DeviceData(T).split(old, old, old, old)

    def normal_equations(self, c):
        old, new = c.shape
        Pc = _empty((old, old))
        gemm(_lib.OP_J, _lib.OP_T, old, old, new, c._t, new, c._t, new, Pc)
        Pc = DeviceData(Pc)
        RRc = self.RR0.absorbMatrixAt(0, c.transpose()).absorbMatrixAt(2, c.adjoint())         # [n, q, n', q']
        W = RRc.absorbMatrixAt(1, Pc.conj()).absorbMatrixAt(3, Pc)                             # [n, j, n', j']
        Wg = W.join((1, 3), (0, 2))                                                            # [(j j'), (n n')]
        G4 = self.LLg.contractWith(Wg, (1,), (0,)).split(old, old, new, new)                   # [i, i', n, n']
        gram = G4.join((0, 2), (1, 3))                                                         # [(i n), (i' n')]
        U = self.T.absorbMatrixAt(2, c.transpose())                                            # [i, j, n, q]
        rhs = U.contractWith(Pc.conj(), (1, 3), (0, 1)).ravel()                                # [(i n)]
        return gram, rhs


def computeProductCompressor(L, R, new_dimension, initial=None, sweeps=4, regularization=1e-10, left_gram=None,
                             right_gram=None):
    """reference compression.py:26-45.  ``initial`` (optional) is the random [old, new] draw; by default it is drawn
    from the host NumPy stream with ``newRandom`` exactly where the reference draws it.  ``left_gram`` / ``right_gram``
    (optional, operator bond 1 only) are L^H L / R R^H over the outer legs when the caller already has them
    (``factored_side_gram``)."""
    if L.shape[1] != L.shape[2]:
        raise ValueError("left inward dimensions do not match (given " + str(L.shape) + ")")
    if R.shape[0] != R.shape[1]:
        raise ValueError("right inward dimensions do not match (given " + str(R.shape) + ")")
    if L.shape[1] != R.shape[1]:
        raise ValueError("left and right shapes are incompatible (given " + str(L.shape) + " and " + str(R.shape) + ")")
    old_dimension = L.shape[1]
    if initial is None:
        initial = DeviceData.newRandom(old_dimension, new_dimension)
    if new_dimension == old_dimension:
        # every unitary preserves the product exactly: the starting point already is a global minimiser of the
        # ALS objective (the reference still runs its four rounds and lands on some other unitary)
        return initial.unitize().transpose()
    m = old_dimension * new_dimension
    if L.shape[3] == 1 and new_dimension <= 80 and not _PYTHON_ALS:
        # the whole fit in one library call (csrc/recipes.cu: carc_product_compressor): Gram form, no host round trip
        out = _empty((new_dimension, old_dimension))
        check(lib.carc_product_compressor(
            _ptr(L._t), L.shape[0], _ptr(R._t), R.shape[3], old_dimension, new_dimension, _ptr(initial._t), int(sweeps),
            float(regularization), _ptr(left_gram._t) if left_gram is not None else None,
            _ptr(right_gram._t) if right_gram is not None else None, _ptr(out), _stream()))
        return DeviceData(out)
    compressor = initial.unitize()
    if L.shape[3] == 1:
        form = _GramForm(L, R, left_gram, right_gram)
    else:
        form = None
        b = L.contractWith(R, (1, 2, 3), (0, 1, 2)).ravel()
        rows = b.shape[0]
    for _ in range(sweeps):
        if form is not None:
            gram, rhs = form.normal_equations(compressor)
        else:
            A = formProductCompressorMatrix(L, compressor, R)
            gram = _empty((m, m))
            gemm(_lib.OP_C, _lib.OP_N, m, m, rows, A._t, m, A._t, m, gram)          # A^H A
            rhs = _empty((m,))
            gemm(_lib.OP_C, _lib.OP_N, m, 1, rows, A._t, m, b._t, 1, rhs)           # A^H b
            del A
            gram, rhs = DeviceData(gram), DeviceData(rhs)
        x = _solve_normal_equations(gram, rhs, regularization)
        compressor = x.split(old_dimension, new_dimension).unitize()
    return compressor.transpose()


__all__ = ["computeProductCompressor"]
