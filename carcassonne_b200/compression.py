"""State-bond compression: ``computeProductCompressor`` (reference carcassonne/compression.py:26-45) on device.

Find the isometry c [new, old] such that projecting the shared (state, state*) bond pair between L and R with
c (x) c* preserves the product L.R: alternating least squares from a random start, each round a least-squares
problem min_x |A(c) x - b| followed by a polar projection (``unitize``).

Device formulation: the (l r) x (old new) matrix A(c) of the reference's generated ``formMatrix`` is produced by
three small mode products and one GEMM, the normal equations A^H A x = A^H b are formed in Gram form by two more
GEMMs (an (old new)^2 matrix -- a few hundred on a side), solved with the device GMRES exactly as the reference
solves them (restart 20, rtol 1e-5: a Krylov solver, because A^H A is singular whenever the product is exactly
compressible), and the polar factor comes from the device Jacobi SVD.
"""
import ctypes as C

from . import _lib
from ._lib import lib, check
from .data import DeviceData, _empty, _ptr, _stream, gemm
from .utils import SolverDidNotConverge, _DenseOperator


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
    # Rc[n, q, p, r] = sum_k conj(c)[k, n] R[k, q, p, r]
    Rc = R.absorbMatrixAt(0, c.adjoint())
    # A[l, i, n, r] = sum_{q, p} Lc[l, i, q, p] Rc[n, q, p, r]
    A4 = Lc.contractWith(Rc, (2, 3), (1, 2))
    return A4.join((0, 3), (1, 2))


def computeProductCompressor(L, R, new_dimension, initial=None, sweeps=4):
    """reference compression.py:26-45.  ``initial`` (optional) is the random [old, new] draw; by default it is drawn
    from the host NumPy stream with ``newRandom`` exactly where the reference draws it."""
    if L.shape[1] != L.shape[2]:
        raise ValueError("left inward dimensions do not match (given " + str(L.shape) + ")")
    if R.shape[0] != R.shape[1]:
        raise ValueError("right inward dimensions do not match (given " + str(R.shape) + ")")
    if L.shape[1] != R.shape[1]:
        raise ValueError("left and right shapes are incompatible (given " + str(L.shape) + " and " + str(R.shape) + ")")
    old_dimension = L.shape[1]
    b = L.contractWith(R, (1, 2, 3), (0, 1, 2)).ravel()
    if initial is None:
        initial = DeviceData.newRandom(old_dimension, new_dimension)
    compressor = initial.unitize()
    m = old_dimension * new_dimension
    rows = b.shape[0]
    for _ in range(sweeps):
        A = formProductCompressorMatrix(L, compressor, R)
        gram = _empty((m, m))
        gemm(_lib.OP_C, _lib.OP_N, m, m, rows, A._t, m, A._t, m, gram)          # A^H A
        rhs = _empty((m,))
        gemm(_lib.OP_C, _lib.OP_N, m, 1, rows, A._t, m, b._t, 1, rhs)           # A^H b
        x = _gmres_dense(DeviceData(gram), DeviceData(rhs))
        compressor = x.split(old_dimension, new_dimension).unitize()
    return compressor.transpose()


__all__ = ["computeProductCompressor"]
