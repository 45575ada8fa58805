"""Device factorisations behind DeviceData.qr / svd / unitize / normalizeAxis.

Every matrix here is tall and skinny ((D^3 d) x D, (chi D) x chi, ...).  The pipeline is
Householder QR (one persistent CTA, LAPACK conventions) -> one-sided Jacobi SVD of the n x n factor ->
DMMA GEMMs for the products, all inside libcarc_b200.so.  Reference: data/__init__.py:263-301 (normalizeAxis),
344-346 (svd), 315-317 (qr), utils.py:879-881 (unitize).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check
from .data import DeviceData, _empty, _ptr, _stream, gemm

MAX_SMALL = 80   # carc_svd_small limit


def _qr_raw(M):
    """M: DeviceData [m, n], m >= n.  Returns torch buffers (Q [m,n], R [n,n])."""
    m, n = M.shape
    work = M.copy()._t
    R = _empty((n, n))
    Q = _empty((m, n))
    tau = _empty((n,))
    check(lib.carc_qr(_ptr(work), m, n, _ptr(R), _ptr(Q), _ptr(tau), _stream()))
    return Q, R


def qr(M, mode="full"):
    """scipy.linalg.qr(M, mode='economic') for m >= n (the only shape the path uses: newEnlargener)."""
    if M.ndim != 2:
        raise ValueError("qr needs a matrix")
    if mode != "economic":
        raise NotImplementedError("only mode='economic' is implemented on device")
    m, n = M.shape
    if m < n:
        raise NotImplementedError("economic QR of a wide matrix is not on the hot path")
    Q, R = _qr_raw(M)
    return DeviceData(Q), DeviceData(R)


def _svd_tall(M):
    """Thin SVD pieces of a tall matrix: Q [m,n], U_R [n,n], S [n] (complex (s,0)), Vh [n,n] as torch buffers."""
    m, n = M.shape
    if n > MAX_SMALL:
        raise NotImplementedError("device SVD supports at most {} columns (got {})".format(MAX_SMALL, n))
    Q, R = _qr_raw(M)
    U = _empty((n, n))
    S = _empty((n,))
    Vh = _empty((n, n))
    check(lib.carc_svd_small(_ptr(R), n, _ptr(U), _ptr(S), _ptr(Vh), _stream()))
    return Q, U, S, Vh


def svd(M, full_matrices=True):
    """scipy.linalg.svd(M, full_matrices=False): (U, S, Vh) with S a real-valued complex vector, descending."""
    if M.ndim != 2:
        raise ValueError("svd needs a matrix")
    m, n = M.shape
    if full_matrices and m != n:
        raise NotImplementedError("full_matrices=True is only implemented for square matrices")
    if m >= n:
        Q, U_R, S, Vh = _svd_tall(M)
        U = _empty((m, n))
        gemm(_lib.OP_N, _lib.OP_N, m, n, n, Q, n, U_R, n, U)
        return DeviceData(U), DeviceData(S), DeviceData(Vh)
    # wide: M^H = U' S V'^H  ->  M = V' S U'^H
    U2, S, Vh2 = svd(M.adjoint(), full_matrices=False)
    return Vh2.adjoint(), S, U2.adjoint()


def _normalizer_pieces(U_R, S, Vh, n, dont_recip_under):
    outs = [_empty((n, n)) for _ in range(5)]
    check(lib.carc_normalizer_matrices(_ptr(U_R), _ptr(S), _ptr(Vh), n, float(dont_recip_under or 0.0),
                                       *[_ptr(o) for o in outs], _stream()))
    return outs  # polar, normalizer, denormalizer, normalizer_sqrt, denormalizer_sqrt


def unitize(M):
    """utils.py:879-881: U Vh of the thin SVD (the polar isometry)."""
    m, n = M.shape
    if m < n:
        return unitize(M.adjoint()).adjoint()
    Q, U_R, S, Vh = _svd_tall(M)
    polar = _normalizer_pieces(U_R, S, Vh, n, 1e-14)[0]
    out = _empty((m, n))
    gemm(_lib.OP_N, _lib.OP_N, m, n, n, Q, n, polar, n, out)
    return DeviceData(out)


def normalize_axis(t, axis, sqrt_svals=False, dont_recip_under=1e-14):
    """NDArrayData.normalizeAxis (data/__init__.py:263-301)."""
    axis = int(axis)
    if t.shape[axis] == 1:
        n = t.norm()
        if sqrt_svals:
            n = np.sqrt(n)
            return DeviceData.fromArray(np.array([[1 / n]])), DeviceData.fromArray(np.array([[n]]))
        return t * (1.0 / n), DeviceData.fromArray(np.array([[1 / n]])), DeviceData.fromArray(np.array([[n]]))
    n = t.shape[axis]
    m = t.size() // n
    if m < n:
        raise ValueError("the total number of degrees of freedom in all other axes ({}) are not enough to normalize "
                         "axis ({}) with dimension ({})".format(m, axis, n))
    if n > MAX_SMALL:
        raise NotImplementedError("device SVD supports at most {} columns (got {})".format(MAX_SMALL, n))
    # one library call: permute to [(others), axis], Householder QR, Jacobi SVD of R, normaliser pieces and the isometric
    # tensor Q (U V^H) written with the normalised axis already back in place (csrc/recipes.cu: carc_normalize_axis)
    shape = (C.c_int64 * t.ndim)(*t.shape)
    nrm, den = _empty((n, n)), _empty((n, n))
    iso = None if sqrt_svals else _empty(t.shape)
    check(lib.carc_normalize_axis(_ptr(t._t), shape, t.ndim, axis, int(bool(sqrt_svals)), float(dont_recip_under or 0.0),
                                  _ptr(iso) if iso is not None else None, _ptr(nrm), _ptr(den), _stream()))
    if sqrt_svals:
        return DeviceData(nrm), DeviceData(den)
    return DeviceData(iso), DeviceData(nrm), DeviceData(den)
