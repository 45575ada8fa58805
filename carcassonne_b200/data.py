"""Device-resident tensor class with the duck-typed interface of the reference's ``NDArrayData``
(reference carcassonne/data/__init__.py:30-365).

A ``DeviceData`` owns a C-contiguous complex128 buffer in B200 HBM (allocated through torch's caching
allocator -- torch is plumbing: memory, streams, NCCL) and performs every operation by calling
libcarc_b200.so through ctypes with raw device pointers.  Nothing is computed on the host and there is no
CPU fallback; host arrays appear only in ``fromArray`` / ``toArray`` / scalar read-backs and for random
draws, which stay on the NumPy RNG so that seeded runs consume the same stream as the reference
(SURVEY.md section 8b).

Tensors behave as immutable values (``System.__copy__`` is shallow and shares them, reference
system/_2d.py:122-131); only ``+=`` / ``*=`` mutate, and the library applies those to freshly created results
only (reference sparse.py:236).
"""
import ctypes as C
from math import prod as _prod

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

_c128 = torch.complex128


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _empty(shape):
    return torch.empty(tuple(int(s) for s in shape), dtype=_c128, device="cuda")


def _ptr(t, offset_elems=0):
    return C.c_void_p(t.data_ptr() + 16 * int(offset_elems))


def _scalar_buffer():
    return torch.empty(2, dtype=torch.float64, device="cuda")


_MAX_BATCH = 65535


def gemm(opA, opB, M, N, K, A, lda, B, ldb, Cbuf, alpha=1.0, beta=0.0, out_map=None, k_map=None,
         batch=1, strideA=0, strideB=0, strideC=0):
    """Thin wrapper over carc_zgemm on raw torch buffers (batches beyond the grid limit are chunked)."""
    om = (C.c_int64 * 6)(*out_map) if out_map is not None else None
    km = (C.c_int64 * 4)(*k_map) if k_map is not None else None
    done = 0
    while done < batch:
        nb = min(_MAX_BATCH, batch - done)
        check(lib.carc_zgemm(opA, opB, M, N, K, _lib.cplx2(alpha), _ptr(A, done * strideA), lda,
                             _ptr(B, done * strideB), ldb, _lib.cplx2(beta), _ptr(Cbuf, done * strideC), om, km, nb,
                             strideA, strideB, strideC, _stream()))
        done += nb


def gemm_hermitian(opA, opB, N, K, A, lda, B, ldb, Cbuf):
    """C[N, N] = op(A) op(B) for a Hermitian product (Gram matrices): upper tiles computed, lower mirrored."""
    check(lib.carc_zgemm_hermitian(opA, opB, N, K, _ptr(A), lda, _ptr(B), ldb, _ptr(Cbuf), _stream()))


_TABLES = {}


def index_table(levels):
    """Device int64 table t[i] = sum_l digit_l(i) * stride_l for levels = ((extent, stride), ...), digits taken
    row-major (last level fastest).  Cached: the same few layouts recur on every sweep iteration."""
    levels = tuple((int(e), int(s)) for e, s in levels if int(e) != 1) or ((1, 0),)
    key = (torch.cuda.current_device(), levels)
    t = _TABLES.get(key)
    if t is None:
        if len(_TABLES) > 512:
            _TABLES.clear()
        total = _prod(e for e, _ in levels)
        t = torch.empty(max(total, 1), dtype=torch.int64, device="cuda")
        nl = len(levels)
        check(lib.carc_index_table(nl, (C.c_int64 * nl)(*[e for e, _ in levels]),
                                   (C.c_int64 * nl)(*[s for _, s in levels]), C.c_void_p(t.data_ptr()), _stream()))
        _TABLES[key] = t
    return t


def gemm_scatter(opA, opB, M, N, K, A, lda, B, ldb, Cbuf, row_levels, col_levels, alpha=1.0, beta=0.0,
                 batch=1, strideA=0, strideB=0, strideC=0, a_offset=0, b_offset=0, c_offset=0):
    """C[rowoff(m) + coloff(n)] = alpha * op(A) op(B) + beta * C: a contraction written straight into the layout
    the reference's final ``join`` would produce (carc_zgemm_tab)."""
    if M * N * batch == 0:
        return
    rt = index_table(row_levels)
    ct = index_table(col_levels)
    done = 0
    while done < batch:
        nb = min(_MAX_BATCH, batch - done)
        check(lib.carc_zgemm_tab(opA, opB, M, N, K, _lib.cplx2(alpha), _ptr(A, a_offset + done * strideA), lda,
                                 _ptr(B, b_offset + done * strideB), ldb, _lib.cplx2(beta),
                                 _ptr(Cbuf, c_offset + done * strideC), C.c_void_p(rt.data_ptr()),
                                 C.c_void_p(ct.data_ptr()), nb, strideA, strideB, strideC, _stream()))
        done += nb


class DeviceData:
    """complex128 tensor in device memory; API of the reference's NDArrayData."""

    # _factors: optional record of how this tensor was produced (set by the center -> side absorption, read by the
    # state-bond compression to build Gram matrices from the factors); dropped by every in-place update
    __slots__ = ["_t", "_factors"]

    # -- construction -----------------------------------------------------------------------------------
    def __init__(self, t):
        if isinstance(t, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.complex128)).to("cuda")
        if t.dtype != _c128 or not t.is_cuda:
            raise TypeError("DeviceData wraps complex128 CUDA buffers")
        self._t = t if t.is_contiguous() else t.contiguous()
        self._factors = None

    @classmethod
    def fromArray(cls, arr):
        return cls(np.asarray(arr, dtype=np.complex128))

    @classmethod
    def newCollected(cls, datas):  # data/__init__.py:35-37
        datas = list(datas)
        out = _empty((len(datas),) + tuple(datas[0].shape))
        for i, d in enumerate(datas):
            check(lib.carc_axpby(d.size(), _lib.cplx2(1), _ptr(d._t), _lib.cplx2(0), C.c_void_p(out[i].data_ptr()), 0,
                                 _stream()))
        return cls(out)

    @classmethod
    def newDiagonal(cls, data):  # data/__init__.py:38-41
        return cls.fromArray(np.diag(np.asarray(data)))

    @classmethod
    def newEnlargener(cls, old_dimension, new_dimension, dtype=None):  # data/__init__.py:42-50
        if new_dimension == old_dimension:
            return (cls.newIdentity(new_dimension),) * 2
        matrix = cls.newRandom(new_dimension, old_dimension).qr(mode="economic")[0]
        return matrix, matrix.conj()

    @classmethod
    def newFilled(cls, shape, value, dtype=None):  # data/__init__.py:51-58
        return cls.fromArray(np.full(shape, value, dtype=np.complex128))

    @classmethod
    def newIdentity(cls, N, dtype=None):  # data/__init__.py:59-62
        return cls.fromArray(np.identity(N, dtype=np.complex128))

    @classmethod
    def newOuterProduct(cls, *factors):  # data/__init__.py:63-66
        from functools import reduce
        return cls.fromArray(reduce(np.multiply.outer, [np.asarray(f, dtype=np.complex128) for f in factors]))

    @classmethod
    def newRandom(cls, *shape):
        """data/__init__.py:67-70 -> utils.randomComplexSample (utils.py:795-797): host NumPy RNG, uploaded."""
        sample = np.random.random_sample(shape) * 2 - 1 + np.random.random_sample(shape) * 2j - 1j
        return cls.fromArray(sample)

    @classmethod
    def newRandomHermitian(cls, *shape):  # data/__init__.py:71-76
        data = cls.newRandom(*shape)
        data += data.transpose().conj()
        return data

    @classmethod
    def newNormalizedRandom(cls, *shape):  # data/__init__.py:77-82
        sample = np.random.random_sample(shape) * 2 - 1 + np.random.random_sample(shape) * 2j - 1j
        sample /= np.linalg.norm(sample)
        return cls.fromArray(sample)

    @classmethod
    def newTrivial(cls, shape, dtype=None):  # data/__init__.py:83-86
        return cls.fromArray(np.ones(shape, dtype=np.complex128))

    @classmethod
    def newZeros(cls, shape, dtype=None):  # data/__init__.py:87-90
        return cls(torch.zeros(tuple(shape), dtype=_c128, device="cuda"))

    # -- properties -------------------------------------------------------------------------------------
    shape = property(lambda self: tuple(self._t.shape))
    ndim = property(lambda self: self._t.dim())
    dtype = property(lambda self: np.dtype(np.complex128))

    def size(self):
        return int(self._t.numel())

    def toArray(self):
        return self._t.cpu().numpy()

    def toNDArrayData(self):
        raise TypeError("DeviceData has no host twin; use toArray()")

    def __repr__(self):
        return "DeviceData(shape={})".format(self.shape)

    # -- elementwise ------------------------------------------------------------------------------------
    def _touch(self):
        """Called by everything that writes this buffer through a raw pointer: drops the derived record and bumps
        torch's version counter, which every view of the same storage shares -- the environment cache keys on it
        (tensors/_2d/sparse.py), so an in-place update can never be answered with stale multipliers."""
        self._factors = None
        torch.autograd.graph.increment_version(self._t)

    @property
    def version(self):
        return self._t._version

    def _axpby(self, alpha, x, beta, conj_x=0):
        """self = alpha * x + beta * self (in place)."""
        self._touch()
        check(lib.carc_axpby(self.size(), _lib.cplx2(alpha), _ptr(x._t), _lib.cplx2(beta), _ptr(self._t), conj_x,
                             _stream()))
        return self

    def _scaled(self, alpha, conj=0):
        out = DeviceData(_empty(self.shape))
        return out._axpby(alpha, self, 0.0, conj)

    def _check_same_shape(self, other):
        if self.shape != other.shape:
            raise ValueError("shape mismatch: {} vs {}".format(self.shape, other.shape))

    def __add__(self, other):
        self._check_same_shape(other)
        return self._scaled(1.0)._axpby(1.0, other, 1.0)

    def __sub__(self, other):
        self._check_same_shape(other)
        return self._scaled(1.0)._axpby(-1.0, other, 1.0)

    def __iadd__(self, other):
        self._check_same_shape(other)
        return self._axpby(1.0, other, 1.0)

    def __neg__(self):
        return self._scaled(-1.0)

    def __copy__(self):
        return self._scaled(1.0)

    def copy(self):
        return self.__copy__()

    def __mul__(self, other):
        if isinstance(other, DeviceData):
            if other.shape == self.shape:
                out = self._scaled(1.0)
                check(lib.carc_mul(out.size(), _ptr(other._t), _ptr(out._t), _stream()))
                return out
            return self._broadcast_mul(other)
        return self._scaled(other)

    __rmul__ = __mul__

    def __imul__(self, other):
        self._check_same_shape(other)
        self._touch()
        check(lib.carc_mul(self.size(), _ptr(other._t), _ptr(self._t), _stream()))
        return self

    def _broadcast_mul(self, other):
        """Row scaling M * s with s of shape (n, 1) (``V*S`` in normalizeAxis, data/__init__.py:292-301):
        out[i, :] = s[i] * M[i, :] as n batched 1 x m x 1 GEMMs."""
        if self.ndim == 2 and other.shape == (self.shape[0], 1):
            n, m = self.shape
            out = DeviceData(_empty((n, m)))
            gemm(_lib.OP_N, _lib.OP_N, 1, m, 1, other._t, 1, self._t, m, out._t, batch=n, strideA=1, strideB=m,
                 strideC=m)
            return out
        raise ValueError("unsupported broadcast {} * {}".format(self.shape, other.shape))

    def __truediv__(self, other):
        if isinstance(other, DeviceData):
            raise NotImplementedError("elementwise division of device tensors is not on the hot path")
        return self._scaled(1.0 / other)

    def conj(self):
        return self._scaled(1.0, conj=1)

    # -- reductions (scalar read-backs synchronise) -------------------------------------------------------
    def norm(self):
        buf = _scalar_buffer()
        check(lib.carc_sumsq(self.size(), _ptr(self._t), _ptr(buf), _stream()))
        return float(np.sqrt(buf.cpu().numpy()[0]))

    def hasNaN(self):
        buf = _scalar_buffer()
        check(lib.carc_count_nonfinite(self.size(), _ptr(self._t), _ptr(buf), _stream()))
        return bool(buf.cpu().numpy()[1] > 0)

    def contractWithAlongAll(self, other):
        """sum_i self_i * other_i (no conjugation; data/__init__.py:160-163); returns a host complex scalar."""
        self._check_same_shape(other)
        buf = _scalar_buffer()
        check(lib.carc_dotu(self.size(), _ptr(self._t), _ptr(other._t), _ptr(buf), _stream()))
        r = buf.cpu().numpy()
        return np.complex128(complex(r[0], r[1]))

    def extractScalar(self):
        if self.ndim != 0:
            raise ValueError("tensor is not a scalar")
        return self._t.cpu().numpy()

    def allcloseTo(self, other, rtol=1e-05, atol=1e-08):
        return bool(np.allclose(self.toArray(), other.toArray(), rtol=rtol, atol=atol))

    def isCloseTo(self, other, rtol=1e-7, atol=1e-7):
        ndiff = (self - other).norm()
        return ndiff <= atol or ndiff / (self.norm() + other.norm()) / 2 <= rtol

    def normalized(self):
        return self._scaled(1.0 / self.norm())

    # -- views / permutes -------------------------------------------------------------------------------
    def split(self, *splits):
        return DeviceData(self._t.reshape(tuple(int(s) for s in splits)))

    def splitAt(self, index, *split):
        shape = list(self.shape)
        assert _prod(split) == shape[index]
        return self.split(*(shape[:index] + list(split) + shape[index + 1:]))

    def ravel(self):
        return DeviceData(self._t.reshape(-1))

    def dropUnitAxis(self, axis):
        if self.shape[axis] != 1:
            raise ValueError("Axis {} has non-unit dimension {}.".format(axis, self.shape[axis]))
        shape = list(self.shape)
        del shape[axis]
        return self.split(*shape)

    def _permuted(self, perm, conj=0):
        perm = [int(p) for p in perm]
        shape = self.shape
        new_shape = tuple(shape[p] for p in perm)
        if not conj and perm == sorted(perm):
            return self
        out = _empty(new_shape)
        nd = len(perm)
        check(lib.carc_permute(_ptr(self._t), _ptr(out), nd, (C.c_int64 * max(nd, 1))(*shape),
                               (C.c_int32 * max(nd, 1))(*perm), conj, 0, _stream()))
        return DeviceData(out)

    def transpose(self, *args):
        if len(args) == 0:
            perm = list(range(self.ndim))[::-1]
        elif len(args) == 1 and hasattr(args[0], "__len__"):
            perm = list(args[0])
        else:
            perm = list(args)
        return self._permuted(perm)

    def join(self, *groups):
        """data/__init__.py:247-256: transpose so that each group's axes are adjacent, then merge each group."""
        groups = [[g] if isinstance(g, int) else list(g) for g in groups]
        perm = [i for g in groups for i in g]
        if sorted(perm) != list(range(self.ndim)):
            raise ValueError("join groups {} are not a permutation of the {} axes".format(groups, self.ndim))
        shape = self.shape
        new_shape = [_prod(shape[i] for i in g) for g in groups]
        return DeviceData(self._permuted(perm)._t.reshape(new_shape))

    def fold(self, axis):  # data/__init__.py:204-208
        others = list(range(self.ndim))
        del others[axis]
        return self.join(axis, others)

    def adjoint(self):
        if self.ndim != 2:
            raise ValueError("Adjoint may only be computed for rank 2 tensors.")
        return self._permuted([1, 0], conj=1)

    def reverseLastAxis(self):  # data/__init__.py:321-323
        return DeviceData(torch.flip(self._t, dims=(-1,)).contiguous())

    def splitAtByRoot(self, index, root):  # data/__init__.py:336-340
        d = round(self.shape[index] ** (1.0 / root))
        return self.splitAt(index, *(d,) * root)

    def __getitem__(self, index):
        """Sub-block copy (used for the column slices of the operator compressors, system/_2d.py:312-356)."""
        return DeviceData(self._t[index].contiguous())

    def __setitem__(self, index, value):
        self._factors = None
        self._t[index] = value._t if isinstance(value, DeviceData) else value

    def __str__(self):
        return "DeviceData({})".format(self.toArray())

    # -- contraction ------------------------------------------------------------------------------------
    def contractWith(self, other, self_axes, other_axes):
        """numpy.tensordot(self, other, (self_axes, other_axes)) on the FP64 tensor pipe
        (data/__init__.py:157-159).  Operands already laid out as [free, contracted] or [contracted, free] are
        used in place (transposed-operand GEMM); anything else is permuted once."""
        a_axes = [int(a) % max(self.ndim, 1) for a in self_axes]
        b_axes = [int(b) % max(other.ndim, 1) for b in other_axes]
        if len(a_axes) != len(b_axes):
            raise ValueError("axis lists differ in length")
        for ia, ib in zip(a_axes, b_axes):
            if self.shape[ia] != other.shape[ib]:
                raise ValueError("shape mismatch for contraction: {}[{}] vs {}[{}]".format(self.shape, ia,
                                                                                        other.shape, ib))
        pairs = sorted(zip(a_axes, b_axes))  # contracted axes in self's storage order
        a_axes = [p[0] for p in pairs]
        b_axes = [p[1] for p in pairs]
        a_free = [i for i in range(self.ndim) if i not in a_axes]
        b_free = [i for i in range(other.ndim) if i not in b_axes]
        M = _prod(self.shape[i] for i in a_free)
        N = _prod(other.shape[i] for i in b_free)
        K = _prod(self.shape[i] for i in a_axes)
        out_shape = tuple(self.shape[i] for i in a_free) + tuple(other.shape[i] for i in b_free)

        def layout(t, free, contracted):
            order = list(range(t.ndim))
            if order == free + contracted:
                return t, True          # [free, K]: K contiguous
            if order == contracted + free:
                return t, False         # [K, free]
            return t._permuted(free + contracted), True

        A, a_kc = layout(self, a_free, a_axes)
        B, b_kc = layout(other, b_free, b_axes)
        out = _empty(out_shape)
        if M * N > 0:
            if K == 0:
                out.zero_()
            else:
                gemm(_lib.OP_N if a_kc else _lib.OP_T, _lib.OP_T if b_kc else _lib.OP_N, M, N, K,
                     A._t, K if a_kc else M, B._t, K if b_kc else N, out)
        return DeviceData(out)

    def absorbMatrixAt(self, axis, matrix):
        """out[..., j, ...] = sum_k matrix[j, k] self[..., k, ...]   (data/__init__.py:148-150), written by one
        batched GEMM straight into the final layout (no axis-rotation copy)."""
        if matrix.ndim != 2 or matrix.shape[1] != self.shape[axis]:
            raise ValueError("matrix of shape {} cannot be absorbed at axis {} of {}".format(matrix.shape, axis,
                                                                                          self.shape))
        shape = self.shape
        pre = _prod(shape[:axis])
        post = _prod(shape[axis + 1:])
        k, j = shape[axis], matrix.shape[0]
        out = _empty(shape[:axis] + (j,) + shape[axis + 1:])
        if pre * post * j > 0 and j <= 16 and k <= 640 and post >= 32:
            # a short matrix against a long tensor (compressor projections): streaming kernel, bound by reading self
            check(lib.carc_mode_product(_ptr(matrix._t), _ptr(self._t), _ptr(out), j, k, pre, post, _stream()))
        elif pre * post * j > 0:
            # for each leading index: out[pre][j, post] = matrix[j, k] . self[pre][k, post]
            gemm(_lib.OP_N, _lib.OP_N, j, post, k, matrix._t, k, self._t, post, out, batch=pre, strideA=0,
                 strideB=k * post, strideC=j * post)
        return DeviceData(out)

    def matvecWith(self, v):  # data/__init__.py:257-259
        return v.absorbMatrixAt(0, self)

    # -- factorisations: device kernels live in carcassonne_b200/linalg.py -------------------------------
    def normalizeAxis(self, axis, sqrt_svals=False, dont_recip_under=1e-14):
        from . import linalg
        return linalg.normalize_axis(self, axis, sqrt_svals, dont_recip_under)

    def normalizeAxisAndDenormalize(self, axis_to_norm, axis_to_denorm, data_to_denormalize=None, sqrt_svals=False,
                                    dont_recip_under=1e-14):  # data/__init__.py:302-310
        if data_to_denormalize is None:
            data_to_denormalize = self
        if self.shape[axis_to_norm] != data_to_denormalize.shape[axis_to_denorm]:
            raise ValueError("Normalized axis and denormalized axis have different sizes ({} != {}).".format(
                self.shape[axis_to_norm], data_to_denormalize.shape[axis_to_denorm]))
        normalized_data, _, denormalizer = self.normalizeAxis(axis_to_norm, sqrt_svals, dont_recip_under)
        return normalized_data, data_to_denormalize.absorbMatrixAt(axis_to_denorm, denormalizer)

    def qr(self, mode="full"):
        from . import linalg
        return linalg.qr(self, mode)

    def svd(self, full_matrices=True):
        from . import linalg
        return linalg.svd(self, full_matrices)

    def unitize(self):
        from . import linalg
        return linalg.unitize(self)


DeviceData.I = None
DeviceData.X = None
DeviceData.Y = None
DeviceData.Z = None


def _init_constants():
    """Pauli constants (data/__init__.py:360-363); created lazily because they need a CUDA device."""
    if DeviceData.I is None:
        DeviceData.I = DeviceData.fromArray(np.array([[1, 0], [0, 1]], dtype=np.complex128))
        DeviceData.X = DeviceData.fromArray(np.array([[0, 1], [1, 0]], dtype=np.complex128))
        DeviceData.Y = DeviceData.fromArray(np.array([[0, -1j], [1j, 0]], dtype=np.complex128))
        DeviceData.Z = DeviceData.fromArray(np.array([[1, 0], [0, -1]], dtype=np.complex128))
