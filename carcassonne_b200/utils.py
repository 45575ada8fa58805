"""Solver utilities of the center-site path: ``Multiplier``, ``relaxOver``, ``computeCompressor``, index helpers
and the exception types -- the names and semantics of the reference's ``carcassonne/utils.py`` for this path.

``relaxOver`` hands the whole restarted-Arnoldi iteration to ``carc_relax`` (csrc/solver.cu): state vector, Krylov
basis, Gram-Schmidt coefficients and the k x k eigenproblem all live on the device; only the stopping rule is
evaluated on the host from one small read-back per restart.
"""
import ctypes as C
from math import prod

import numpy as np


# -- exceptions (reference utils.py:13-41) -------------------------------------------------------------------
class DimensionMismatchError(ValueError):
    def __init__(self, left_tensor_number, left_index, left_dimension, right_tensor_number, right_index,
                 right_dimension):
        self.left_tensor_number = left_tensor_number
        self.left_index = left_index
        self.left_dimension = left_dimension
        self.right_tensor_number = right_tensor_number
        self.right_index = right_index
        self.right_dimension = right_dimension
        ValueError.__init__(
            self, "tensor {}'s index {} has dimension {}, whereas tensor {}'s index {} has dimension {}".format(
                left_tensor_number, left_index, left_dimension, right_tensor_number, right_index, right_dimension))


class InvariantViolatedError(Exception):
    pass


class RelaxFailed(Exception):
    def __init__(self, initial_value, final_value):
        Exception.__init__(self, "{} --> {}".format(initial_value, final_value))
        self.initial_value = initial_value
        self.final_value = final_value

    def __repr__(self):
        return "RelaxFailed({},{})".format(self.initial_value, self.final_value)


class UnexpectedTensorRankError(ValueError):
    def __init__(self, tensor_number, expected_rank, actual_rank):
        self.tensor_number = tensor_number
        self.expected_rank = expected_rank
        self.actual_rank = actual_rank
        ValueError.__init__(self, "tensor {} was expected to have rank {} but actually has rank {}".format(
            tensor_number, expected_rank, actual_rank))


class SolverDidNotConverge(AssertionError):
    """The reference's ``assert info == 0`` after GMRES (utils.py:824, compression.py:43)."""


# -- index helpers (reference utils.py:886-892) ----------------------------------------------------------------
def O(i):
    return (i + 2) % 4


def L(i):
    return (i + 1) % 4


def R(i):
    return (i - 1) % 4


def A(d, i):
    return i - 1 if i > d else i


def OA(i):
    return A(i, O(i))


def LA(i):
    return A(i, L(i))


def RA(i):
    return A(i, R(i))


def computeNewDimension(old_dimension, by=None, to=None):
    """reference utils.py:363-374."""
    if by is None and to is None:
        raise ValueError("Either 'by' or 'to' must not be None.")
    if by is not None and to is not None:
        raise ValueError("Both 'by' and 'to' cannot be None.")
    new_dimension = old_dimension + by if by is not None else to
    assert new_dimension >= old_dimension     # as the reference: a bond never shrinks through this door
    return new_dimension


def dropAt(iterable, index):
    return type(iterable)(x for i, x in enumerate(iterable) if i != index)


def randomComplexSample(shape):
    """reference utils.py:795-797: uniform [-1,1) + i[-1,1) from the host NumPy stream (real parts first)."""
    return np.random.random_sample(shape) * 2 - 1 + np.random.random_sample(shape) * 2j - 1j


def crand(*shape):
    return np.random.rand(*shape) * 2 - 1 + np.random.rand(*shape) * 2j - 1j


class Pauli:
    I = np.identity(2, dtype=np.complex128)
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
    Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)


# -- Multiplier (reference utils.py:180-207) ---------------------------------------------------------------------
class Multiplier:
    """A matvec with its cmac cost and a way to form its matrix.  Multipliers built by the tensors layer also
    carry ``device_operator`` (a ``Stage3Operator``), which ``relaxOver`` hands to the device solver."""

    def __init__(self, shape, multiply, cost_of_multiply, formMatrix, cost_of_formMatrix):
        self.shape = shape
        self.multiply = multiply
        self.cost_of_multiply = cost_of_multiply
        self.formMatrix = formMatrix
        self.cost_of_formMatrix = cost_of_formMatrix
        self.device_operator = None

    def __call__(self, vector):
        return self.multiply(vector)

    @classmethod
    def fromMatrix(cls, matrix):
        m, n = matrix.shape
        return cls(matrix.shape, lambda v: matrix.matvecWith(v), m * n, lambda: matrix, 0)

    def isCheaperToFormMatrix(self, estimated_number_of_applications):
        return estimated_number_of_applications * self.cost_of_multiply > \
            self.cost_of_formMatrix + estimated_number_of_applications * self.shape[0] * self.shape[1]


# -- relaxOver (reference utils.py:805-878) ------------------------------------------------------------------------
class _DenseOperator:
    """carc_operator over a dense [n, n] device matrix."""

    def __init__(self, matrix):
        from ._lib import lib, check
        from .data import _ptr
        if matrix.ndim != 2 or matrix.shape[0] != matrix.shape[1]:
            raise ValueError("dense operator needs a square matrix, not {}".format(matrix.shape))
        self.matrix = matrix
        self._finalized = True
        self._handle = C.c_void_p()
        check(lib.carc_operator_create_dense(C.byref(self._handle), _ptr(matrix._t), matrix.shape[0]))

    def close(self):
        from ._lib import lib
        if self._handle:
            lib.carc_operator_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LUFactors:
    """scipy.linalg.lu_factor / lu_solve on device (carc_lu_factor / carc_lu_solve)."""

    def __init__(self, matrix, try_cholesky=False, hermitian_tolerance=1e-11):
        """``try_cholesky``: the caller expects a Hermitian positive definite matrix (the normalization matrix of a
        double-layer environment).  It is then factorised by the device Cholesky (about 3x faster: no pivot search)
        and stored in the same LU form; a matrix that is not Hermitian to ``hermitian_tolerance`` (relative,
        Frobenius) or meets a non-positive pivot goes through the general LU as in the reference."""
        import torch
        from ._lib import lib, check
        from .data import _ptr, _stream
        n = matrix.shape[0]
        self.n = n
        self.piv = torch.empty(n, dtype=torch.int32, device="cuda")
        self.method = "lu"
        self.singular = False
        self.lu = matrix.copy()
        if try_cholesky:
            status = C.c_int(0)
            check(lib.carc_cholesky_factor_as_lu(_ptr(self.lu._t), n, C.c_void_p(self.piv.data_ptr()),
                                                 float(hermitian_tolerance), C.byref(status), _stream()))
            if status.value == 0:
                self.method = "cholesky"
            elif status.value == 1:
                self.lu = matrix.copy()        # the failed attempt overwrote its copy
        if self.method == "lu":
            singular = C.c_int(0)
            check(lib.carc_lu_factor(_ptr(self.lu._t), n, C.c_void_p(self.piv.data_ptr()), C.byref(singular),
                                     _stream()))
            self.singular = bool(singular.value)
        self.inv_blocks = torch.empty(int(lib.carc_lu_inverse_blocks_elems(n)), dtype=torch.complex128, device="cuda")
        check(lib.carc_lu_invert_diagonal_blocks(_ptr(self.lu._t), n, C.c_void_p(self.inv_blocks.data_ptr()), _stream()))

    def solve(self, b):
        from ._lib import lib, check
        from .data import _ptr, _stream
        x = b.copy()
        check(lib.carc_lu_solve_blocks(_ptr(self.lu._t), self.n, C.c_void_p(self.piv.data_ptr()),
                                       C.c_void_p(self.inv_blocks.data_ptr()), _ptr(x._t), _stream()))
        return x

    def solve_reference(self, b):
        """Plain blocked substitution (carc_lu_solve), kept as the cross-check of the fast path."""
        from ._lib import lib, check
        from .data import _ptr, _stream
        x = b.copy()
        check(lib.carc_lu_solve(_ptr(self.lu._t), self.n, C.c_void_p(self.piv.data_ptr()), _ptr(x._t), _stream()))
        return x


def _operator_handle(multiplier, prefer_matrix, keep):
    """Device operator for a Multiplier: its stage-3 term list, or its dense matrix when that is cheaper (or all
    there is)."""
    dev = getattr(multiplier, "device_operator", None)
    if dev is not None and not prefer_matrix:
        if not dev._finalized:
            dev.finalize()
        keep.append(dev)
        return dev._handle
    dense = _DenseOperator(multiplier.formMatrix())
    keep.append(dense)
    return dense._handle


def relaxOver(initial, expectation_multiplier, normalization_multiplier=None, maximum_number_of_multiplications=None,
              tolerance=1e-7, dimension_of_krylov_space=None, gmres_rtol=1e-5, statistics=None):
    """Minimise <v|H|v>/<v|N|v> by a restarted Arnoldi iteration on N^-1 H, on device.

    Branch selection follows the reference's cost model: N^-1 by LU of the dense normalization matrix when
    ``isCheaperToFormMatrix(10*2*k)`` else by GMRES on the operator; H as a dense matrix when
    ``isCheaperToFormMatrix(2*k)`` else through its term list.  Unlike the reference, ``initial`` is not normalised
    in place (utils.py:808-809 mutates the caller's array); the returned tensor is the same."""
    from ._lib import lib, check, CarcError, ERR_RELAX_FAILED, ERR_NO_CONVERGENCE
    from .data import DeviceData, _ptr, _stream
    shape = initial.shape
    n = prod(shape)
    k = 3 if dimension_of_krylov_space is None else int(dimension_of_krylov_space)
    keep = []
    lu = None
    n_handle = None
    if normalization_multiplier is not None:
        if normalization_multiplier.isCheaperToFormMatrix(10 * 2 * k) or \
                getattr(normalization_multiplier, "device_operator", None) is None:
            lu = LUFactors(normalization_multiplier.formMatrix(), try_cholesky=True)
            keep.append(lu)
        else:
            n_handle = _operator_handle(normalization_multiplier, False, keep)
    h_handle = _operator_handle(expectation_multiplier, expectation_multiplier.isCheaperToFormMatrix(2 * k), keep)
    v = initial.copy()
    info = (C.c_double * 9)()
    rc = lib.carc_relax(h_handle, n_handle, _ptr(lu.lu._t) if lu else None,
                        C.c_void_p(lu.piv.data_ptr()) if lu else None,
                        C.c_void_p(lu.inv_blocks.data_ptr()) if lu else None, _ptr(v._t),
                        int(maximum_number_of_multiplications or 0), float(tolerance), k, float(gmres_rtol), 20, 0,
                        info, _stream())
    if statistics is not None:
        statistics.update(initial_value=complex(info[0], info[1]), final_value=complex(info[2], info[3]),
                          ritz_value=complex(info[4], info[5]), counted=int(info[6]), multiplications=int(info[7]),
                          gmres_iterations=int(info[8]),
                          normalization=lu.method if lu else ("gmres" if n_handle else None))
    if rc == ERR_RELAX_FAILED:
        raise RelaxFailed(complex(info[0], info[1]), complex(info[2], info[3]))
    if rc == ERR_NO_CONVERGENCE:
        raise SolverDidNotConverge(lib.carc_last_error().decode("utf-8", "replace"))
    check(rc)
    del keep
    return DeviceData(v._t.reshape(shape))


# -- compressors (reference utils.py:268-321) ----------------------------------------------------------------------
def _dominant_eigenpairs(multiplier, size, count, dtype):
    """(eigenvalues ascending, eigenvectors as columns) of the ``count`` dominant eigenpairs of a Hermitian positive
    semi-definite operator of dimension ``size``.  Dense LAPACK when at least about half the spectrum is wanted -- the
    reference's switch-over, new >= old // 2 (utils.py:279) -- otherwise ARPACK on the matvec, as the reference does."""
    from scipy.linalg import eigh
    from scipy.sparse.linalg import LinearOperator, eigsh
    if count < size // 2:
        return eigsh(LinearOperator(shape=(size, size), matvec=multiplier, dtype=dtype), k=count)
    matrix = multiplier.formMatrix()
    matrix = matrix.toArray() if hasattr(matrix, "toArray") else np.asarray(matrix)
    if matrix.shape != (size, size):
        raise ValueError("Multiplier matrix has shape {} but the old dimension is {}.".format(matrix.shape, size))
    values, vectors = eigh(matrix)
    return values[size - count:], vectors[:, size - count:]


def computeCompressor(old_dimension, new_dimension, multiplier, dtype=np.complex128, normalize=False):
    """Compressor pair for an operator bond (same contract as the reference's ``computeCompressor``, utils.py:268-303;
    a #terms x #terms problem, kept on the host as SURVEY.md section 2 (K17) prescribes): the rows of both returned
    [new, old] matrices are the dominant eigenvectors (transposed, not conjugated) of the Gram operator ``multiplier``;
    with ``normalize`` the first is scaled by the square roots of the eigenvalues and the second by their inverses, so
    that the pair still multiplies to the projector.  Raises ``ValueError`` for an inconsistent dimension and when every
    retained eigenvalue is numerically zero."""
    if new_dimension < 0:
        raise ValueError("New dimension ({}) must be non-negative.".format(new_dimension))
    if new_dimension > old_dimension:
        raise ValueError("New dimension ({}) must be less than or equal to the old dimension ({}).".format(
            new_dimension, old_dimension))
    if new_dimension == 0:
        empty = np.zeros((0, old_dimension), dtype=dtype)
        return empty, empty
    values, vectors = _dominant_eigenpairs(multiplier, old_dimension, new_dimension, dtype)
    if not np.any(np.abs(values) >= 1e-15):
        raise ValueError("Input is filled with near-zero elements.")
    rows = vectors.T
    if not normalize:
        return rows, rows
    weights = np.sqrt(values)[:, None]
    return rows * weights, rows / weights


def computeCompressorForMatrixTimesItsDagger(old_dimension, new_dimension, matrix, normalize=False):
    """Compressor of the Gram operator M^H M of a [other, old] matrix, applied without forming it when ARPACK is used
    (reference utils.py:305-321)."""
    adjoint = matrix.conj().T
    rows = matrix.shape[0]
    gram = Multiplier((old_dimension, old_dimension), lambda v: adjoint @ (matrix @ v), 2 * old_dimension * rows,
                      lambda: adjoint @ matrix, old_dimension * old_dimension * rows)
    return computeCompressor(old_dimension, new_dimension, gram, matrix.dtype, normalize)
