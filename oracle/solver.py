"""Center-site eigen-solver, compressors -- restated from the reference.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.
"""
import numpy as np
import scipy.linalg as sla
from scipy.sparse.linalg import LinearOperator, eigsh, gmres

from . import linalg as _l


class RelaxFailed(Exception):
    def __init__(self, initial_value, final_value):
        super().__init__("{} --> {}".format(initial_value, final_value))
        self.initial_value = initial_value
        self.final_value = final_value


class Mult:
    """reference utils.py:180-207 (Multiplier): a matvec with its cmac cost, and a way to form its matrix."""

    def __init__(self, shape, multiply, cost_of_multiply, form_matrix, cost_of_form_matrix):
        self.shape = shape
        self.multiply = multiply
        self.cost_of_multiply = cost_of_multiply
        self.form_matrix = form_matrix
        self.cost_of_form_matrix = cost_of_form_matrix

    def __call__(self, v):
        return self.multiply(v)

    def is_cheaper_to_form_matrix(self, n):  # utils.py:203-206
        return n * self.cost_of_multiply > self.cost_of_form_matrix + n * self.shape[0] * self.shape[1]


def relax_over(initial, H, N=None, maximum_number_of_multiplications=None, tolerance=1e-7,
               dimension_of_krylov_space=None, stats=None):
    """reference utils.py:805-878 (relaxOver): restarted Arnoldi of dimension k (default 3) on N^-1 H with
    classical Gram-Schmidt, a non-Hermitian k x k eig, argmin of the real part, restart on the Ritz vector.

    `initial` is a tensor (any shape); H, N are `Mult`s acting on tensors of that shape.  Returns the new
    tensor.  Unlike the reference, `initial` is not normalised in place (utils.py:808-809 mutates the
    caller's array); the returned value is identical.
    """
    shape = initial.shape
    v0 = np.array(initial, dtype=np.complex128).ravel()
    v0 /= sla.norm(v0)
    n = len(v0)
    k = 3 if dimension_of_krylov_space is None else dimension_of_krylov_space

    if N is None:
        solve = lambda x: x
    elif N.is_cheaper_to_form_matrix(10 * 2 * k):
        lu = sla.lu_factor(N.form_matrix())
        solve = lambda x: sla.lu_solve(lu, x)
    else:
        op = LinearOperator(matvec=lambda x: N(x.reshape(shape)).ravel(), shape=(n, n), dtype=np.complex128)

        def solve(x):
            y, info = gmres(op, x)
            assert info == 0
            return y

    if H.is_cheaper_to_form_matrix(2 * k):
        hmat = H.form_matrix()
        apply_h = lambda x: hmat @ x
    else:
        apply_h = lambda x: H(x.reshape(shape)).ravel()

    nmul = [0]

    def multiply(x):
        nmul[0] += 1
        return solve(apply_h(x))

    initial_value = np.vdot(v0, multiply(v0))
    count = 0
    last = None
    complete = k == n
    start = v0
    while True:
        basis = np.zeros((k, n), dtype=np.complex128)
        mbasis = np.zeros((k, n), dtype=np.complex128)
        basis[0] = start
        for i in range(k):
            mbasis[i] = multiply(basis[i])
            if i < k - 1:
                w = mbasis[i] - (basis[:i + 1].conj() @ mbasis[i]) @ basis[:i + 1]
                nrm = sla.norm(w)
                if nrm <= 1e-14:
                    complete = True
                    basis = basis[:i + 1]
                    mbasis = mbasis[:i + 1]
                    break
                basis[i + 1] = w / nrm
        count += k
        small = basis.conj() @ mbasis.T
        evals, evecs = sla.eig(small)
        j = int(np.argmin(evals.real))
        lam = evals[j]
        y = evecs[:, j]
        ritz = y @ basis
        ritz /= sla.norm(ritz)
        done = (
            complete
            or (last is not None and abs(last - lam) <= tolerance)
            or (maximum_number_of_multiplications is not None and count >= maximum_number_of_multiplications)
        )
        if done:
            final_value = np.vdot(ritz, multiply(ritz))
            if stats is not None:
                stats.update(multiplications=nmul[0], counted=count, initial_value=initial_value,
                             final_value=final_value, ritz_value=lam)
            if ((final_value - initial_value) / (abs(final_value) + abs(initial_value))).real > 1 + 1e-7:
                raise RelaxFailed(initial_value, final_value)
            return ritz.reshape(shape)
        start = ritz
        last = lam


def compute_compressor(old, new, multiplier, normalize=False):
    """reference utils.py:268-303 (computeCompressor): top-`new` eigenpairs of a Hermitian PSD matrix."""
    if new < 0:
        raise ValueError("New dimension ({}) must be non-negative.".format(new))
    if new > old:
        raise ValueError("New dimension ({}) must be <= the old dimension ({}).".format(new, old))
    if new == 0:
        z = np.zeros((0, old), dtype=np.complex128)
        return z, z
    if new >= old // 2:
        m = multiplier.form_matrix()
        if tuple(m.shape) != (old, old):
            raise ValueError("Multiplier matrix has shape {} but the old dimension is {}.".format(m.shape, old))
        evals, evecs = sla.eigh(m)
        evals, evecs = evals[-new:], evecs[:, -new:]
    else:
        op = LinearOperator(shape=(old, old), matvec=multiplier, dtype=np.complex128)
        evals, evecs = eigsh(op, k=new)
    evecs = evecs.T
    while new > 0 and abs(evals[new - 1]) < 1e-15:
        new -= 1
    if new == 0:
        raise ValueError("Input is filled with near-zero elements.")
    if normalize:
        ev = np.sqrt(evals).reshape(new, 1)
        return evecs * ev, evecs / ev
    return evecs, evecs


def product_compressor_matrix(Lt, c, Rt):
    """reference compression.py:11-25: the generated formMatrix(L, c*, c^H, c^T, R) with c of shape [old,new]:
    A[(l r),(i n)] = sum L[l,i,j,p] conj(c)[j,m] conj(c)[k,n] c[q,m] R[k,q,p,r]  (rows [L0 R3], cols [L1 c^H_0])."""
    cc = c.conj()
    out = np.einsum("lijp,jm,nk,mq,kqpr->lrin", Lt, cc, cc.T, c.T, Rt, optimize=True)
    l, i = Lt.shape[0], Lt.shape[1]
    r, n = Rt.shape[3], c.shape[1]
    return np.ascontiguousarray(out).reshape(l * r, i * n)


def product_compressor(Lt, Rt, new, initial=None, sweeps=4):
    """reference compression.py:26-45 (computeProductCompressor): alternating least squares for the isometry
    c [new, old] that best preserves L.R when the shared (state, state*) bond pair is projected by c (x) c*.

    `initial` is the random [old, new] draw (compression.py:35 draws it with newRandom); returns c [new, old].
    """
    if Lt.shape[1] != Lt.shape[2]:
        raise ValueError("left inward dimensions do not match (given {})".format(Lt.shape))
    if Rt.shape[0] != Rt.shape[1]:
        raise ValueError("right inward dimensions do not match (given {})".format(Rt.shape))
    if Lt.shape[1] != Rt.shape[1]:
        raise ValueError("left and right shapes are incompatible (given {} and {})".format(Lt.shape, Rt.shape))
    old = Lt.shape[1]
    b = np.tensordot(Lt, Rt, axes=([1, 2, 3], [0, 1, 2])).ravel()
    if initial is None:
        initial = _l.random_complex(np.random, old, new)
    c = _l.unitize(np.asarray(initial, dtype=np.complex128))
    for _ in range(sweeps):
        A = product_compressor_matrix(Lt, c, Rt)
        Ah = A.conj().T
        x, info = gmres(LinearOperator((old * new,) * 2, lambda v: Ah @ (A @ v), dtype=np.complex128), Ah @ b)
        assert info == 0
        c = _l.unitize(x.reshape(old, new))
    return np.ascontiguousarray(c.T)
