"""Synthetic environments for benchmarks and large-size parity properties.

Random corner / side tensors with the double-layer structure a real environment has -- each is a sum of
``rank`` ket (x) bra products -- so the normalization matrix is Hermitian positive semi-definite and the
generalised eigenproblem of ``minimizeExpectation`` is well posed at any (chi, D), unlike the uniformly random
Hermitian tensors of ``System.newRandom``.  Draws come from a seeded NumPy ``Generator`` on the host and are
returned as plain ndarrays, so the device system and the CPU oracle can be fed the same numbers.
"""
import numpy as np


def _crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def double_layer_environment(chi, D, d=2, rank=2, seed=0):
    """-> (corners, sides, center) as ndarrays: corners [chi,chi,1,chi,chi,1], sides [chi,chi,1,chi,chi,1,D,D],
    center [D,D,D,D,d] (normalised)."""
    rng = np.random.default_rng(seed)
    corners, sides = [], []
    for _ in range(4):
        kets = _crand(rng, rank, chi, chi, D)
        side = np.einsum("kadg,kbeh->abdegh", kets, kets.conj()).reshape(chi, chi, 1, chi, chi, 1, D, D)
        sides.append(np.ascontiguousarray(side / np.linalg.norm(side) * chi))
    for _ in range(4):
        kets = _crand(rng, rank, chi, chi)
        corner = np.einsum("kad,kbe->abde", kets, kets.conj()).reshape(chi, chi, 1, chi, chi, 1)
        corners.append(np.ascontiguousarray(corner / np.linalg.norm(corner) * chi))
    center = _crand(rng, D, D, D, D, d)
    center /= np.linalg.norm(center)
    return corners, sides, center


def tfim_operator_arrays(J=1.0):
    """(Os, OO_UDs, OO_LRs) of the transverse-field Ising Hamiltonian -Z - J XX as ndarrays."""
    Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    return [-Z], [(X, -J * X)], [(X, -J * X)]


def heisenberg_operator_arrays():
    X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
    Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
    Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
    pairs = [(X, X), (Y, Y), (Z, Z)]
    return [], list(pairs), list(pairs)


def device_system(chi, D, model="tfim", J=1.0, rank=2, seed=0):
    """A ``System`` on the current CUDA device holding the synthetic environment (Identity tags only)."""
    from .data import DeviceData
    from .sparse import Identity
    from .system import System
    from .sparse import makeSparseOperator
    corners, sides, center = double_layer_environment(chi, D, 2, rank, seed)
    Os, UDs, LRs = tfim_operator_arrays(J) if model == "tfim" else heisenberg_operator_arrays()
    dev = DeviceData.fromArray
    operator = makeSparseOperator([dev(o) for o in Os], [(dev(a), dev(b)) for a, b in UDs],
                                  [(dev(a), dev(b)) for a, b in LRs])
    return System([{Identity(): dev(c)} for c in corners], [{Identity(): dev(s)} for s in sides], dev(center),
                  operator)
