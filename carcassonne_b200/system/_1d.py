"""The 1D system: an MPS center tensor between a left and a right environment, driven by an MPO -- the device twin
of the reference's ``carcassonne/system/_1d.py``.  It exists as an independent in-process oracle for 2D runs driven
along one axis (the reference's ``system/_1d2d.py`` pattern): same policies, same run loop, every tensor on device.
"""
from copy import copy

import numpy as np

from ..data import DeviceData
from ..tensors import _1d as t1
from ..utils import relaxOver
from .base import BaseSystem


def _outer(*vectors):
    out = np.asarray(vectors[0], dtype=np.complex128)
    for v in vectors[1:]:
        out = np.multiply.outer(out, np.asarray(v, dtype=np.complex128))
    return out


class System(BaseSystem):
    def __init__(self, right_operator_boundary, left_operator_boundary, operator_center_data, state_center_data=None,
                 right_state_boundary=None, left_state_boundary=None):
        BaseSystem.__init__(self)
        as_array = lambda x: x.toArray() if hasattr(x, "toArray") else np.asarray(x, dtype=np.complex128)
        rob, lob = as_array(right_operator_boundary), as_array(left_operator_boundary)
        self.right_operator_boundary = DeviceData.fromArray(rob)
        self.left_operator_boundary = DeviceData.fromArray(lob)
        rsb = [1] if right_state_boundary is None else as_array(right_state_boundary)
        lsb = [1] if left_state_boundary is None else as_array(left_state_boundary)
        self.right_environment = DeviceData.fromArray(_outer(rob, rsb, np.conj(rsb)))
        self.left_environment = DeviceData.fromArray(_outer(lob, lsb, np.conj(lsb)))
        self.operator_center_data = operator_center_data if isinstance(operator_center_data, DeviceData) \
            else DeviceData.fromArray(as_array(operator_center_data))
        if state_center_data is None:
            state_center_data = DeviceData.newTrivial((1, 1, self.operator_center_data.shape[2]))
        elif not isinstance(state_center_data, DeviceData):
            state_center_data = DeviceData.fromArray(as_array(state_center_data))
        self.setStateCenter(state_center_data)
        self.just_increased_bandwidth = False
        assert self.left_operator_boundary.ndim == 1 and self.right_operator_boundary.ndim == 1
        assert self.left_environment.ndim == 3 and self.right_environment.ndim == 3
        assert self.operator_center_data.ndim == 4 and self.state_center_data.ndim == 3

    @classmethod
    def newRandom(cls, operator_dimension, state_dimension, physical_dimension):
        """reference system/_1d.py:15-25 (draws from the host NumPy stream in the reference's order)."""
        from ..utils import crand
        rob, lob = crand(operator_dimension), crand(operator_dimension)
        op = crand(operator_dimension, operator_dimension, physical_dimension, physical_dimension)
        op = op + op.transpose(0, 1, 3, 2).conj()
        return cls(rob, lob, op, crand(state_dimension, state_dimension, physical_dimension),
                   crand(state_dimension), crand(state_dimension))

    def __copy__(self):
        other = object.__new__(type(self))
        BaseSystem.__init__(other)
        other.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_policies"})
        return other

    # -- expectation ------------------------------------------------------------------------------------------------
    def formExpectationMultiplier(self):
        return t1.formExpectationMultiplier(self.right_environment, self.left_environment, self.operator_center_data)

    def formExpectationMatrix(self):
        return self.formExpectationMultiplier().formMatrix()

    def computeScalarUsingMultiplier(self, multiply):
        return self.state_center_data_conj.contractWithAlongAll(multiply(self.state_center_data))

    def computeExpectation(self):
        return self.computeScalarUsingMultiplier(self.formExpectationMultiplier()).real

    def computeEstimatedOneSiteExpectation(self, direction=0):
        system = copy(self)
        before = system.computeExpectation()
        system.contractTowards(direction)
        return system.computeExpectation() - before

    def computeOneSiteExpectation(self):
        """Energy per site in the thermodynamic limit from the dominant Jordan block of the MPO transfer matrix
        (reference system/_1d.py:61-97 -> utils.computeAbsoluteLimitingLinearCoefficient, utils.py:333-362).  The
        transfer matrices are built on device; the eigen-analysis of the (operator bond x D^2)-sized matrix is host
        LAPACK like the reference's."""
        from scipy.linalg import eigvals, svd
        normalized = self.state_center_data.normalizeAxis(1)[0]
        normalized_conj = normalized.conj()
        O = self.operator_center_data
        D = normalized.shape[0]
        # TO[(o',a,b),(o,s,t)] = sum_{q,p} O[o,o',q,p] S[s,a,p] S*[t,b,q];  TN[(a,b),(s,t)] = sum_p S[s,a,p] S*[t,b,p]
        t = O.contractWith(normalized, (3,), (2,))                    # [o, o', q, s, a]
        t = t.contractWith(normalized_conj, (2,), (2,))               # [o, o', s, a, t, b]
        TO = t.join((1, 3, 5), (0, 2, 4)).toArray()
        TN = normalized.contractWith(normalized_conj, (2,), (2,)).join((1, 3), (0, 2)).toArray()
        n = TO.shape[0]
        matrix = TO.T                                                  # rows = images of unit vectors, as the reference
        evals = eigvals(matrix)
        lam = evals[np.argmax(abs(evals))]
        shifted = matrix - lam * np.identity(n)
        ovecs = svd(shifted @ shifted)[-1][-2:]
        apply_o = lambda v: TO @ v
        omatrix = np.array([[np.dot(ovecs[i].conj(), apply_o(ovecs[j])) for j in range(2)] for i in range(2)])
        numerator = np.sqrt(np.trace(omatrix.T.conj() @ omatrix) - 2)
        nob = self.left_operator_boundary.shape[0]
        lob, rob = self.left_operator_boundary.toArray(), self.right_operator_boundary.toArray()
        project = lambda boundary: np.tensordot(boundary, ovecs.reshape(2, nob, D, D), (0, 1)).reshape(2, D * D)
        lnvecs, rnvecs = project(lob), project(rob)
        nmatrix = np.array([[np.dot(lnvecs[i].conj(), TN @ rnvecs[j]) for j in range(2)] for i in range(2)])
        denominator = np.sqrt(np.trace(nmatrix.T.conj() @ nmatrix))
        return numerator / denominator

    # -- absorption -------------------------------------------------------------------------------------------------
    def contractLeftUnnormalized(self, state_center_data):
        self.left_environment = t1.absorbCenterOSSIntoLeftEnvironment(
            self.left_environment, self.operator_center_data, state_center_data, state_center_data.conj())

    def contractRightUnnormalized(self, state_center_data=None):
        if state_center_data is None:
            state_center_data = self.state_center_data
        self.right_environment = t1.absorbCenterOSSIntoRightEnvironment(
            self.right_environment, self.operator_center_data, state_center_data, state_center_data.conj())

    def contractLeftNormalized(self, state_center_data):
        if state_center_data.shape[1] != self.left_environment.shape[1]:
            raise ValueError("state dimension of the left environment ({}) does not match the left dimension of the "
                             "center state ({})".format(self.left_environment.shape[1], state_center_data.shape[1]))
        self.contractLeftUnnormalized(state_center_data.normalizeAxis(0)[0])

    def contractRightNormalized(self, state_center_data):
        if state_center_data.shape[0] != self.right_environment.shape[1]:
            raise ValueError("state dimension of the right environment ({}) does not match the right dimension of the "
                             "center state ({})".format(self.right_environment.shape[1], state_center_data.shape[0]))
        self.contractRightUnnormalized(state_center_data.normalizeAxis(1)[0])

    def contractUnnormalizedTowards(self, direction, state_center_data=None):
        if state_center_data is None:
            state_center_data = self.state_center_data
        if direction == 0:
            self.contractRightUnnormalized(state_center_data)
        elif direction == 1:
            self.contractLeftUnnormalized(state_center_data)
        else:
            raise ValueError("Direction must be 0 for right or 1 for left, not {}.".format(direction))

    def contractNormalizedTowards(self, direction, state_center_data=None):
        if state_center_data is None:
            state_center_data = self.state_center_data
        if direction == 0:
            self.contractRightNormalized(state_center_data)
        elif direction == 1:
            self.contractLeftNormalized(state_center_data)
        else:
            raise ValueError("Direction must be 0 for right or 1 for left, not {}.".format(direction))

    def contractTowards(self, direction):
        isometry, _, denormalizer = self.state_center_data.normalizeAxis(1 - direction)
        self.contractUnnormalizedTowards(direction, isometry)
        self.setStateCenter(self.state_center_data.normalizeAxis(direction)[0].absorbMatrixAt(direction, denormalizer))

    # -- optimisation / bandwidth -----------------------------------------------------------------------------------
    def increaseBandwidth(self, direction=0, by=None, to=None, do_as_much_as_possible=False, enlargeners=None):
        if direction != 0:
            raise ValueError("Direction for bandwidth increase must be 0, not {}.".format(direction))
        result = self._increaseBandwidth(0, by, to, do_as_much_as_possible, enlargeners)
        self.just_increased_bandwidth = False       # the 1D system has no such invariant (reference system/_1d.py)
        return result

    def minimizeExpectation(self):
        self.setStateCenter(relaxOver(initial=self.state_center_data,
                                      expectation_multiplier=self.formExpectationMultiplier(),
                                      maximum_number_of_multiplications=100))

    def minimizeExpectationUsingFullEigensolver(self):
        from scipy.linalg import eigh
        evals, evecs = eigh(self.formExpectationMatrix().toArray())
        self.setStateCenter(DeviceData.fromArray(evecs[:, 0].reshape(self.state_center_data.shape)))
        return evals[0]

    def setStateCenter(self, state_center_data, state_center_data_conj=None):
        self.state_center_data = state_center_data
        self.state_center_data_conj = state_center_data.conj() if state_center_data_conj is None \
            else state_center_data_conj


__all__ = ["System"]
