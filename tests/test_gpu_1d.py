"""The 1D (MPS/MPO) system on device: recipes against einsum, multiplier against its matrix, and the reference's
own 1D simulator runs (tests/test_simulator_1d.py): transverse Ising, XY, Heisenberg -- plus agreement with the 2D
system driven along one axis."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
I2 = np.eye(2, dtype=complex)


@pytest.fixture(scope="module")
def dd():
    from carcassonne_b200.data import DeviceData, _init_constants
    _init_constants()
    return DeviceData


def crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


def relerr(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_1d_recipes(dd):
    from carcassonne_b200.tensors import _1d as t1
    rng = np.random.default_rng(0)
    o, s, a, d = 3, 4, 5, 2
    L, R = crand(rng, o, s, s), crand(rng, o, s, s)
    O = crand(rng, o, o, d, d)
    S = crand(rng, a, s, d)        # [right, left, phys] with left bond s
    Sr = crand(rng, s, a, d)       # right bond s
    dev = dd.fromArray
    out = t1.absorbCenterOSSIntoLeftEnvironment(dev(L), dev(O), dev(S), dev(S.conj())).toArray()
    assert relerr(out, np.einsum("ost,uoqp,asp,btq->uab", L, O, S, S.conj())) < 1e-13
    out = t1.absorbCenterOSSIntoRightEnvironment(dev(R), dev(O), dev(Sr), dev(Sr.conj())).toArray()
    assert relerr(out, np.einsum("ost,ouqp,sap,tbq->uab", R, O, Sr, Sr.conj())) < 1e-13
    L2, R2 = crand(rng, s, s), crand(rng, s, s)
    out = t1.absorbCenterSSIntoLeftEnvironment(dev(L2), dev(S), dev(S.conj())).toArray()
    assert relerr(out, np.einsum("st,asp,btp->ab", L2, S, S.conj())) < 1e-13
    out = t1.absorbCenterSSIntoRightEnvironment(dev(R2), dev(Sr), dev(Sr.conj())).toArray()
    assert relerr(out, np.einsum("st,sap,tbp->ab", R2, Sr, Sr.conj())) < 1e-13
    Rm, Lm = crand(rng, o, s, s), crand(rng, o, a, a)
    Sc = crand(rng, s, a, d)
    m = t1.formExpectationMultiplier(dev(Rm), dev(Lm), dev(O))
    ref = np.einsum("osa,utb,ouqp,stp->abq", Rm, Lm, O, Sc)
    assert relerr(m(dev(Sc)).toArray(), ref) < 1e-13
    assert relerr((m.formMatrix().toArray() @ Sc.ravel()).reshape(Sc.shape), ref) < 1e-13


def _policies(system, sweep, run, increment, pattern):
    from carcassonne_b200 import policies as pol
    system.setPolicy("sweep convergence", sweep)
    system.setPolicy("run convergence", run)
    system.setPolicy("bandwidth increase", pol.OneDirectionIncrementBandwidthIncreasePolicy(0, increment))
    system.setPolicy("contraction", pol.RepeatPatternContractionPolicy(pattern))


def test_1d_transverse_ising_matches_2d(dd):
    """reference tests/test_simulator_1d.py:54-73 and tests/test_simulator_2d_in_1d.py:36-47: the same chain through
    the 1D system and through the 2D system driven along one axis."""
    from carcassonne_b200 import policies as pol
    from carcassonne_b200.sparse import makeMPO
    from carcassonne_b200.system import System as System2D
    from carcassonne_b200.system._1d import System as System1D
    np.random.seed(3)
    random.seed(3)
    tensor, right, _, left, _ = makeMPO(I2, Os=[-Z], OOs=[(X, -0.01 * X)])
    s1 = System1D(right, left, tensor, np.ones((1, 1, 2)))
    _policies(s1, pol.RelativeStateDifferenceThresholdConvergencePolicy(1e-5),
              pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-7), 2, [0, 1])
    s1.runUntilConverged()
    e1 = s1.computeOneSiteExpectation()          # the limiting coefficient is an absolute value (utils.py:333)
    assert abs(e1 - 1.0000250001562545) < 1e-6
    s2 = System2D.newTrivialWithSimpleSparseOperator(O=-dd.Z, OO_LR=[dd.X, -0.01 * dd.X])
    _policies(s2, pol.RelativeStateDifferenceThresholdConvergencePolicy(1e-5),
              pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-7), 2, [0, 2])
    s2.runUntilConverged()
    e2 = s2.computeOneSiteExpectation()
    assert abs(e1 - abs(e2)) < 1e-6


def test_1d_heisenberg(dd):
    """reference tests/test_simulator_1d.py:146-170: energy per bond 1/4 - ln 2 to 3 places."""
    from carcassonne_b200 import policies as pol
    from carcassonne_b200.system._1d import System as System1D
    np.random.seed(4)
    tensor = np.zeros((5, 5, 2, 2), dtype=complex)
    tensor[0, 0] = I2
    tensor[0, 1], tensor[0, 2], tensor[0, 3] = X, Y, Z
    tensor[1, 4], tensor[2, 4], tensor[3, 4] = -X, -Y, Z
    tensor[4, 4] = I2
    s = System1D([1, 0, 0, 0, 0], [0, 0, 0, 0, 1], tensor, np.ones((1, 1, 2)))
    _policies(s, pol.RelativeEstimatedOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-5),
              pol.RelativeEstimatedOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-3), 2, [0, 1])
    s.runUntilConverged()
    assert abs(s.computeEstimatedOneSiteExpectation() / 4 - (-0.4431471805599)) < 2e-3


def test_1d_full_eigensolver_cross_check(dd):
    """reference system/_1d.py:176-180: the dense eigen-solve is the system's own cross-check of minimizeExpectation."""
    from carcassonne_b200.sparse import makeMPO
    from carcassonne_b200.system._1d import System as System1D
    rng = np.random.default_rng(6)
    tensor, right, _, left, _ = makeMPO(I2, Os=[-Z], OOs=[(X, -0.7 * X)])
    s = System1D(right, left, tensor, crand(rng, 3, 3, 2), crand(rng, 3), crand(rng, 3))
    h = s.formExpectationMatrix().toArray()
    assert relerr(h, h.conj().T) < 1e-13
    lowest = np.linalg.eigvalsh(h)[0]
    before = s.computeExpectation() / s.state_center_data.norm() ** 2
    for _ in range(30):
        s.minimizeExpectation()
    after = s.computeExpectation()               # relaxOver returns a normalised state
    assert after <= before + 1e-9
    assert abs(after - lowest) < 1e-8 * max(1, abs(lowest))
