"""Closed-form two-site-operator walks on trivial (bond dimension 1) systems with physical dimension 1..5 -- the
reference's tests/test_two_site_operator.py: product-state expectations and width x height counting pin the
sparse tag algebra exactly."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dd():
    from carcassonne_b200.data import DeviceData, _init_constants
    _init_constants()
    return DeviceData


def herm(rng, d):
    m = rng.uniform(-1, 1, (d, d)) + 1j * rng.uniform(-1, 1, (d, d))
    return m + m.conj().T


def state(rng, d):
    v = rng.uniform(-1, 1, d) + 1j * rng.uniform(-1, 1, d)
    return v / np.linalg.norm(v)


def site_exp(O, v):
    return np.vdot(v, O @ v)


def make(dd, OO_UD=None, OO_LR=None):
    from carcassonne_b200.system import System
    kw = {}
    if OO_UD is not None:
        kw["OO_UD"] = [dd.fromArray(o) for o in OO_UD]
    if OO_LR is not None:
        kw["OO_LR"] = [dd.fromArray(o) for o in OO_LR]
    return System.newTrivialWithSimpleSparseOperator(**kw)


def center(dd, v):
    return dd.fromArray(v.reshape(1, 1, 1, 1, -1))


def close(a, b, tol=1e-10):
    return abs(a - b) <= tol * max(1.0, abs(b))


@pytest.mark.parametrize("d", [1, 2, 3, 5])
def test_no_steps_and_orthogonal_steps_vanish(dd, d):
    rng = np.random.default_rng(d)
    Os = [herm(rng, d) for _ in range(4)]
    s = make(dd, OO_UD=Os[:2], OO_LR=Os[2:])
    assert close(s.computeExpectation(), 0)
    pyrng = random.Random(d)
    s = make(dd, OO_UD=Os[:2])
    for _ in range(pyrng.randint(1, 4)):
        s.contractTowards(pyrng.choice((0, 2)))
    assert close(s.computeExpectation(), 0)
    s = make(dd, OO_LR=Os[2:])
    for _ in range(pyrng.randint(1, 4)):
        s.contractTowards(pyrng.choice((1, 3)))
    assert close(s.computeExpectation(), 0)


@pytest.mark.parametrize("d", [1, 2, 4, 5])
@pytest.mark.parametrize("direction", [0, 1, 2, 3])
def test_one_step_product_expectation(dd, d, direction):
    """reference test_LR_one_step_right/left, test_UD_one_step_up/down: absorb a site in state a, put state b in the
    center -> <O_first>_{towards} <O_second>."""
    rng = np.random.default_rng(10 * d + direction)
    OO = [herm(rng, d), herm(rng, d)]
    horizontal = direction in (0, 2)
    s = make(dd, OO_LR=OO) if horizontal else make(dd, OO_UD=OO)
    absorbed, remaining = state(rng, d), state(rng, d)
    s.setStateCenter(center(dd, absorbed))
    assert close(s.computeNormalization(), 1)
    s.contractTowards(direction)
    s.setStateCenter(center(dd, remaining))
    assert close(s.computeNormalization(), 1)
    # OO_LR = (O_L, O_R): the left site carries O_L; OO_UD = (O_U, O_D) likewise
    if direction in (0, 1):      # absorbed site sits to the right / above
        first, second = (remaining, absorbed) if horizontal else (absorbed, remaining)
    else:
        first, second = (absorbed, remaining) if horizontal else (remaining, absorbed)
    expected = site_exp(OO[0], first) * site_exp(OO[1], second)
    assert close(s.computeExpectation(), expected)


@pytest.mark.parametrize("d", [2, 3])
@pytest.mark.parametrize("direction", [0, 1, 2, 3])
def test_two_steps_chain(dd, d, direction):
    """reference test_LR_two_steps_*/test_UD_two_steps_*: three sites in a row -> two bond energies."""
    rng = np.random.default_rng(100 * d + direction)
    OO = [herm(rng, d), herm(rng, d)]
    horizontal = direction in (0, 2)
    s = make(dd, OO_LR=OO) if horizontal else make(dd, OO_UD=OO)
    a, b, c = state(rng, d), state(rng, d), state(rng, d)
    s.setStateCenter(center(dd, a))
    s.contractTowards(direction)
    s.setStateCenter(center(dd, b))
    s.contractTowards(direction)
    s.setStateCenter(center(dd, c))
    # sites along the axis, in the order (left to right) or (up to down)
    if horizontal:
        chain = [c, b, a] if direction == 0 else [a, b, c]
    else:
        chain = [a, b, c] if direction == 1 else [c, b, a]
    expected = sum(site_exp(OO[0], chain[i]) * site_exp(OO[1], chain[i + 1]) for i in range(2))
    assert close(s.computeExpectation(), expected)
    assert close(s.computeNormalization(), 1)


@pytest.mark.parametrize("d", [1, 2, 3, 5])
@pytest.mark.parametrize("which", ["LR", "UD", "both"])
def test_many_steps_uniform(dd, d, which):
    """reference test_*_many_steps_uniform: a uniform product state on a width x height patch has
    (#horizontal bonds) E_LR + (#vertical bonds) E_UD."""
    rng = np.random.default_rng(7 * d + len(which))
    pyrng = random.Random(7 * d + len(which))
    LR, UD = [herm(rng, d), herm(rng, d)], [herm(rng, d), herm(rng, d)]
    s = make(dd, OO_LR=LR if which != "UD" else None, OO_UD=UD if which != "LR" else None)
    v = state(rng, d)
    s.setStateCenter(center(dd, v))
    e_lr = site_exp(LR[0], v) * site_exp(LR[1], v) if which != "UD" else 0
    e_ud = site_exp(UD[0], v) * site_exp(UD[1], v) if which != "LR" else 0
    width = height = 1
    for _ in range(pyrng.randint(1, 6)):
        direction = pyrng.randint(0, 3)
        s.contractTowards(direction)
        if direction in (0, 2):
            width += 1
        else:
            height += 1
    expected = (width - 1) * height * e_lr + width * (height - 1) * e_ud
    assert close(s.computeExpectation(), expected, 1e-9)
    assert close(s.computeNormalization(), 1)


# -- operator-bond compression that survives absorption --------------
def _twin_systems(dd, model):
    from carcassonne_b200.system import System
    if model == "tfim":
        make = lambda: System.newTrivialWithSparseOperator(Os=[-dd.Z], OO_UDs=[(dd.X, -0.3 * dd.X)],  # noqa: E731
                                                           OO_LRs=[(dd.X, -0.3 * dd.X)])
    else:
        pairs = [(dd.X, dd.X), (dd.Y, dd.Y), (dd.Z, dd.Z)]
        make = lambda: System.newTrivialWithSparseOperator(OO_UDs=pairs, OO_LRs=pairs)  # noqa: E731
    return make(), make()


@pytest.mark.parametrize("model", ["tfim", "heisenberg"])
def test_edge_operator_compression_survives_absorption(dd, model):
    """Twenty absorb + compress rounds: a system whose two-site halves are folded, edge by edge, into compressed operator
    bonds after EVERY absorption keeps <H> and <N> of its uncompressed twin to 1e-9, and its number of stage-3 terms
    stops growing while the twin's grows by 2 (TFIM) / 6 (Heisenberg) per round.  The reference's per-junction routine
    (system/_2d.py:229-363) is only invariant for a single application (tests/test_system.py:82-178)."""
    from carcassonne_b200.sparse import TwoSiteOperator, TwoSiteOperatorCompressed
    plain, folded = _twin_systems(dd, model)
    rng = np.random.default_rng(11)
    terms_plain, terms_folded = [], []
    for round_ in range(20):
        v = rng.uniform(-1, 1, plain.state_center_data.shape) + 1j * rng.uniform(-1, 1, plain.state_center_data.shape)
        v /= np.linalg.norm(v)
        for system in (plain, folded):
            system.setStateCenter(dd.fromArray(v))
            system.contractTowards(round_ % 4)
        for edge in range(4):
            n = folded.edgeTwoSiteOperatorBondDimension(edge)
            if n:
                folded.compressEdgeTwoSiteOperators(edge, n)          # full rank: exact
        e0, n0 = plain.computeExpectationAndNormalization()
        e1, n1 = folded.computeExpectationAndNormalization()
        assert abs(e1 - e0) <= 1e-9 * max(1.0, abs(e0)), (round_, e0, e1)
        assert abs(n1 - n0) <= 1e-9 * abs(n0), (round_, n0, n1)
        terms_plain.append(len(plain.formExpectationMultiplier().terms))
        terms_folded.append(len(folded.formExpectationMultiplier().terms))
    assert any(isinstance(t, TwoSiteOperatorCompressed) for c in folded.corners for t in c)
    assert not any(isinstance(t, TwoSiteOperator) and t.direction in (0, 1) for c in folded.corners for t in c)
    assert terms_plain[-1] >= terms_plain[4] + 6            # the twin keeps growing (cross terms of ever deeper halves) ...
    assert terms_folded[-1] == terms_folded[8]              # ... the folded system does not
    assert terms_folded[-1] < terms_plain[-1]
    assert abs(e0) > 1e-3


def test_edge_operator_compression_with_bonds_and_state_compression(dd):
    """The same invariance with non-trivial bonds: chi = D = 2, every absorption followed by the state-bond compression
    back to chi = 2 (identical on both twins: the compressors come from the Identity tensors and the same host draws),
    the folded twin additionally by the per-edge operator compression at full rank."""
    from carcassonne_b200 import synthetic
    twins = [synthetic.device_system(2, 2, J=0.4, seed=21) for _ in range(2)]
    plain, folded = twins
    rng = np.random.default_rng(12)
    for round_ in range(8):
        v = rng.uniform(-1, 1, plain.state_center_data.shape) + 1j * rng.uniform(-1, 1, plain.state_center_data.shape)
        v /= np.linalg.norm(v)
        state = np.random.get_state()
        for system in twins:
            np.random.set_state(state)
            system.setStateCenter(dd.fromArray(v))
            system.contractTowards(round_ % 4)
            for corner_id in range(4):
                for direction in range(2):
                    system.compressCornerStateTowards(corner_id, direction, 2)
        for edge in range(4):
            n = folded.edgeTwoSiteOperatorBondDimension(edge)
            if n:
                folded.compressEdgeTwoSiteOperators(edge, n)
        e0, n0 = plain.computeExpectationAndNormalization()
        e1, n1 = folded.computeExpectationAndNormalization()
        assert abs(e1 - e0) <= 1e-8 * max(1.0, abs(e0)), (round_, e0, e1)
        assert abs(n1 - n0) <= 1e-8 * abs(n0), (round_, n0, n1)


@pytest.mark.parametrize("direction", [0, 1])
def test_run_with_operator_compression_policy(dd, direction):
    """The reference's transverse-Ising chain run (tests/test_simulator_2d_in_1d.py:36-47) with the operator-compression
    slot filled: the per-edge policy, truncating to 3 channels, applied after every absorption of the whole run.  The
    energy still lands on the known answer and the number of stage-3 terms stays bounded."""
    from carcassonne_b200 import policies as pol
    from carcassonne_b200.system import System
    np.random.seed(7 + direction)
    random.seed(7 + direction)
    kw = {"OO_LR" if direction == 0 else "OO_UD": [dd.X, -0.01 * dd.X]}
    system = System.newTrivialWithSimpleSparseOperator(O=-dd.Z, **kw)
    system.setPolicy("operator compression", pol.ConstantOperatorCompressionPolicy(3))
    system.setPolicy("sweep convergence", pol.RelativeStateDifferenceThresholdConvergencePolicy(1e-5))
    system.setPolicy("run convergence", pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-7))
    system.setPolicy("bandwidth increase", pol.OneDirectionIncrementBandwidthIncreasePolicy(direction, 2))
    system.setPolicy("contraction", pol.RepeatPatternContractionPolicy([0 + direction, 2 + direction]))
    system.runUntilConverged()
    energy = system.computeOneSiteExpectation()
    assert abs(energy - (-1.0000250001562545)) < 1e-6
    assert len(system.formExpectationMultiplier().terms) <= 12
