"""Host logic of the run loop and its policies (reference carcassonne/policies.py, system/base.py:58-111) against
stand-in systems: no device call is made, so these run in the CPU suite."""
import pytest

from carcassonne_b200 import policies as pol
from carcassonne_b200.system.base import BaseSystem
from carcassonne_b200.utils import RelaxFailed


class Recorder(BaseSystem):
    """A system whose 'energy' follows a scripted sequence; records the calls the policies make."""

    def __init__(self, energies):
        super().__init__()
        self.energies = list(energies)
        self.calls = []
        self.bandwidth = 1

    def minimizeExpectation(self):
        self.calls.append("minimize")

    def contractTowards(self, direction):
        self.calls.append(("contract", direction))

    def compressCornerStateTowards(self, corner_id, direction, new_dimension):
        self.calls.append(("compress", corner_id, direction, new_dimension))

    def increaseBandwidth(self, direction, by=None, to=None, do_as_much_as_possible=False, enlargeners=None):
        self.calls.append(("grow", direction, by))
        self.bandwidth += by

    def computeOneSiteExpectation(self):
        return self.energies.pop(0) if len(self.energies) > 1 else self.energies[0]


def test_templates_are_shared_but_state_is_per_binding():
    template = pol.RepeatPatternContractionPolicy([0, 2])
    a, b = Recorder([1.0]), Recorder([1.0])
    a.setPolicy("contraction", template)
    b.setPolicy("contraction", template)
    for _ in range(3):
        a._applyPolicy("contraction")
    b._applyPolicy("contraction")
    assert a.calls == [("contract", 0), ("contract", 2), ("contract", 0)]
    assert b.calls == [("contract", 0)]
    assert template.position == 0                     # writes stayed on the bindings
    a._resetPolicy("contraction")
    a._applyPolicy("contraction")
    assert a.calls[-1] == ("contract", 0)


def test_empty_contraction_pattern_is_an_error():
    s = Recorder([1.0])
    s.setPolicy("contraction", pol.RepeatPatternContractionPolicy([]))
    with pytest.raises(ValueError):
        s._applyPolicy("contraction")


def test_constant_state_compression_visits_every_corner_and_direction():
    s = Recorder([1.0])
    s.setPolicy("state compression", pol.ConstantStateCompressionPolicy(5))
    s._applyPolicy("state compression")
    assert s.calls == [("compress", c, d, 5) for c in range(4) for d in range(2)]


def test_sweep_and_run_loops_follow_the_reference_order():
    # energies seen by the convergence policies: sweep 1 converges at the third update, the run after two sweeps
    s = Recorder([-1.0, -1.5, -1.5 - 1e-9, -1.5 - 1e-9, -1.5 - 1e-9, -1.5 - 1e-9])
    s.setPolicy("sweep convergence", pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-6))
    s.setPolicy("run convergence", pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-6))
    s.setPolicy("contraction", pol.RepeatPatternContractionPolicy(range(4)))
    s.setPolicy("state compression", pol.ConstantStateCompressionPolicy(2))
    s.setPolicy("bandwidth increase", pol.AllDirectionsIncrementBandwidthIncreasePolicy(1))
    s.runUntilConverged()
    kinds = [c if isinstance(c, str) else c[0] for c in s.calls]
    # first sweep: minimize, then (contract, 8 compressions, minimize) until two successive energies agree
    assert kinds[:11] == ["minimize", "contract"] + ["compress"] * 8 + ["minimize"]
    assert ("grow", 0, 1) in s.calls and ("grow", 1, 1) in s.calls
    assert s.number_of_sweeps == 2
    assert s.calls[1] == ("contract", 0)
    # the contraction pattern restarts with every sweep
    first_contract_of_second_sweep = s.calls[max(i for i, c in enumerate(s.calls) if c[0] == "grow") + 2]
    assert first_contract_of_second_sweep == ("contract", 0)


def test_failed_relaxation_is_swallowed_inside_a_sweep_but_not_at_its_start():
    class Failing(Recorder):
        def __init__(self, fail_on):
            super().__init__([-1.0, -1.2, -1.2, -1.2])
            self.fail_on = fail_on
            self.count = 0

        def minimizeExpectation(self):
            self.count += 1
            if self.count == self.fail_on:
                raise RelaxFailed(0.0, 1.0)
            super().minimizeExpectation()

    for fail_on in (2, 1):
        s = Failing(fail_on)
        s.setPolicy("sweep convergence", pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-6))
        s.setPolicy("contraction", pol.RepeatPatternContractionPolicy([0]))
        if fail_on == 1:
            with pytest.raises(RelaxFailed):          # the reference only guards the loop body (base.py:104-110)
                s.sweepUntilConverged()
        else:
            s.sweepUntilConverged()
            assert s.count >= 3


def test_relative_state_difference_policy_with_stand_in_tensors():
    class Vec:
        def __init__(self, *x):
            self.x = x
            self.shape = (len(x),)

        def __add__(self, o):
            return Vec(*[a + b for a, b in zip(self.x, o.x)])

        def __sub__(self, o):
            return Vec(*[a - b for a, b in zip(self.x, o.x)])

        def norm(self):
            return sum(abs(a) ** 2 for a in self.x) ** 0.5

    class S:
        state_center_data = Vec(1.0, 0.0)

    system = S()
    p = pol.RelativeStateDifferenceThresholdConvergencePolicy(1e-3).createBindingToSystem(system)
    p.reset()
    p.update()
    assert p.converged() is False                      # nothing to compare with yet
    system.state_center_data = Vec(1.0, 0.5)
    p.update()
    assert p.converged() is False
    system.state_center_data = Vec(1.0, 0.5 + 1e-5)
    p.update()
    assert p.converged() is True
    system.state_center_data = Vec(1.0, 0.5, 0.0)      # a bandwidth increase changes the shape: never converged
    p.update()
    assert p.converged() is False


def test_hook_policy_accepts_a_bound_method_callback():
    """A callable stored on a policy that is a bound method of ANOTHER object must reach the policy untouched (the
    reference's Proxy forwards plain attributes, policies.py:12-44); only the template's own methods are rebound."""
    class Spy:
        def __init__(self):
            self.seen = []

        def hook(self, system):
            self.seen.append(system)
            return "called"

    spy = Spy()
    system = Recorder([1.0])
    binding = pol.HookPolicy(spy.hook).createBindingToSystem(system)
    assert binding.apply() == "called"
    assert spy.seen == [system]
    assert binding.callback.__self__ is spy


def test_compression_policy_rejects_bonds_beyond_the_device_svd_up_front():
    from carcassonne_b200.linalg import MAX_SMALL
    pol.ConstantStateCompressionPolicy(MAX_SMALL)
    with pytest.raises(NotImplementedError):
        pol.ConstantStateCompressionPolicy(MAX_SMALL + 1)
