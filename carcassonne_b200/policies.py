"""Run-loop policies: when to grow the bandwidth, which way to contract, how hard to compress, when to stop.

Host-side orchestration with the class names and behaviour of the reference's ``carcassonne/policies.py``; the
numerical work each policy triggers runs on the device through the bound ``System``.  A policy object is a
template: ``createBindingToSystem(system)`` returns a bound copy-on-read view whose ``apply`` / ``update`` /
``reset`` / ``converged`` see ``self.system`` (reference policies.py:12-44).
"""
import inspect
import logging

import numpy as np

from .utils import O

log = logging.getLogger(__name__)


class _Binding:
    """View of a policy bound to one system.  Attribute reads fall through to the template policy; attribute writes
    stay on the binding, so one template can serve several systems."""

    def __init__(self, policy, system):
        object.__setattr__(self, "forward", policy)
        object.__setattr__(self, "system", system)

    def __getattr__(self, name):
        forward = object.__getattribute__(self, "forward")
        value = getattr(forward, name)
        # rebind the template's OWN methods so that they see self.system; any other attribute -- including a bound
        # method of some other object stored on the policy, e.g. HookPolicy(recorder.hook) -- is forwarded untouched,
        # as the reference's Proxy does (policies.py:12-44)
        if inspect.ismethod(value) and value.__self__ is forward:
            return value.__func__.__get__(self, type(self))
        return value


class Policy:
    def createBindingToSystem(self, system):
        return _Binding(self, system)


# kept for API compatibility with code that names the reference's proxy classes
Proxy = ApplyProxy = ConvergedProxy = ResetProxy = UpdateProxy = _Binding


# -- bandwidth increase ---------------------------------------------------------------------------------------------
class BandwidthIncreasePolicy(Policy):
    pass


class AllDirectionsIncrementBandwidthIncreasePolicy(BandwidthIncreasePolicy):
    def __init__(self, increment=1):
        self.increment = increment

    def apply(self):
        log.info("Increasing bandwidth in all directions by %s", self.increment)
        for direction in (0, 1):
            self.system.increaseBandwidth(direction, by=self.increment, do_as_much_as_possible=True)


class OneDirectionIncrementBandwidthIncreasePolicy(BandwidthIncreasePolicy):
    def __init__(self, direction, increment=1):
        self.direction = direction
        self.increment = increment

    def apply(self):
        log.info("Increasing bandwidth in direction %s by %s", self.direction, self.increment)
        self.system.increaseBandwidth(self.direction, by=self.increment, do_as_much_as_possible=True)


class AlternatingDirectionsIncrementBandwidthIncreasePolicy(BandwidthIncreasePolicy):
    """The reference's version reads an ``increment`` it never sets (policies.py:50-56) and so cannot run; here the
    increment is a constructor argument (default 1) and the direction alternates between the two axes."""

    def __init__(self, directions=(0, 1), increment=1):
        self.directions = directions
        self.increment = increment
        self.direction = 0

    def apply(self):
        self.system.increaseBandwidth(self.direction, by=self.increment, do_as_much_as_possible=True)
        self.direction = 1 - self.direction


# -- compression ----------------------------------------------------------------------------------------------------
class CompressionPolicy(Policy):
    pass


class ConstantStateCompressionPolicy(CompressionPolicy):
    def __init__(self, new_dimension):
        from .linalg import MAX_SMALL
        if new_dimension > MAX_SMALL:     # the polar projection of the compressor is a device SVD of `new` columns
            raise NotImplementedError("state compression to {} exceeds the device SVD limit of {} columns".format(
                new_dimension, MAX_SMALL))
        self.new_dimension = new_dimension

    def apply(self):
        log.debug("Compressing to %s", self.new_dimension)
        for corner_id in range(4):
            for direction in range(2):
                self.system.compressCornerStateTowards(corner_id, direction, self.new_dimension)


class ConstantOperatorCompressionPolicy(CompressionPolicy):
    """Fills the reference's empty "operator compression" slot (system/base.py:49): folds the two-site halves of every
    edge of the ring into compressed operator bonds of at most ``new_dimension``, which
    bounds the number of stage-3 terms -- otherwise it grows with every absorption round.

    ``per_edge`` (default): ``System.compressEdgeTwoSiteOperators`` -- one channel basis per edge, applied to the four
    tensors whose operator bonds face each other across that edge, so that <H> and <N> stay invariant (at full rank)
    under any number of absorb + compress rounds (tests/test_gpu_two_site.py: 20 rounds, 1e-9).  ``per_edge=False`` is
    the reference's per-junction routine (``compressCornerTwoSiteOperatorTowards``, system/_2d.py:229-363), exact for
    ONE application only: the two ends of a side are rotated at different junctions by different unitaries and stop
    matching once a corner has absorbed the side."""

    def __init__(self, new_dimension, normalize=False, per_edge=True):
        self.new_dimension = new_dimension
        self.normalize = normalize
        self.per_edge = per_edge

    def apply(self):
        if self.per_edge:
            for edge in range(4):
                old_dimension = self.system.edgeTwoSiteOperatorBondDimension(edge)
                if old_dimension:
                    self.system.compressEdgeTwoSiteOperators(edge, min(self.new_dimension, old_dimension), self.normalize)
            return
        for corner_id in range(4):
            for direction in range(2):
                old_dimension = self.system.twoSiteOperatorBondDimension(corner_id, direction)
                if old_dimension:
                    self.system.compressCornerTwoSiteOperatorTowards(
                        corner_id, direction, min(self.new_dimension, old_dimension), self.normalize)


# -- contraction ----------------------------------------------------------------------------------------------------
class ContractionPolicy(Policy):
    pass


class RepeatPatternContractionPolicy(ContractionPolicy):
    def __init__(self, directions):
        self.directions = directions
        self.position = 0

    def apply(self):
        directions = list(self.directions)
        if not directions:
            raise ValueError("An empty sequence of contraction directions was provided! ({})".format(self.directions))
        direction = directions[self.position % len(directions)]
        self.position = self.position % len(directions) + 1
        log.debug("Contracting towards direction %s", direction)
        self.system.contractTowards(direction)

    def reset(self):
        self.position = 0


# -- convergence ----------------------------------------------------------------------------------------------------
class ConvergencePolicy(Policy):
    def reset(self):
        pass

    def update(self):
        pass


def _relative_change(current, last):
    return abs(current - last) / abs(current + last) * 2


class PeriodicyThresholdConvergencePolicy(ConvergencePolicy):
    def __init__(self, threshold, *directions):
        if not directions:
            raise ValueError("at least one direction must be specified")
        self.threshold = threshold
        self.directions = directions

    def converged(self):
        state = self.system.state_center_data
        difference = 0
        for direction in self.directions:
            normalized = state.normalizeAxis(direction)[0]
            denormalizer = state.normalizeAxis(O(direction))[-1]
            difference += (state - normalized.absorbMatrixAt(direction, denormalizer)).norm()
        return difference < self.threshold


class _LastCurrent(ConvergencePolicy):
    def reset(self):
        self.last = None
        self.current = None


class RelativeEstimatedOneSiteExpectationDifferenceThresholdConvergencePolicy(_LastCurrent):
    def __init__(self, threshold, direction=0):
        self.threshold = threshold
        self.direction = direction
        self.last = self.current = None

    def converged(self):
        if self.last is None or self.current is None:
            return None
        magnitude = abs(self.current + self.last)
        return magnitude < 1e-15 or _relative_change(self.current, self.last) < self.threshold

    def update(self):
        self.last = self.current
        self.current = self.system.computeEstimatedOneSiteExpectation(self.direction)


class RelativeExpectationDifferenceDifferenceThresholdConvergencePolicy(ConvergencePolicy):
    def __init__(self, threshold):
        self.threshold = threshold
        self.reset()

    def reset(self):
        self.last_value = self.last_difference = self.current_value = self.current_difference = None

    def converged(self):
        last, current = self.last_difference, self.current_difference
        if last is None or current is None:
            return None
        absolute = abs(current - last)
        return absolute < 1e-15 or absolute / abs(current + last) * 2 < self.threshold

    def update(self):
        self.last_value, self.last_difference = self.current_value, self.current_difference
        self.current_value = self.system.computeExpectation()
        if self.last_value is not None:
            self.current_difference = self.current_value - self.last_value


class RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(_LastCurrent):
    def __init__(self, threshold):
        self.threshold = threshold
        self.last = self.current = None

    def converged(self):
        if self.last is None or self.current is None:
            return None
        if (self.current - self.last).real > self.threshold:
            log.info("Current expectation (%s) is greater than last expectation (%s)!", self.current, self.last)
        absolute = abs(self.current - self.last)
        return absolute < 1e-15 or _relative_change(self.current, self.last) < self.threshold

    def update(self):
        self.last = self.current
        self.current = self.system.computeOneSiteExpectation()


class RelativeStateDifferenceThresholdConvergencePolicy(_LastCurrent):
    """Compares successive center tensors.  The comparison runs on device (two norms) instead of pulling both
    tensors to the host as the reference does (policies.py:217-232)."""

    def __init__(self, threshold):
        self.threshold = threshold
        self.last = self.current = None

    def converged(self):
        if self.last is None or self.current is None or self.last.shape != self.current.shape:
            return False
        magnitude = (self.current + self.last).norm()
        if magnitude < 1e-15:
            return True
        return (self.current - self.last).norm() / magnitude * 2 < self.threshold

    def update(self):
        self.last = self.current
        self.current = self.system.state_center_data


class BoundedConvergencePolicy(ConvergencePolicy):
    """Not in the reference.  Wraps another convergence policy: additionally reports convergence after
    ``maximum_updates`` updates, and raises ``FloatingPointError`` as soon as the wrapped policy's tracked value stops
    being finite -- a NaN compares false against every threshold, so the reference's run loop (base.py:87-111) would
    never return once the environment's norm has under- or overflowed (the reference never
    renormalises a full-2D run)."""

    def __init__(self, policy, maximum_updates):
        self.policy = policy
        self.maximum_updates = maximum_updates

    def createBindingToSystem(self, system):
        binding = _Binding(self, system)
        binding.inner = self.policy.createBindingToSystem(system)
        binding.updates = 0
        return binding

    def reset(self):
        self.inner.reset()
        self.updates = 0

    def update(self):
        self.inner.update()
        self.updates += 1
        value = getattr(self.inner, "current", None)
        if isinstance(value, (int, float, complex, np.number)) and not np.isfinite(value):
            raise FloatingPointError("convergence value is not finite after {} updates: {}".format(self.updates, value))

    def converged(self):
        if self.updates >= self.maximum_updates:
            return True
        return self.inner.converged()


# -- hooks ----------------------------------------------------------------------------------------------------------
class HookPolicy(Policy):
    def __init__(self, callback):
        self.callback = callback

    def apply(self):
        return self.callback(self.system)


__all__ = [
    "Policy", "Proxy", "ApplyProxy", "ConvergedProxy", "ResetProxy", "UpdateProxy",
    "BandwidthIncreasePolicy", "AllDirectionsIncrementBandwidthIncreasePolicy",
    "AlternatingDirectionsIncrementBandwidthIncreasePolicy", "OneDirectionIncrementBandwidthIncreasePolicy",
    "CompressionPolicy", "ConstantStateCompressionPolicy", "ConstantOperatorCompressionPolicy",
    "ContractionPolicy", "RepeatPatternContractionPolicy",
    "ConvergencePolicy", "PeriodicyThresholdConvergencePolicy",
    "RelativeEstimatedOneSiteExpectationDifferenceThresholdConvergencePolicy",
    "RelativeExpectationDifferenceDifferenceThresholdConvergencePolicy",
    "RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy",
    "RelativeStateDifferenceThresholdConvergencePolicy", "BoundedConvergencePolicy",
    "HookPolicy",
]
