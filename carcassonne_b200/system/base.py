"""Run loop shared by the lattice systems: named policy slots, sweep / run drivers and the bandwidth increase.

Host-side orchestration with the interface of the reference's ``carcassonne/system/base.py``; every tensor
operation it triggers (normalizeAxis, absorbMatrixAt, absorption, the eigen-solve) runs on the device.
"""
import logging
from copy import copy

from ..utils import O, RelaxFailed, computeNewDimension

log = logging.getLogger(__name__)

POLICY_SLOTS = (
    "bandwidth increase",
    "contraction",
    "operator compression",
    "run convergence",
    "post-contraction hook",
    "pre-optimization hook",
    "post-optimization hook",
    "state compression",
    "sweep convergence",
)


class BaseSystem:
    def __init__(self):
        self._policies = dict.fromkeys(POLICY_SLOTS)
        self.number_of_sweeps = 0
        self.number_of_iterations = 0
        self.iteration_number_for_sweep = 0

    # -- policies ---------------------------------------------------------------------------------------------------
    def setPolicy(self, policy_name, policy):
        if policy_name not in self._policies:
            raise ValueError("No such policy name " + policy_name)
        if self._policies[policy_name] is not None:
            raise ValueError("Policy " + policy_name + " has already been set.")
        self._policies[policy_name] = policy.createBindingToSystem(self)

    def _policy(self, policy_name, optional=False):
        if policy_name not in self._policies:
            raise ValueError("No such policy name " + policy_name)
        policy = self._policies[policy_name]
        if policy is None and not optional:
            raise ValueError("Policy " + policy_name + " has not been set.")
        return policy

    def _call(self, policy_name, method, optional=False):
        policy = self._policy(policy_name, optional)
        return getattr(policy, method)() if policy is not None else None

    def _applyPolicy(self, policy_name, optional=False):
        return self._call(policy_name, "apply", optional)

    def _resetPolicy(self, policy_name, optional=False):
        return self._call(policy_name, "reset", optional)

    def _updatePolicy(self, policy_name, optional=False):
        return self._call(policy_name, "update", optional)

    def _hasConverged(self, policy_name):
        return self._call(policy_name, "converged")

    # -- drivers (reference base.py:58-111) ---------------------------------------------------------------------------
    def computeEstimatedOneSiteExpectation(self, direction=0):
        system = copy(self)
        before = system.computeExpectation()
        system.contractTowards(direction)
        return system.computeExpectation() - before

    def _optimize(self):
        self._applyPolicy("pre-optimization hook", optional=True)
        self.minimizeExpectation()
        self._applyPolicy("post-optimization hook", optional=True)
        self._updatePolicy("sweep convergence")

    def sweepUntilConverged(self):
        self.number_of_sweeps += 1
        sweep = self.number_of_sweeps
        log.info("Starting sweep #%d", sweep)
        self._resetPolicy("contraction")
        self._resetPolicy("sweep convergence")
        self.iteration_number_for_sweep = 1
        self._optimize()
        while not self._hasConverged("sweep convergence"):
            self._applyPolicy("contraction")
            self._applyPolicy("post-contraction hook", optional=True)
            self._applyPolicy("state compression", optional=True)
            self._applyPolicy("operator compression", optional=True)
            self.iteration_number_for_sweep += 1
            self.number_of_iterations += 1
            log.info("Iteration #%d of sweep #%d", self.iteration_number_for_sweep, sweep)
            try:
                self._optimize()
            except RelaxFailed:
                pass  # the reference swallows a failed relaxation and keeps sweeping (base.py:104-110)

    def runUntilConverged(self):
        log.info("Beginning run.")
        self.number_of_sweeps = 0
        self.number_of_iterations = 0
        self.sweepUntilConverged()
        self._updatePolicy("run convergence")
        while not self._hasConverged("run convergence"):
            self._applyPolicy("bandwidth increase")
            self.sweepUntilConverged()
            self._updatePolicy("run convergence")
        log.info("Finished run with %d total sweeps and %d total iterations.", self.number_of_sweeps,
                 self.number_of_iterations)

    # -- bandwidth increase (reference base.py:114-160) ---------------------------------------------------------------
    def _increaseBandwidth(self, axis, by=None, to=None, do_as_much_as_possible=False, enlargeners=None):
        center = self.state_center_data
        opposite = O(axis) if center.ndim == 5 else 1 - axis
        physical_dimension = center.shape[-1]
        old_dimension = center.shape[axis]
        new_dimension = computeNewDimension(old_dimension, by=by, to=to)
        if new_dimension == old_dimension:
            return None
        limit = physical_dimension * old_dimension
        if new_dimension > limit:
            if not do_as_much_as_possible:
                raise ValueError("New dimension must be less than the physical dimension times the old dimension "
                                 "({} > {}*{}).".format(new_dimension, physical_dimension, old_dimension))
            new_dimension = limit
        # fail BEFORE anything is mutated: the device SVD behind normalizeAxis handles bonds up to linalg.MAX_SMALL
        from ..linalg import MAX_SMALL
        if new_dimension > MAX_SMALL:
            raise NotImplementedError(
                "bond dimension {} exceeds the device SVD limit of {} columns (carc_svd_small); the system has not been "
                "changed".format(new_dimension, MAX_SMALL))

        towards_opposite = center.normalizeAxis(opposite)[0]
        towards_axis = center.normalizeAxis(axis)[0]
        if enlargeners is None:
            enlargeners = center.newEnlargener(old_dimension, new_dimension)
        grow, grow_conj = enlargeners

        center = center.absorbMatrixAt(axis, grow)
        towards_opposite = towards_opposite.absorbMatrixAt(opposite, grow_conj)
        towards_opposite, center = towards_opposite.normalizeAxisAndDenormalize(opposite, axis, center)

        towards_axis = towards_axis.absorbMatrixAt(axis, grow)
        center = center.absorbMatrixAt(opposite, grow_conj)
        towards_axis, center = towards_axis.normalizeAxisAndDenormalize(axis, opposite, center)

        self.setStateCenter(center)
        self.contractUnnormalizedTowards(axis, towards_opposite)
        self.contractUnnormalizedTowards(opposite, towards_axis)
        self.just_increased_bandwidth = True
        return grow, grow_conj
