"""Sparse (tagged) versions of the 2D contraction recipes and the expectation / normalization multipliers.

Mirror of the reference's ``carcassonne/tensors/_2d/sparse.py``.  The host plans which tag pairs multiply
(``carcassonne_b200.sparse``); the device executes one GEMM per pair with the ``+=`` of the result tag folded into
the epilogue.  Stage 2 writes its output directly in the layout the fused stage-3 matvec streams, and stage 3
collects all sparse terms into ONE device operator (one kernel launch per matvec over a device-side term table)
instead of the reference's Python loop over per-term multipliers (sparse.py:129-133).
"""
import itertools
from math import prod

import numpy as np

from ...sparse import (Identity, contractSparseTensors, getInformationFromOperatorCenter,
                       rule_center_into_side, rule_side_into_corner_from_left, rule_side_into_corner_from_right,
                       rule_stage1, rule_stage2, stage3_term_allowed)
from ...utils import Multiplier
from . import dense as _dense
from .dense import *  # noqa: F401,F403  (the reference re-exports the dense recipes from here too)


class _Memo:
    """Small identity-keyed cache of environment pieces.  The convergence policies rebuild the environment of an
    unchanged system several times per sweep iteration (``computeEstimatedOneSiteExpectation`` copies the system,
    evaluates <H>, absorbs once and evaluates again -- 37 % of the reference's run time);
    device tensors are immutable values shared by the shallow ``System.__copy__``, so a stage-1 / stage-2 result
    can be reused whenever exactly the same input tensors come back.  Entries keep their inputs alive (ids stay
    valid) and are bounded in number and in bytes."""

    def __init__(self, max_entries=16, max_bytes=4 << 30):
        from collections import OrderedDict
        self.entries = OrderedDict()
        self.max_entries = max_entries
        self.max_bytes = max_bytes
        self.hits = self.misses = 0

    @staticmethod
    def _key(sparse_tensors, extra):
        return tuple(tuple((tag, id(data), data.version) for tag, data in t.items()) for t in sparse_tensors) + (extra,)

    def get(self, sparse_tensors, extra, compute):
        key = self._key(sparse_tensors, extra)
        hit = self.entries.get(key)
        if hit is not None:
            self.entries.move_to_end(key)
            self.hits += 1
            return hit[1]
        self.misses += 1
        value = compute()
        # what an entry pins: its value AND its inputs (which the system may otherwise have released)
        size = sum(16 * data.size() for data in value.values()) + \
            sum(16 * data.size() for t in sparse_tensors for data in t.values())
        if size <= self.max_bytes // 4:
            self.entries[key] = ([list(t.values()) for t in sparse_tensors], value, size)
            while len(self.entries) > self.max_entries or sum(e[2] for e in self.entries.values()) > self.max_bytes:
                self.entries.popitem(last=False)
        return value

    def clear(self):
        self.entries.clear()


environment_cache = _Memo()


class Stage2Half(dict):
    """{tag: tensor} of one half-ring in the stage-3 streaming layout [(x y), D*, D*, D, D] (``half`` = 0 or 1);
    ``bonds`` keeps the two environment bond extents of the reference layout."""

    def __init__(self, half, bonds=None, slab=None):
        dict.__init__(self)
        self.half = half
        self.bonds = bonds
        self.slab = slab          # (rank, world) when only this rank's slab of X is held (multi-GPU), else None
        self.full_X = {}          # tag -> extent of the whole joined bond X (== shape[0] when not sharded)


def absorbSparseSideIntoCornerFromLeft(corner, side):
    """reference tensors/_2d/sparse.py:12-23."""
    return contractSparseTensors(rule_side_into_corner_from_left,
                                 lambda c, s, acc, _: _dense.absorbDenseSideIntoCornerFromLeft(c, s, acc), corner, side)


def absorbSparseSideIntoCornerFromRight(corner, side):
    """reference tensors/_2d/sparse.py:24-35."""
    return contractSparseTensors(rule_side_into_corner_from_right,
                                 lambda c, s, acc, _: _dense.absorbDenseSideIntoCornerFromRight(c, s, acc), corner, side)


def absorbSparseCenterSOSIntoSide(direction, side, state_center_data, operator_center_data,
                                  state_center_data_conj=None):
    """reference tensors/_2d/sparse.py:36-58.  The double-layer center tensor of each site operator is built once
    and shared by all side tags that pair with it."""
    if state_center_data_conj is None:
        state_center_data_conj = state_center_data.conj()
    cache = {}

    def absorb(side_data, operator_data, acc, needs_operator):
        return _dense.absorbDenseCenterSOSIntoSide(direction, side_data, state_center_data,
                                                   operator_data if needs_operator else None,
                                                   state_center_data_conj, acc, cache)

    return contractSparseTensors(lambda s, c: rule_center_into_side(direction, s, c), absorb, side,
                                 operator_center_data)


def formExpectationStage1(corner, side):
    """reference tensors/_2d/sparse.py:72-85 (memoised on the identity of the input tensors)."""
    return environment_cache.get(
        (corner, side), "stage1",
        lambda: contractSparseTensors(rule_stage1, lambda c, s, acc, _: _dense.formNormalizationStage1(c, s, acc),
                                      corner, side))


def formExpectationStage2(right, left, half=None):
    """reference tensors/_2d/sparse.py:86-99.  ``half`` = 0 / 1 produces a ``Stage2Half`` in the stage-3 layout; in
    the multi-GPU mode (``distributed.shard_environment``) it holds this rank's slab of X only."""
    from ... import distributed as _dist
    sharding = _dist.environment_sharding() if half is not None else None
    slab = sharding.slab if sharding is not None else None

    def compute():
        full_X = {}

        def dense(r, l, acc, _):
            result = _dense.formNormalizationStage2(r, l, acc, half=half, slab=slab)
            full_X[id(result)] = l.shape[0] * r.shape[1]      # extent of the whole joined bond (B0 A1)
            return result

        result = contractSparseTensors(rule_stage2, dense, right, left)
        if half is None:
            return result
        out = Stage2Half(half, (left[Identity()].shape[0], right[Identity()].shape[1]), slab)
        out.update(result)
        out.full_X = {tag: full_X[id(data)] for tag, data in result.items()}
        return out

    return environment_cache.get((right, left), ("stage2", half, slab), compute)


def _as_half(stage2, half):
    if isinstance(stage2, Stage2Half):
        if stage2.half != half:
            raise ValueError("stage-2 tensor was built for half {} but is used as half {}".format(stage2.half, half))
        return stage2
    out = Stage2Half(half, tuple(stage2[Identity()].shape[:2]))
    for tag, data in stage2.items():
        out[tag] = _dense.prejoinStage2(data, half)
        out.full_X[tag] = out[tag].shape[0]
    return out


def stage3Terms(stage2_0, stage2_1, operator_center):
    """Term list [(tag_0, tag_1, tag_center)] in the reference's itertools.product order (sparse.py:119-127)."""
    return [tags for tags in itertools.product(stage2_0, stage2_1, operator_center) if stage3_term_allowed(*tags)]


def formExpectationStage3(stage2_0, stage2_1, operator_center):
    """reference tensors/_2d/sparse.py:100-161 -> (expectation Multiplier, normalization Multiplier)."""
    from ...operator import Stage3Operator
    from ... import distributed as _dist
    half_0, half_1 = _as_half(stage2_0, 0), _as_half(stage2_1, 1)
    if half_0.slab != half_1.slab:
        raise ValueError("stage-2 halves were built for different X slabs: {} vs {}".format(half_0.slab, half_1.slab))
    sharding = _dist.environment_sharding() if half_0.slab is not None else None
    if half_0.slab is not None and (sharding is None or sharding.slab != half_0.slab):
        raise ValueError("stage-2 halves hold an X slab but the multi-GPU mode that built them is no longer active")
    physical_dimension, _, DataClass = getInformationFromOperatorCenter(operator_center)
    A_id, B_id = half_0[Identity()], half_1[Identity()]
    state_shape = (A_id.shape[3], A_id.shape[4], B_id.shape[3], B_id.shape[4], physical_dimension)
    dimension = prod(state_shape)
    terms = stage3Terms(half_0, half_1, operator_center)

    expectation_operator = Stage3Operator(state_shape)
    cost_of_multiply = cost_of_formMatrix = 0
    for x, y, z in terms:
        site = None if z == Identity() else operator_center[z]
        if half_0[x].shape[0] > 0:          # an empty slab (more ranks than slow bond indices) contributes nothing
            expectation_operator.add_term(half_0[x], half_1[y], site)
        # costs are those of the WHOLE operator, so that every rank (and a single GPU) takes the same solver branches
        cost_of_multiply += _dense.stage3CostOfMultiply(half_0[x], half_1[y], physical_dimension, site is not None,
                                                        half_0.full_X[x])
        cost_of_formMatrix += _dense.stage3CostOfFormMatrix(half_0[x], half_1[y], physical_dimension, half_0.full_X[x])
    expectation_operator.finalize()
    if sharding is not None:
        sharding.attach(expectation_operator)

    identity = np.eye(physical_dimension, dtype=np.complex128)

    def formExpectationMatrix():
        matrix = DataClass.newZeros((dimension, dimension))
        for x, y, z in terms:
            _dense.stage3FormMatrix(half_0[x], half_1[y], identity if z == Identity() else operator_center[z], matrix)
        return sharding.sum_matrix_(matrix) if sharding is not None else matrix

    expectation_multiplier = Multiplier((dimension, dimension), expectation_operator, cost_of_multiply,
                                        formExpectationMatrix, cost_of_formMatrix)
    expectation_multiplier.device_operator = expectation_operator
    expectation_multiplier.terms = terms
    normalization_multiplier = _dense._stage3_multiplier(A_id, B_id, None, physical_dimension, sharding,
                                                         half_0.full_X[Identity()])
    return expectation_multiplier, normalization_multiplier


def formExpectationAndNormalizationMultipliers(corners, sides, operator_center_data):
    """reference tensors/_2d/sparse.py:59-71."""
    return formExpectationStage3(
        formExpectationStage2(formExpectationStage1(corners[0], sides[0]),
                              formExpectationStage1(corners[1], sides[1]), half=0),
        formExpectationStage2(formExpectationStage1(corners[2], sides[2]),
                              formExpectationStage1(corners[3], sides[3]), half=1),
        operator_center_data,
    )
