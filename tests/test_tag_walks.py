"""The product's sparse-tag planner (carcassonne_b200/sparse.py: pairing rules, planSparseContraction,
stage3_term_allowed -- pure host logic) against the oracle's restatement of the reference's rules
(oracle/tags.py, pinned to the reference's own walks by tests/test_oracle_golden.py): random absorption walks of
the transverse-Ising and Heisenberg operators, compared tag by tag and term by term after every absorption.  The
product side moves tags only; the oracle side runs its CPU system on trivial (all bonds 1) tensors."""
import itertools

import numpy as np
import pytest

from carcassonne_b200 import sparse as sp
from carcassonne_b200.utils import L, R
from oracle import tags as otags
from oracle.system import System as OracleSystem

X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)


class Site:
    """Stand-in for a d x d device matrix: makeSparseOperator only needs its shape and an identity."""

    def __init__(self, array):
        self.array = np.asarray(array, dtype=complex)
        self.shape = self.array.shape

    def newIdentity(self, n):
        return Site(np.eye(n))


def results(rule, tags_1, tags_2):
    """Result tags of one sparse contraction in first-appearance order (the key order of the reference's result dict)."""
    return list(dict.fromkeys(res for _, _, res, _ in sp.planSparseContraction(rule, tags_1, tags_2)))


class TagSystem:
    """system/_2d.py:443-459 on tags alone."""

    def __init__(self, operator_tags):
        self.corners = [[sp.Identity()] for _ in range(4)]
        self.sides = [[sp.Identity()] for _ in range(4)]
        self.operator = list(operator_tags)

    def contract_towards(self, d):
        self.corners[d] = results(sp.rule_side_into_corner_from_left, self.corners[d], self.sides[L(d)])
        self.sides[d] = results(lambda s, c: sp.rule_center_into_side(d, s, c), self.sides[d], self.operator)
        self.corners[R(d)] = results(sp.rule_side_into_corner_from_right, self.corners[R(d)], self.sides[R(d)])

    def terms(self):
        stage1 = [results(sp.rule_stage1, self.corners[i], self.sides[i]) for i in range(4)]
        half_0 = results(sp.rule_stage2, stage1[0], stage1[1])
        half_1 = results(sp.rule_stage2, stage1[2], stage1[3])
        return [t for t in itertools.product(half_0, half_1, self.operator) if sp.stage3_term_allowed(*t)]


def as_tuples(tags):
    return [otags.from_reference_tag(t) for t in tags]


MODELS = {
    "tfim": dict(Os=[-Z], OO_UDs=[(X, -0.7 * X)], OO_LRs=[(X, -0.7 * X)]),
    "heisenberg": dict(Os=[], OO_UDs=[(X, X), (Y, Y), (Z, Z)], OO_LRs=[(X, X), (Y, Y), (Z, Z)]),
    "chain": dict(Os=[-Z], OO_UDs=[], OO_LRs=[(X, -0.3 * X)]),
}


@pytest.mark.parametrize("model", sorted(MODELS))
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_walk_matches_the_oracle(model, seed):
    spec = MODELS[model]
    product_operator = sp.makeSparseOperator([Site(o) for o in spec["Os"]],
                                             [(Site(a), Site(b)) for a, b in spec["OO_UDs"]],
                                             [(Site(a), Site(b)) for a, b in spec["OO_LRs"]])
    oracle = OracleSystem.new_trivial(otags.make_sparse_operator(spec["Os"], spec["OO_UDs"], spec["OO_LRs"]))
    product = TagSystem(product_operator)
    assert as_tuples(product.operator) == list(oracle.operator_center)
    rng = np.random.default_rng(seed)
    walk = list(range(4)) + [int(d) for d in rng.integers(0, 4, 8)]
    counts = []
    for step, d in enumerate(walk):
        product.contract_towards(d)
        oracle.contract_towards(d)
        for i in range(4):
            assert as_tuples(product.corners[i]) == list(oracle.corners[i]), (step, "corner", i)
            assert as_tuples(product.sides[i]) == list(oracle.sides[i]), (step, "side", i)
        H, _ = oracle.multipliers()
        got = [tuple(as_tuples(t)) for t in product.terms()]
        assert got == [tuple(t) for t in H.terms], (step, d)
        counts.append(len(got))
    # 9 terms (TFIM) / 20 (Heisenberg) after the first round of four directions
    if model == "tfim":
        assert counts[3] == 9
    if model == "heisenberg":
        assert counts[3] == 20
    assert counts[-1] >= counts[3]
