"""Host logic of the system-level multi-GPU mode ("stage-2 construction shards the same way:
each GPU builds its X-slab, slice x of the stage-1 tensors"), without a device.

``stage2SlabPlan`` (carcassonne_b200/tensors/_2d/dense.py) is the launch plan of the batched scatter-epilogue GEMM
that builds one rank's slab of a stage-2 half; here a NumPy emulation of that GEMM executes the plan and the result is
compared with the oracle's stage 2 + pre-join (reference dense.py:102-112, 130-131) sliced along X.  Also: the slabs of
the two halves pair up index by index, so the sum over ranks of the slab matvecs is the full matvec.
"""
import itertools
import os
import sys
from math import prod

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from carcassonne_b200.tensors._2d.dense import stage2SlabPlan  # noqa: E402
from carcassonne_b200.distributed import balance_terms  # noqa: E402
from oracle import dense  # noqa: E402


def _table(levels):
    """rowoff / coloff table of carc_index_table: digits row-major over the level extents."""
    extents = [e for e, _ in levels]
    out = np.zeros(prod(extents), dtype=np.int64)
    for i, digits in enumerate(itertools.product(*[range(e) for e in extents])):
        out[i] = sum(d * s for d, (_, s) in zip(digits, levels))
    return out


def run_plan(plan, s1a, s1b):
    """What carc_zgemm_tab does with the plan: C_b[rowoff(m) + coloff(n)] = op_T(A) . B_b."""
    A, B = s1a.ravel(), s1b.ravel()
    out = np.full(prod(plan["out_shape"]), np.nan + 0j)
    rowoff, coloff = _table(plan["rows"]), _table(plan["cols"])
    K, M, N = plan["K"], plan["M"], plan["N"]
    k = np.arange(K)
    for batch in range(plan["batch"]):
        At = A[plan["a_offset"] + k[:, None] * plan["lda"] + np.arange(M)[None, :]]          # [K, M]
        Bb = B[plan["b_offset"] + batch * plan["strideB"] + k[:, None] * plan["ldb"] + np.arange(N)[None, :]]
        out[batch * plan["strideC"] + rowoff[:, None] + coloff[None, :]] = At.T @ Bb
    return out.reshape(plan["out_shape"])


def _crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


SHAPES = [  # (K, a1, a2, a3), (b0, K, b2, b3)
    ((4, 6, 2, 3), (5, 4, 3, 2)),
    ((1, 1, 1, 1), (1, 1, 1, 1)),
    ((9, 4, 3, 3), (4, 9, 3, 3)),
    ((2, 7, 2, 2), (3, 2, 1, 4)),
]


@pytest.mark.parametrize("a_shape,b_shape", SHAPES)
def test_unsharded_plan_matches_oracle(a_shape, b_shape):
    rng = np.random.default_rng(1)
    s1a, s1b = _crand(rng, *a_shape), _crand(rng, *b_shape)
    s2 = dense.stage2(s1a, s1b)
    np.testing.assert_allclose(run_plan(stage2SlabPlan(a_shape, b_shape, None), s1a, s1b), s2, rtol=1e-13, atol=1e-13)
    A, _ = dense.stage3_prejoin(s2, s2.transpose(1, 0, 2, 3, 4, 5))
    np.testing.assert_allclose(run_plan(stage2SlabPlan(a_shape, b_shape, 0), s1a, s1b), A, rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("a_shape,b_shape", SHAPES)
def test_slabs_tile_the_joined_bond(a_shape, b_shape, world):
    rng = np.random.default_rng(2)
    s1a, s1b = _crand(rng, *a_shape), _crand(rng, *b_shape)
    s2 = dense.stage2(s1a, s1b)                                      # [B0, A1, A2, B2, A3, B3]
    b0, a1 = s2.shape[:2]
    full = {0: s2.transpose(0, 1, 4, 5, 2, 3).reshape(b0 * a1, *[s2.shape[i] for i in (4, 5, 2, 3)]),
            1: s2.transpose(1, 0, 4, 5, 2, 3).reshape(b0 * a1, *[s2.shape[i] for i in (4, 5, 2, 3)])}
    for half in (0, 1):
        covered = 0
        for rank in range(world):
            plan = stage2SlabPlan(a_shape, b_shape, half, (rank, world))
            lo, hi = plan["x_range"]
            assert lo == covered and plan["full_X"] == b0 * a1
            assert plan["out_shape"][0] == hi - lo
            covered = hi
            if hi > lo:
                np.testing.assert_allclose(run_plan(plan, s1a, s1b), full[half][lo:hi], rtol=1e-13, atol=1e-13)
        assert covered == b0 * a1


@pytest.mark.parametrize("world", [2, 3, 4])
def test_slab_matvecs_sum_to_the_full_matvec(world):
    """Half 0 is sliced on B0, half 1 on A1: the same range of the SAME ring bond, so slab r of A pairs with slab r of B."""
    rng = np.random.default_rng(3)
    chi2, D, d = 5, 2, 2
    # half 0 from stage-1 tensors (a, b), half 1 from (c, e); the ring closes: B0 of half 0 <-> A1 of half 1 (extent x),
    # A1 of half 0 <-> B0 of half 1 (extent y)
    x, y, K0, K1 = 5, 3, 4, 6
    a, b = _crand(rng, K0, y, D, D), _crand(rng, x, K0, D, D)
    c, e = _crand(rng, K1, x, D, D), _crand(rng, y, K1, D, D)
    v = _crand(rng, D, D, D, D, d)
    A, B = dense.stage3_prejoin(dense.stage2(a, b), dense.stage2(c, e))
    want = dense.stage3_multiply_joined(A, B, v)
    got = np.zeros_like(want)
    for rank in range(world):
        pa = stage2SlabPlan(a.shape, b.shape, 0, (rank, world))
        pb = stage2SlabPlan(c.shape, e.shape, 1, (rank, world))
        assert pa["x_range"] == pb["x_range"]
        if pa["out_shape"][0]:
            got += dense.stage3_multiply_joined(run_plan(pa, a, b), run_plan(pb, c, e), v)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    del chi2


def test_slabs_need_a_stage3_layout():
    with pytest.raises(ValueError):
        stage2SlabPlan((2, 2, 2, 2), (2, 2, 2, 2), None, (0, 2))


def test_balance_terms_is_a_partition_and_balanced():
    costs = [7, 7, 7, 1, 1, 1, 1, 5, 5, 5, 5, 5, 5, 3, 3, 3, 3, 3, 3, 3]
    for world in (1, 2, 4, 8):
        owner = balance_terms(costs, world)
        assert sorted(t for o in owner for t in o) == list(range(len(costs)))
        loads = [sum(costs[t] for t in o) for o in owner]
        assert max(loads) - min(loads) <= max(costs)
