"""Host helpers of carcassonne_b200.utils with the reference's semantics (reference utils.py:180-207, 363-374,
886-892): direction arithmetic, bandwidth bookkeeping, the Multiplier cost model."""
import pytest

from carcassonne_b200 import utils


def test_direction_helpers():
    # directions 0..3 run counter-clockwise: opposite, left neighbour, right neighbour; A() renumbers a center axis
    # after axis d has been removed
    assert [utils.O(i) for i in range(4)] == [2, 3, 0, 1]
    assert [utils.L(i) for i in range(4)] == [1, 2, 3, 0]
    assert [utils.R(i) for i in range(4)] == [3, 0, 1, 2]
    for i in range(4):
        assert utils.L(utils.R(i)) == i and utils.O(utils.O(i)) == i
        others = [a for a in range(4) if a != i]
        assert sorted([utils.OA(i), utils.LA(i), utils.RA(i)]) == [0, 1, 2]
        assert utils.OA(i) == others.index(utils.O(i))
        assert utils.LA(i) == others.index(utils.L(i))
        assert utils.RA(i) == others.index(utils.R(i))


def test_compute_new_dimension():
    assert utils.computeNewDimension(3, by=2) == 5
    assert utils.computeNewDimension(3, to=7) == 7
    assert utils.computeNewDimension(3, to=3) == 3
    with pytest.raises(ValueError):
        utils.computeNewDimension(3)
    with pytest.raises(ValueError):
        utils.computeNewDimension(3, by=1, to=4)
    with pytest.raises(AssertionError):
        utils.computeNewDimension(3, to=2)


def test_drop_at_keeps_the_container_type():
    assert utils.dropAt((1, 2, 3), 1) == (1, 3)
    assert utils.dropAt([1, 2, 3], 0) == [2, 3]


def test_multiplier_cost_model():
    calls = []
    m = utils.Multiplier((10, 10), lambda v: calls.append(v) or "out", cost_of_multiply=1000,
                         formMatrix=lambda: "matrix", cost_of_formMatrix=5000)
    assert m("v") == "out" and calls == ["v"]
    # forming the matrix pays off once n * cost_of_multiply exceeds cost_of_formMatrix + n * rows * cols
    assert not m.isCheaperToFormMatrix(5)          # 5000 > 5000 + 500 is false
    assert m.isCheaperToFormMatrix(6)              # 6000 > 5000 + 600
    dense = utils.Multiplier.fromMatrix(type("M", (), {"shape": (4, 6), "matvecWith": lambda self, v: ("mv", v)})())
    assert dense.shape == (4, 6) and dense.cost_of_multiply == 24 and dense.cost_of_formMatrix == 0
    assert dense("x") == ("mv", "x")
    assert not dense.isCheaperToFormMatrix(100)    # already a matrix: nothing to gain


def test_error_types_mirror_the_reference():
    assert issubclass(utils.DimensionMismatchError, ValueError)
    assert issubclass(utils.UnexpectedTensorRankError, ValueError)
    failure = utils.RelaxFailed(1.0, 2.0)
    assert (failure.initial_value, failure.final_value) == (1.0, 2.0)
