"""Star decomposition of the sparse term list (csrc/stage3.cu: stage3_plan_host) on the CPU through
carc_stage3_describe_stars.  The terms of formExpectationStage3 (reference tensors/_2d/sparse.py:100-127) form a
bipartite multigraph between half-0 and half-1 stage-2 tensors; a star shares its centre's products."""
import numpy as np

from carcassonne_b200 import _lib


def stars(pairs, X=None):
    n = len(pairs)
    a = np.array([p[0] for p in pairs], dtype=np.int32)
    b = np.array([p[1] for p in pairs], dtype=np.int32)
    x = np.array(X if X is not None else [7] * n, dtype=np.int64)
    ng = np.zeros(1, dtype=np.int32)
    kind, first, count, order = (np.full(max(n, 1), -1, dtype=np.int32) for _ in range(4))
    rc = _lib.lib.carc_stage3_describe_stars(n, a.ctypes.data, b.ctypes.data, x.ctypes.data, ng.ctypes.data,
                                             kind.ctypes.data, first.ctypes.data, count.ctypes.data, order.ctypes.data)
    assert rc == 0
    g = int(ng[0])
    return [(int(kind[i]), [int(t) for t in order[first[i]:first[i] + count[i]]]) for i in range(g)]


def products(groups):
    first = sum(1 if kind else len(ts) for kind, ts in groups)
    second = sum(len(ts) if kind else 1 for kind, ts in groups)
    return first, second


def check_partition(pairs, groups, X=None):
    seen = sorted(t for _, ts in groups for t in ts)
    assert seen == list(range(len(pairs)))
    for kind, ts in groups:
        centres = {pairs[t][0 if kind else 1] for t in ts}
        assert len(centres) == 1                      # a star: one shared tensor
        if X is not None:
            assert len({X[t] for t in ts}) == 1       # and one environment extent


def test_transverse_ising_term_list():
    # the 9 terms bench.py reads off the planner after one absorption round (6 + 6 tensors, Identity = tensor 0)
    pairs = [(1, 0), (0, 1), (0, 0), (4, 0), (5, 0), (0, 4), (0, 5), (3, 2), (2, 3)]
    groups = stars(pairs)
    check_partition(pairs, groups)
    assert [(k, sorted(ts)) for k, ts in groups] == [(0, [0, 2, 3, 4]), (1, [1, 5, 6]), (0, [7]), (0, [8])]
    assert products(groups) == (7, 6)                 # instead of the reference's 9 + 9


def test_heisenberg_like_term_list():
    # 7 terms with Identity (tensor 0) on half 1, 8 with Identity on half 0 (one of them both), 6 cross terms with their
    # own pair of tensors: the larger star (half-0 Identity, 8 terms) is peeled first, then the 6 left of the other
    pairs = [(i, 0) for i in range(7)] + [(0, i) for i in range(1, 7)] + [(10 + i, 10 + i) for i in range(6)] + [(0, 7)]
    groups = stars(pairs)
    check_partition(pairs, groups)
    assert groups[0][0] == 1 and len(groups[0][1]) == 8 and groups[1][0] == 0 and len(groups[1][1]) == 6
    assert len(pairs) == 20 and products(groups) == (13, 15)      # 28 products per x instead of 40


def test_random_term_lists():
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(0, 25))
        na, nb = int(rng.integers(1, 8)), int(rng.integers(1, 8))
        pairs = [(int(rng.integers(na)), int(rng.integers(nb))) for _ in range(n)]
        X = [int(rng.choice([5, 9])) for _ in range(n)]
        groups = stars(pairs, X)
        check_partition(pairs, groups, X)
        first, second = products(groups)
        assert first <= n and second <= n and first + second == n + len(groups)
        # greedy: group sizes never increase, and the first star is a largest one
        sizes = [len(ts) for _, ts in groups]
        assert sizes == sorted(sizes, reverse=True)
        if n:
            best = 0
            for t in range(n):
                best = max(best, sum(1 for u in range(n) if pairs[u][1] == pairs[t][1] and X[u] == X[t]),
                           sum(1 for u in range(n) if pairs[u][0] == pairs[t][0] and X[u] == X[t]))
            assert sizes[0] == best


def test_duplicate_terms_and_single_term():
    assert stars([(3, 4)]) == [(0, [0])]
    groups = stars([(1, 1), (1, 1), (1, 1)])
    check_partition([(1, 1)] * 3, groups)
    assert len(groups) == 1 and products(groups) in ((3, 1), (1, 3))
    assert stars([]) == []
