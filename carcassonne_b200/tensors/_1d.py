"""1D (matrix-product state / operator) contraction recipes on device tensors -- the counterpart of the reference's
``carcassonne/tensors/_1d.py``, used as an independent in-process cross-check of 1D-in-2D runs (SURVEY.md section
8f item 4).

Leg conventions (read off the reference's Join specs): environments ``L`` / ``R`` are [operator, state, state*];
the MPO center ``O`` is [right operator bond, left operator bond, physical (conjugate side), physical (state side)];
the state center ``S`` is [right bond, left bond, physical].
"""
from math import prod

from ..utils import Multiplier


def absorbCenterOSSIntoLeftEnvironment(L, O, S, S_conj):
    """reference tensors/_1d.py:12-24: new L[o', a, b] = sum L[o,s,t] O[o',o,q,p] S[a,s,p] S*[b,t,q]."""
    t = L.contractWith(S, (1,), (1,))                    # [o, t, a, p]
    t = t.contractWith(O, (0, 3), (1, 3))                # [t, a, o', q]
    t = t.contractWith(S_conj, (0, 3), (1, 2))           # [a, o', b]
    return t.join(1, 0, 2)


def absorbCenterOSSIntoRightEnvironment(R, O, S, S_conj):
    """reference tensors/_1d.py:26-38: new R[o', a, b] = sum R[o,s,t] O[o,o',q,p] S[s,a,p] S*[t,b,q]."""
    t = R.contractWith(S, (1,), (0,))                    # [o, t, a, p]
    t = t.contractWith(O, (0, 3), (0, 3))                # [t, a, o', q]
    t = t.contractWith(S_conj, (0, 3), (0, 2))           # [a, o', b]
    return t.join(1, 0, 2)


def absorbCenterSSIntoLeftEnvironment(L, S, S_conj):
    """reference tensors/_1d.py:40-50: new L[a, b] = sum L[s,t] S[a,s,p] S*[b,t,p]."""
    return L.contractWith(S, (0,), (1,)).contractWith(S_conj, (0, 2), (1, 2))


def absorbCenterSSIntoRightEnvironment(R, S, S_conj):
    """reference tensors/_1d.py:52-62: new R[a, b] = sum R[s,t] S[s,a,p] S*[t,b,p]."""
    return R.contractWith(S, (0,), (0,)).contractWith(S_conj, (0, 2), (0, 2))


def multiplyExpectation(R, L, O, S):
    """out[a, b, q] = sum R[o,s,a] L[o',t,b] O[o,o',q,p] S[s,t,p]   (reference tensors/_1d.py:74-88)."""
    t = R.contractWith(S, (1,), (0,))                    # [o, a, t, p]
    t = t.contractWith(O, (0, 3), (0, 3))                # [a, t, o', q]
    t = t.contractWith(L, (1, 2), (1, 0))                # [a, q, b]
    return t.join(0, 2, 1)


def formExpectationMatrix(R, L, O):
    """M[(a b q), (s t p)] = sum R[o,s,a] L[o',t,b] O[o,o',q,p]   (reference tensors/_1d.py:64-72)."""
    t = R.contractWith(O, (0,), (0,))                    # [s, a, o', q, p]
    t = t.contractWith(L, (2,), (0,))                    # [s, a, q, p, t, b]
    return t.join((1, 5, 2), (0, 4, 3))


def formExpectationMultiplier(R, L, O):
    """reference tensors/_1d.py:64-95; costs are the pairwise cmac counts of the contractions above."""
    o, s, a = R.shape
    o2, t, b = L.shape
    d = O.shape[2]
    n = a * b * d
    cost_multiply = o * s * a * t * d + a * t * o * d * o2 * d + a * d * t * o2 * b
    cost_matrix = s * a * o * o2 * d * d + s * a * d * d * o2 * t * b
    return Multiplier((n, n), lambda S: multiplyExpectation(R, L, O, S), cost_multiply,
                      lambda: formExpectationMatrix(R, L, O), cost_matrix)
