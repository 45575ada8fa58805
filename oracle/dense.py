"""Dense contraction recipes of the 2D boundary environment, restated as einsum strings.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.

Leg conventions (reference ``sketches/tensors/dense/{corner,side}.svg``):
  corner  [a b o | d e p]           legs 0-2 = left bond (state, state*, operator), 3-5 = right bond
  side    [a b o | d e p | g h]     legs 0-2 left, 3-5 right, 6 = state leg facing the center, 7 = its conjugate
  center  [n0 n1 n2 n3 s]           bond towards direction 0..3 (right, up, left, down) and the physical index
All arrays are complex128, C-contiguous.
"""
import numpy as np

_c128 = np.complex128


def L(i):  # reference utils.py:886-888
    return (i + 1) % 4


def R(i):
    return (i - 1) % 4


def O(i):
    return (i + 2) % 4


def absorb_side_into_corner_from_left(corner, side):
    """reference tensors/_2d/dense.py:11-15: sum corner(0,1,2)=side(3,4,5) -> [s0][s1][s2][c3 s6][c4 s7][c5]."""
    g, h, i = side.shape[:3]
    d, e, f = corner.shape[3:]
    j, k = side.shape[6:]
    out = np.einsum("abcdef,ghiabcjk->ghidjekf", corner, side)
    return np.ascontiguousarray(out).reshape(g, h, i, d * j, e * k, f)


def absorb_side_into_corner_from_right(corner, side):
    """reference tensors/_2d/dense.py:17-21: sum corner(3,4,5)=side(0,1,2) -> [c0 s6][c1 s7][c2][s3][s4][s5]."""
    a, b, c = corner.shape[:3]
    g, h, i, j, k = side.shape[3:]
    out = np.einsum("abcdef,defghijk->ajbkcghi", corner, side)
    return np.ascontiguousarray(out).reshape(a * j, b * k, c, g, h, i)


def _center_letters(direction):
    """Letters for the center (n*) and its conjugate (m*) with leg `direction` tied to the side's legs 6/7."""
    v = ["q", "r", "t", "u"]
    w = ["v", "w", "x", "y"]
    v[direction] = "g"
    w[direction] = "h"
    return v, w


def _absorb_center_into_side(direction, side, center, center_conj, operator):
    v, w = _center_letters(direction)
    l, r, o = L(direction), R(direction), O(direction)
    out_letters = "a" + v[l] + "b" + w[l] + "c" + "d" + v[r] + "e" + w[r] + "f" + v[o] + w[o]
    if operator is None:
        expr = "abcdefgh,{}s,{}s->{}".format("".join(v), "".join(w), out_letters)
        out = np.einsum(expr, side, center, center_conj, optimize=True)
    else:
        # reference dense.py:52-57: Join(3,0,2,4) ties O's leg 0 to the conjugate's physical leg,
        # Join(3,1,1,4) ties O's leg 1 to the state's physical leg.
        expr = "abcdefgh,{}s,{}z,zs->{}".format("".join(v), "".join(w), out_letters)
        out = np.einsum(expr, side, center, center_conj, operator, optimize=True)
    sh = dict(zip("abcdef", side.shape[:6]))
    for letters, arr in ((v, center), (w, center_conj)):
        for ax, letter in enumerate(letters):
            sh[letter] = arr.shape[ax]
    new_shape = (
        sh["a"] * sh[v[l]], sh["b"] * sh[w[l]], sh["c"],
        sh["d"] * sh[v[r]], sh["e"] * sh[w[r]], sh["f"],
        sh[v[o]], sh[w[o]],
    )
    return np.ascontiguousarray(out).reshape(new_shape)


def absorb_center_ss_into_side(direction, side, center, center_conj=None):
    """reference tensors/_2d/dense.py:23-49 (absorbDenseCenterSSIntoSide)."""
    if center_conj is None:
        center_conj = center.conj()
    return _absorb_center_into_side(direction, side, center, center_conj, None)


def absorb_center_sos_into_side(direction, side, center, operator, center_conj=None):
    """reference tensors/_2d/dense.py:51-81 (absorbDenseCenterSOSIntoSide)."""
    if center_conj is None:
        center_conj = center.conj()
    return _absorb_center_into_side(direction, side, center, center_conj, operator)


def stage1(corner, side):
    """reference tensors/_2d/dense.py:96-99: sum corner(3,4,5)=side(0,1,2) -> [(c0 c1 c2)][(s3 s4 s5)][s6][s7]."""
    a, b, c = corner.shape[:3]
    g, h, i, j, k = side.shape[3:]
    out = np.einsum("abcdef,defghijk->abcghijk", corner, side)
    return np.ascontiguousarray(out).reshape(a * b * c, g * h * i, j, k)


def stage2(s1a, s1b):
    """reference tensors/_2d/dense.py:102-112: sum A0=B1 -> [B0][A1][A2][B2][A3][B3]."""
    return np.ascontiguousarray(np.einsum("kaij,bklm->bailjm", s1a, s1b))


def stage3_prejoin(s2_0, s2_1):
    """reference dense.py:130-131 / 162-163: A=[(x y),D0*,D1*,D0,D1], B=[(y' x'),D2*,D3*,D2,D3]."""
    x, y, a, b, c, d = s2_0.shape
    A = np.ascontiguousarray(s2_0.transpose(0, 1, 4, 5, 2, 3)).reshape(x * y, c, d, a, b)
    xp, yp, e, f, g, h = s2_1.shape
    B = np.ascontiguousarray(s2_1.transpose(1, 0, 4, 5, 2, 3)).reshape(yp * xp, g, h, e, f)
    return A, B


def stage3_multiply_joined(A, B, v, operator=None):
    """The center-site matvec on pre-joined halves.

    reference dense.py:115-128 (normalization: t=A.(3,4)*v.(0,1); out=B.(3,4,0)*t.(3,4,0); join [2][3][0][1][4])
    and dense.py:146-160 (operator variant first applies O[s',s] to v's physical leg).
    Two pairwise steps (never a 3-operand einsum) so the cost matches the reference's 2*X*D^6*d cmac.
    """
    if operator is not None:
        v = np.einsum("ts,cdghs->cdght", operator, v)
    X = A.shape[0]
    P = A.shape[1] * A.shape[2]
    Q = A.shape[3] * A.shape[4]
    Rr = B.shape[1] * B.shape[2]
    S = B.shape[3] * B.shape[4]
    d = v.shape[4]
    t = A.reshape(X * P, Q) @ v.reshape(Q, S * d)                    # [X P, S d]
    t = t.reshape(X, P, S, d)
    out = np.tensordot(B.reshape(X, Rr, S), t, axes=([0, 2], [0, 2]))  # [R, P, d]
    out = np.ascontiguousarray(out.transpose(1, 0, 2))
    return out.reshape(A.shape[1], A.shape[2], B.shape[1], B.shape[2], d)


def stage3_multiply(s2_0, s2_1, v, operator=None):
    A, B = stage3_prejoin(s2_0, s2_1)
    return stage3_multiply_joined(A, B, v, operator)


def stage3_cost_of_multiply(s2_0_shape, s2_1_shape, d, with_operator):
    """cmac count the reference's CostTracker assigns to one matvec (data/cost_tracker.py:17-21 run on
    the generated contractors of dense.py:115-128 / 146-160)."""
    x, y, a, b, c, dd = s2_0_shape
    xp, yp, e, f, g, h = s2_1_shape
    X = x * y
    cost = 0
    if with_operator:
        cost += d * d * a * b * e * f                     # O.(1)*v.(4)
    cost += X * c * dd * (e * f * d) * (a * b)            # A.(3,4)*v.(0,1)
    cost += (g * h) * (c * dd * d) * (e * f * X)          # B.(3,4,0)*t.(3,4,0)
    return int(cost)


def stage3_form_matrix(s2_0, s2_1, operator):
    """reference dense.py:176-194: sum s2_0(0,1)=s2_1(1,0), outer with O -> [(D0* D1* D2* D3* s')][(D0 D1 D2 D3 s)]."""
    out = np.einsum("xyabcd,yxefgh,ts->cdghtabefs", s2_0, s2_1, operator, optimize=True)
    a, b, c, d = s2_0.shape[2:]
    e, f, g, h = s2_1.shape[2:]
    t, s = operator.shape
    return np.ascontiguousarray(out).reshape(c * d * g * h * t, a * b * e * f * s)


def stage3_cost_of_form_matrix(s2_0_shape, s2_1_shape, d):
    x, y, a, b, c, dd = s2_0_shape
    xp, yp, e, f, g, h = s2_1_shape
    m = (a * b * c * dd) * (e * f * g * h)
    return int(m * x * y + m * d * d)


def normalization_submatrix(s2_0, s2_1):
    """reference dense.py:205-225."""
    out = np.einsum("xyabcd,yxefgh->cdghabef", s2_0, s2_1, optimize=True)
    a, b, c, d = s2_0.shape[2:]
    e, f, g, h = s2_1.shape[2:]
    return np.ascontiguousarray(out).reshape(c * d * g * h, a * b * e * f)


def normalization_halves(corners, sides):
    """reference dense.py:82-94 (formNormalizationMultiplier) without the final stage."""
    s2_0 = stage2(stage1(corners[0], sides[0]), stage1(corners[1], sides[1]))
    s2_1 = stage2(stage1(corners[2], sides[2]), stage1(corners[3], sides[3]))
    return s2_0, s2_1
