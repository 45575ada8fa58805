"""Sparse operator tag algebra, restated with plain tuples.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.

Tags (reference sparse.py:24-208):
  I                      = ("I",)                 Identity()
  C                      = ("C",)                 Complete()
  ONE                    = ("1",)                 OneSiteOperator(id)  (the reference discards id, sparse.py:36-37)
  two(id, dir, pos)      = ("2", id, dir, pos)    TwoSiteOperator
  zipd(dir)              = ("Z", dir)             TwoSiteOperatorCompressed

In the *center* operator dict ``dir`` is the neighbour the partner sits at (0 right, 1 up, 2 left,
3 down; sparse.py:313-333).  In *environment* dicts ``dir`` is LEFT/RIGHT/CENTER (sparse.py:212-214).
A sparse tensor is a dict {tag: ndarray}; dict order is insertion order, as in the reference.
"""
from . import dense as _d

I = ("I",)
C = ("C",)
ONE = ("1",)
LEFT, RIGHT, CENTER = 0, 1, 2


def two(id_, direction, position=None):
    return ("2", id_, direction, position)


def zipd(direction):
    return ("Z", direction)


def kind(tag):
    return tag[0]


def _move_out(t):  # sparse.py:164-166
    return two(t[1], t[2], t[3] + 1)


def _two_matches(left, right):  # sparse.py:81-89
    if left[1] == right[1] and left[2] == RIGHT and right[2] == LEFT and left[3] == right[3]:
        return C
    return None


def _zip_matches(left, right):  # sparse.py:185-188
    if left[1] == RIGHT and right[1] == LEFT:
        return C
    return None


def contract_sparse(rule, dense_fn, t1, t2):
    """reference sparse.py:223-240 (contractSparseTensors): every (tag1, tag2) pair the rule accepts is
    contracted densely and accumulated under the resulting tag, in dict iteration order."""
    out = {}
    for tag1, data1 in t1.items():
        for tag2, data2 in t2.items():
            res = rule(tag1, tag2)
            if res is None:
                continue
            tag, fn = res if isinstance(res, tuple) and callable(res[-1]) else (res, dense_fn)
            if tag is None:
                continue
            val = fn(data1, data2)
            if tag in out:
                out[tag] = out[tag] + val
            else:
                out[tag] = val
    return out


def _standard(k1, k2):  # sparse.py:218-222
    if (k1, k2) == ("I", "I"):
        return I
    if (k1, k2) in (("C", "I"), ("I", "C")):
        return C
    return False


def absorb_side_into_corner_from_left(corner, side):
    """reference tensors/_2d/sparse.py:12-23; arguments of each rule are (corner tag, side tag)."""
    def rule(c, s):
        k = (kind(c), kind(s))
        std = _standard(*k)
        if std is not False:
            return std
        if k == ("2", "I"):                       # r.matchesSideIdentityOnLeft  (sparse.py:131-134)
            return _move_out(c) if c[2] == RIGHT else None
        if k == ("I", "2"):                       # l.matchesCornerIdentityOnRight (sparse.py:125-130)
            if s[2] == LEFT:
                return s
            if s[2] == CENTER:
                return two(s[1], RIGHT, 0)
            return None
        if k == ("2", "2"):                       # l.matches(r) with l = side, r = corner
            return _two_matches(s, c)
        if k == ("Z", "Z"):
            return _zip_matches(s, c)
        return None
    return contract_sparse(rule, _d.absorb_side_into_corner_from_left, corner, side)


def absorb_side_into_corner_from_right(corner, side):
    """reference tensors/_2d/sparse.py:24-35; TwoSite rules see (l=corner, r=side); the Compressed rule
    is declared ``lambda r,l: l.matches(r)`` so it sees (r=corner, l=side) -- reproduced as written."""
    def rule(c, s):
        k = (kind(c), kind(s))
        std = _standard(*k)
        if std is not False:
            return std
        if k == ("2", "I"):                       # l.matchesSideIdentityOnRight (sparse.py:135-138)
            return _move_out(c) if c[2] == LEFT else None
        if k == ("I", "2"):                       # r.matchesCornerIdentityOnLeft (sparse.py:113-118)
            if s[2] == RIGHT:
                return s
            if s[2] == CENTER:
                return two(s[1], LEFT, 0)
            return None
        if k == ("2", "2"):
            return _two_matches(c, s)
        if k == ("Z", "Z"):
            return _zip_matches(s, c)
        return None
    return contract_sparse(rule, _d.absorb_side_into_corner_from_right, corner, side)


def absorb_center_sos_into_side(direction, side, center, operator_center, center_conj=None):
    """reference tensors/_2d/sparse.py:36-58; rule arguments are (side tag, center-operator tag)."""
    if center_conj is None:
        center_conj = center.conj()

    def ss(side_data, _):
        return _d.absorb_center_ss_into_side(direction, side_data, center, center_conj)

    def sos(side_data, op):
        return _d.absorb_center_sos_into_side(direction, side_data, center, op, center_conj)

    def rule(s, c):
        k = (kind(s), kind(c))
        if k == ("I", "I"):
            return (I, ss)
        if k == ("C", "I"):
            return (C, ss)
        if k == ("I", "1"):
            return (C, sos)
        if k == ("2", "I"):                       # s.matchesCenterIdentity (sparse.py:98-101)
            return (_move_out(s) if s[2] != CENTER else None, ss)
        if k == ("I", "2"):                       # c.matchesSideIdentityOutward(direction) (sparse.py:143-150)
            if c[2] == _d.L(direction):
                return (two(c[1], LEFT, 0), sos)
            if c[2] == _d.R(direction):
                return (two(c[1], RIGHT, 0), sos)
            if c[2] == _d.O(direction):
                return (two(c[1], CENTER, None), sos)
            return None
        if k == ("2", "2"):                       # s.matchesCenter(direction, c) (sparse.py:90-97)
            ok = s[1] == c[1] and s[2] == CENTER and direction == c[2]
            return (C if ok else None, sos)
        if k == ("Z", "I"):
            return (s, ss)
        return None
    return contract_sparse(rule, None, side, operator_center)


def expectation_stage1(corner, side):
    """reference tensors/_2d/sparse.py:72-85; rule arguments (l=corner tag, r=side tag)."""
    def rule(c, s):
        k = (kind(c), kind(s))
        std = _standard(*k)
        if std is not False:
            return std
        if k == ("2", "I"):                       # sparse.py:139-142
            return c if c[2] == LEFT else None
        if k == ("I", "2"):                       # sparse.py:119-122
            return s if s[2] in (RIGHT, CENTER) else None
        if k == ("2", "2"):
            return _two_matches(c, s)
        if k == ("Z", "I"):                       # sparse.py:193-196
            return c if c[1] == LEFT else None
        if k == ("I", "Z"):                       # sparse.py:189-192
            return s if s[1] == RIGHT else None
        if k == ("Z", "Z"):
            return _zip_matches(c, s)
        return None
    return contract_sparse(rule, _d.stage1, corner, side)


def expectation_stage2(right, left):
    """reference tensors/_2d/sparse.py:86-99; rule arguments (r=first tensor's tag, l=second's)."""
    def rule(r, l):
        k = (kind(r), kind(l))
        std = _standard(*k)
        if std is not False:
            return std
        if k == ("2", "I"):                       # r.matchesStage1IdentityOnLeft (sparse.py:151-156)
            if r[2] == RIGHT:
                return r
            if r[2] == CENTER:
                return two(r[1], CENTER, LEFT)
            return None
        if k == ("I", "2"):                       # l.matchesStage1IdentityOnRight (sparse.py:157-162)
            if l[2] == LEFT:
                return l
            if l[2] == CENTER:
                return two(l[1], CENTER, RIGHT)
            return None
        if k == ("2", "2"):
            return _two_matches(l, r)
        if k == ("Z", "I"):
            return r if r[1] == RIGHT else None
        if k == ("I", "Z"):
            return l if l[1] == LEFT else None
        if k == ("Z", "Z"):
            return _zip_matches(l, r)
        return None
    return contract_sparse(rule, _d.stage2, right, left)


def stage3_term_allowed(x, y, z):
    """reference tensors/_2d/sparse.py:101-114 (rule table of formExpectationStage3)."""
    k = (kind(x), kind(y), kind(z))
    if k in (("C", "I", "I"), ("I", "C", "I"), ("I", "I", "1")):
        return True
    if k == ("2", "I", "2"):                      # x.matchesCenterForStage3(0,z) (sparse.py:102-108)
        return x[1] == z[1] and x[2] == CENTER and x[3] + 0 == z[2]
    if k == ("I", "2", "2"):
        return y[1] == z[1] and y[2] == CENTER and y[3] + 2 == z[2]
    if k == ("2", "2", "I"):                      # sparse.py:163-169
        return x[1] == y[1] and (x[2], y[2]) in ((LEFT, RIGHT), (RIGHT, LEFT)) and x[3] == y[3]
    if k == ("Z", "Z", "I"):
        return (x[1], y[1]) in ((LEFT, RIGHT), (RIGHT, LEFT))
    return False


def stage3_terms(s2_0, s2_1, operator_center):
    """Term list [(tag0, tag1, tagc)] in the reference's itertools.product order (sparse.py:119-127)."""
    return [
        (x, y, z)
        for x in s2_0 for y in s2_1 for z in operator_center
        if stage3_term_allowed(x, y, z)
    ]


def make_sparse_operator(Os=(), OO_UDs=(), OO_LRs=()):
    """reference sparse.py:295-339 (makeSparseOperator).  Later one-site operators overwrite earlier
    ones because every OneSiteOperator tag is equal (sparse.py:36-37)."""
    op = {}
    ident = None
    import numpy as np
    for o in Os:
        op[ONE] = np.asarray(o, dtype=np.complex128)
        ident = np.eye(op[ONE].shape[0], dtype=np.complex128) if ident is None else ident
    for id_, (ou, od) in enumerate(OO_UDs):
        op[two(id_, 3, 0)] = np.asarray(ou, dtype=np.complex128)
        op[two(id_, 1, 0)] = np.asarray(od, dtype=np.complex128)
        ident = np.eye(op[two(id_, 3, 0)].shape[0], dtype=np.complex128) if ident is None else ident
    for id_, (ol, orr) in enumerate(OO_LRs):
        op[two(id_, 0, 0)] = np.asarray(ol, dtype=np.complex128)
        op[two(id_, 2, 0)] = np.asarray(orr, dtype=np.complex128)
        ident = np.eye(op[two(id_, 0, 0)].shape[0], dtype=np.complex128) if ident is None else ident
    if ident is None:
        raise ValueError("No terms have been specified.")
    op[I] = ident
    return op


def from_reference_tag(tag):
    """Translate a reference tag object (duck-typed by class name) into the tuple form.  Used only by the
    golden-vector generator."""
    name = type(tag).__name__
    if name == "Identity":
        return I
    if name == "Complete":
        return C
    if name == "OneSiteOperator":
        return ONE
    if name == "TwoSiteOperator":
        return two(tag.id, tag.direction, tag.position)
    if name == "TwoSiteOperatorCompressed":
        return zipd(tag.direction)
    raise TypeError(name)
