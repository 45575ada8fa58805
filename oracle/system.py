"""The 2D corner/side boundary system, restated on plain ndarrays.

Oracle (test infrastructure) -- see ``oracle/__init__.py``.  Follows reference system/_2d.py and
system/base.py; sparse tensors are dicts {tuple tag: ndarray} (``oracle/tags.py``).
"""
import copy as _copy

import numpy as np
import scipy.linalg as sla

from . import dense as _d
from . import linalg as _l
from . import solver as _s
from . import tags as _t


class InvariantViolatedError(Exception):
    pass


def expectation_and_normalization_multipliers(corners, sides, operator_center):
    """reference tensors/_2d/sparse.py:59-71 + 100-161 (formExpectationAndNormalizationMultipliers)."""
    s1 = [_t.expectation_stage1(corners[i], sides[i]) for i in range(4)]
    s2_0 = _t.expectation_stage2(s1[0], s1[1])
    s2_1 = _t.expectation_stage2(s1[2], s1[3])
    return stage3_multipliers(s2_0, s2_1, operator_center)


def stage3_multipliers(s2_0, s2_1, operator_center):
    terms = _t.stage3_terms(s2_0, s2_1, operator_center)
    d = operator_center[_t.I].shape[0]
    ident = np.eye(d, dtype=np.complex128)
    joined = {}

    def halves(x, y):
        if (x, y) not in joined:
            joined[x, y] = _d.stage3_prejoin(s2_0[x], s2_1[y])
        return joined[x, y]

    def multiply_h(v):
        out = np.zeros(v.shape, dtype=np.complex128)
        for x, y, z in terms:
            A, B = halves(x, y)
            out += _d.stage3_multiply_joined(A, B, v, None if z == _t.I else operator_center[z])
        return out

    def matrix_h():
        m = 0
        for x, y, z in terms:
            m = m + _d.stage3_form_matrix(s2_0[x], s2_1[y], operator_center[z])
        return m

    a, b = s2_0[_t.I].shape[2:4]
    e, f = s2_1[_t.I].shape[2:4]
    n = a * b * e * f * d
    cost_mul = sum(
        _d.stage3_cost_of_multiply(s2_0[x].shape, s2_1[y].shape, d, z != _t.I) for x, y, z in terms
    )
    cost_mat = sum(_d.stage3_cost_of_form_matrix(s2_0[x].shape, s2_1[y].shape, d) for x, y, z in terms)
    H = _s.Mult((n, n), multiply_h, cost_mul, matrix_h, cost_mat)

    def multiply_n(v):
        A, B = halves(_t.I, _t.I)
        return _d.stage3_multiply_joined(A, B, v, None)

    N = _s.Mult(
        (n, n), multiply_n,
        _d.stage3_cost_of_multiply(s2_0[_t.I].shape, s2_1[_t.I].shape, d, False),
        lambda: _d.stage3_form_matrix(s2_0[_t.I], s2_1[_t.I], ident),
        _d.stage3_cost_of_form_matrix(s2_0[_t.I].shape, s2_1[_t.I].shape, d),
    )
    H.terms = terms
    return H, N


def side_from_corner(corner_id, direction):  # _2d.py:573-575
    return (corner_id + 1 - direction) % 4


class System:
    """reference system/_2d.py:19-569."""

    def __init__(self, corners, sides, center, operator_center, center_conj=None):
        self.corners = [dict(c) for c in corners]
        self.sides = [dict(s) for s in sides]
        self.center = center
        self.center_conj = center.conj() if center_conj is None else center_conj
        self.operator_center = dict(operator_center)
        self.just_increased_bandwidth = False

    # -- constructors -------------------------------------------------------------------------------
    @classmethod
    def new_trivial(cls, operator_center):
        """_2d.py:71-86."""
        d = next(iter(operator_center.values())).shape[0]
        one6 = np.ones((1,) * 6, dtype=np.complex128)
        one8 = np.ones((1,) * 8, dtype=np.complex128)
        center = np.full((1, 1, 1, 1, d), 1.0 / np.sqrt(d), dtype=np.complex128)
        return cls([{_t.I: one6.copy()} for _ in range(4)], [{_t.I: one8.copy()} for _ in range(4)],
                   center, operator_center)

    def copy(self):
        """_2d.py:122-131: shallow -- dicts copied, tensors shared."""
        s = System(self.corners, self.sides, self.center, self.operator_center, self.center_conj)
        return s

    def set_center(self, center, center_conj=None):  # _2d.py:550-557
        self.center = center
        self.center_conj = center.conj() if center_conj is None else center_conj
        self.just_increased_bandwidth = False

    # -- expectation --------------------------------------------------------------------------------
    def multipliers(self, operator_center=None):  # _2d.py:455-459
        return expectation_and_normalization_multipliers(
            self.corners, self.sides, self.operator_center if operator_center is None else operator_center)

    def scalar(self, mult):  # _2d.py:429-431
        return np.tensordot(self.center_conj, mult(self.center), axes=(range(5), range(5)))[()]

    def expectation_and_normalization(self, operator_center=None):  # _2d.py:373-378
        H, N = self.multipliers(operator_center)
        e = self.scalar(H)
        n = self.scalar(N)
        return e / n, n

    def expectation(self, operator_center=None):
        return self.expectation_and_normalization(operator_center)[0]

    def normalization(self):  # _2d.py:390-392
        s2_0, s2_1 = _d.normalization_halves(
            [c[_t.I] for c in self.corners], [s[_t.I] for s in self.sides])
        return self.scalar(lambda v: _d.stage3_multiply(s2_0, s2_1, v))

    def strip(self):  # _2d.py:558-567
        return System(
            [{_t.I: c[_t.I]} for c in self.corners], [{_t.I: s[_t.I]} for s in self.sides],
            self.center, {_t.I: self.operator_center[_t.I]}, self.center_conj)

    def one_site_expectation(self):
        """_2d.py:396-428."""
        total = 0
        stripped = self.strip()
        for tag, value in self.operator_center.items():
            if _t.kind(tag) == "1":
                s = stripped.copy()
                s.operator_center[tag] = value
                total += s.expectation()
            if _t.kind(tag) == "2":
                if tag[3] != 0:
                    continue
                for direction, partner in ((0, 2), (1, 3)):
                    if tag[2] == direction:
                        other = _t.two(tag[1], partner, 0)
                        s = stripped.copy()
                        s.operator_center[tag] = value
                        s.operator_center[other] = self.operator_center[other]
                        s.contract_towards(direction)
                        total += s.expectation()
        return total

    def estimated_one_site_expectation(self, direction=0):  # base.py:58-64
        s = self.copy()
        e1 = s.expectation()
        s.contract_towards(direction)
        return s.expectation() - e1

    # -- absorption ---------------------------------------------------------------------------------
    def contract_unnormalized_towards(self, direction, center=None, center_conj=None):  # _2d.py:443-454
        if center is None:
            center, center_conj = self.center, self.center_conj
        if center_conj is None:
            center_conj = center.conj()
        i = direction
        self.corners[i] = _t.absorb_side_into_corner_from_left(self.corners[i], self.sides[_d.L(i)])
        self.sides[i] = _t.absorb_center_sos_into_side(i, self.sides[i], center, self.operator_center, center_conj)
        self.corners[_d.R(i)] = _t.absorb_side_into_corner_from_right(self.corners[_d.R(i)], self.sides[_d.R(i)])
        if self.just_increased_bandwidth:
            raise InvariantViolatedError("optimize or replace the center before contracting it")

    def contract_towards(self, direction):  # _2d.py:435-439
        iso, _, den = _l.normalize_axis(self.center, _d.O(direction))
        self.contract_unnormalized_towards(direction, iso)
        self.set_center(_l.absorb_matrix_at(_l.normalize_axis(self.center, direction)[0], direction, den))

    # -- optimisation -------------------------------------------------------------------------------
    def minimize_expectation(self, stats=None):  # _2d.py:489-497
        H, N = self.multipliers()
        self.set_center(_s.relax_over(self.center, H, N, maximum_number_of_multiplications=100, stats=stats))

    def minimize_expectation_full(self):  # _2d.py:498-502
        H, N = self.multipliers()
        evals, evecs = sla.eigh(H.form_matrix(), N.form_matrix())
        self.set_center(evecs[:, 0].reshape(self.center.shape))
        return evals[0]

    # -- compression --------------------------------------------------------------------------------
    def compress_corner_state_towards(self, corner_id, direction, new, initial=None):  # _2d.py:191-228
        am = _l.absorb_matrix_at
        if direction == 0:
            sid = _d.L(corner_id)
            s = self.sides[sid][_t.I]
            c = self.corners[corner_id][_t.I]
            sj = np.ascontiguousarray(s.transpose(0, 1, 2, 6, 7, 3, 4, 5)).reshape(-1, s.shape[3], s.shape[4], s.shape[5])
            cj = c.reshape(c.shape[0], c.shape[1], c.shape[2], -1)
            comp = _s.product_compressor(sj, cj, new, initial)
            self.sides[sid] = {t: am(am(x, 3, comp), 4, comp.conj()) for t, x in self.sides[sid].items()}
            self.corners[corner_id] = {t: am(am(x, 0, comp.conj()), 1, comp) for t, x in self.corners[corner_id].items()}
        elif direction == 1:
            c = self.corners[corner_id][_t.I]
            s = self.sides[corner_id][_t.I]
            cj = c.reshape(-1, c.shape[3], c.shape[4], c.shape[5])
            sj = s.reshape(s.shape[0], s.shape[1], s.shape[2], -1)
            comp = _s.product_compressor(cj, sj, new, initial)
            self.corners[corner_id] = {t: am(am(x, 3, comp), 4, comp.conj()) for t, x in self.corners[corner_id].items()}
            self.sides[corner_id] = {t: am(am(x, 0, comp.conj()), 1, comp) for t, x in self.sides[corner_id].items()}
        else:
            raise ValueError("compression direction must be 0 or 1, not " + str(direction))
        return comp

    # -- bandwidth ----------------------------------------------------------------------------------
    def increase_bandwidth(self, direction, by=None, to=None, do_as_much_as_possible=False, enlargeners=None,
                           sample=None):
        """_2d.py:480-484 -> base.py:114-160.  `sample` is the random (new x old) draw newEnlargener would make."""
        if direction not in (0, 1):
            raise ValueError("Direction for bandwidth increase must be 0 or 1, not {}.".format(direction))
        c = self.center
        axis, oaxis = direction, _d.O(direction)
        d = c.shape[-1]
        old = c.shape[axis]
        if (by is None) == (to is None):
            raise ValueError("exactly one of by= and to= must be given")
        new = old + by if by is not None else to
        if new < old:
            raise ValueError("new dimension must not be smaller than the old one")
        if new == old:
            return None
        if new > d * old:
            if do_as_much_as_possible:
                new = d * old
            else:
                raise ValueError("New dimension must be <= physical dimension times old dimension.")
        n0 = _l.normalize_axis(c, oaxis)[0]
        n1 = _l.normalize_axis(c, axis)[0]
        if enlargeners is None:
            if sample is None:
                sample = _l.random_complex(np.random, new, old)
            ea, eb = _l.enlargener_from_random(sample)
        else:
            ea, eb = enlargeners
        am = _l.absorb_matrix_at
        c = am(c, axis, ea)
        n0 = am(n0, oaxis, eb)
        n0, c = _l.normalize_axis_and_denormalize(n0, oaxis, axis, c)
        n1 = am(n1, axis, ea)
        c = am(c, oaxis, eb)
        n1, c = _l.normalize_axis_and_denormalize(n1, axis, oaxis, c)
        self.set_center(c)
        self.contract_unnormalized_towards(axis, n0)
        self.contract_unnormalized_towards(oaxis, n1)
        self.just_increased_bandwidth = True
        return ea, eb

    # -- snapshots ----------------------------------------------------------------------------------
    def snapshot(self):
        return _copy.deepcopy((self.corners, self.sides, self.center, self.operator_center))
