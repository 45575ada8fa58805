"""Sparse operator tags and the host-side planner that decides which dense device kernels run.

A sparse boundary tensor is a dict ``{tag: DeviceData}`` exactly as in the reference's ``carcassonne/sparse.py``
(tags 24-208, ``contractSparseTensors`` 223-240, ``makeSparseOperator`` 295-339).  The tag classes below expose
the reference's constructors, fields, equality and ``repr``; the pairing rules are stated as small tables keyed on
tag kinds, and the planner turns two tag dicts into a flat list of (tag_1, tag_2, result tag, variant) products --
the device then executes each product as one GEMM whose epilogue accumulates into the result tag's buffer.
"""
from .utils import L, O, R

LEFT, RIGHT, CENTER = 0, 1, 2


# -- tags ---------------------------------------------------------------------------------------------------------
class _Tag:
    __slots__ = ()
    kind = "?"

    def _key(self):
        return ()

    def __eq__(self, other):
        return type(other) is type(self) and other._key() == self._key()

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((self.kind,) + self._key())


class Identity(_Tag):
    """No operator inside this piece of the environment."""
    __slots__ = ()
    kind = "I"
    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = object.__new__(cls)
        return cls._instance

    def __repr__(self):
        return "Identity()"


class Complete(_Tag):
    """One full Hamiltonian term already inside."""
    __slots__ = ()
    kind = "C"
    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = object.__new__(cls)
        return cls._instance

    def __repr__(self):
        return "Complete()"


class OneSiteOperator(_Tag):
    """The reference discards the id (sparse.py:36-37: ``self.id = None``), so every one-site operator shares a
    key and later ones overwrite earlier ones in ``makeSparseOperator``; reproduced."""
    __slots__ = ("id",)
    kind = "1"

    def __init__(self, id):
        self.id = None

    def _key(self):
        return (self.id,)

    def __repr__(self):
        return "OneSiteOperator({})".format(self.id)


class TwoSiteOperator(_Tag):
    """Half of two-site term ``id``.  In the center operator ``direction`` is the neighbour holding the partner
    (0 right, 1 up, 2 left, 3 down); in the environment it is LEFT / RIGHT / CENTER, the leg through which the
    term will be completed, and ``position`` counts sites along the boundary (reference sparse.py:55-172)."""
    __slots__ = ("id", "direction", "position")
    kind = "2"

    def __init__(self, id, direction, position=None):
        self.id = id
        self.direction = direction
        self.position = position

    def _key(self):
        return (self.id, self.direction, self.position)

    def __repr__(self):
        return "TwoSiteOperator({},{},{})".format(self.id, self.direction, self.position)

    def withNewDirectionAndPosition(self, direction, position=None):
        return TwoSiteOperator(self.id, direction, position)

    def moveOut(self):
        return TwoSiteOperator(self.id, self.direction, self.position + 1)


class TwoSiteOperatorCompressed(_Tag):
    """All two-site halves of one direction folded into an operator bond (reference sparse.py:173-208)."""
    __slots__ = ("direction",)
    kind = "Z"

    def __init__(self, direction):
        self.direction = direction

    def _key(self):
        return (self.direction,)

    def __repr__(self):
        return "TwoSiteOperatorCompressed({})".format(self.direction)


_I, _C = Identity(), Complete()


# -- pairing rules ------------------------------------------------------------------------------------------------
def _halves_meet(left, right):
    """Two halves of the same term facing each other complete it (sparse.py:81-89 / 185-188)."""
    if left.kind == "2":
        ok = left.id == right.id and left.direction == RIGHT and right.direction == LEFT and \
            left.position == right.position
    else:
        ok = left.direction == RIGHT and right.direction == LEFT
    return _C if ok else None


def _standard(k1, k2):
    if k1 == "I" and k2 == "I":
        return _I
    if (k1, k2) in (("C", "I"), ("I", "C")):
        return _C
    return False


def rule_side_into_corner_from_left(corner_tag, side_tag):
    """tensors/_2d/sparse.py:12-23, plus the two rules the reference lacks for compressed operator bonds: a corner's
    compressed halves that point RIGHT survive the absorption of an operator-free side, and a side's compressed
    halves that point LEFT become the new corner's (the compressed twins of the TwoSiteOperator rules; without them
    every compressed term is dropped by the next absorption, which is why the reference ships no operator-
    compression policy)."""
    k = (corner_tag.kind, side_tag.kind)
    std = _standard(*k)
    if std is not False:
        return std
    if k == ("2", "I"):
        return corner_tag.moveOut() if corner_tag.direction == RIGHT else None
    if k == ("I", "2"):
        if side_tag.direction == LEFT:
            return side_tag
        if side_tag.direction == CENTER:
            return side_tag.withNewDirectionAndPosition(RIGHT, 0)
        return None
    if k == ("Z", "I"):
        return corner_tag if corner_tag.direction == RIGHT else None
    if k == ("I", "Z"):
        return side_tag if side_tag.direction == LEFT else None
    if k in (("2", "2"), ("Z", "Z")):
        return _halves_meet(side_tag, corner_tag)
    return None


def rule_side_into_corner_from_right(corner_tag, side_tag):
    """tensors/_2d/sparse.py:24-35.  The contraction joins the corner's RIGHT bond to the side's LEFT bond, so halves
    meet as (left = corner, right = side) for compressed bonds exactly as for TwoSiteOperator tags.  (The reference
    declares the compressed rule with swapped lambda arguments, ``lambda r,l: l.matches(r)``, which pairs the two
    bonds the contraction leaves OPEN and would fail at the ``+=`` with a shape mismatch; corrected here, with the
    compressed twins of the survive / carry-over rules added as in the from-left case.)"""
    k = (corner_tag.kind, side_tag.kind)
    std = _standard(*k)
    if std is not False:
        return std
    if k == ("2", "I"):
        return corner_tag.moveOut() if corner_tag.direction == LEFT else None
    if k == ("I", "2"):
        if side_tag.direction == RIGHT:
            return side_tag
        if side_tag.direction == CENTER:
            return side_tag.withNewDirectionAndPosition(LEFT, 0)
        return None
    if k == ("Z", "I"):
        return corner_tag if corner_tag.direction == LEFT else None
    if k == ("I", "Z"):
        return side_tag if side_tag.direction == RIGHT else None
    if k in (("2", "2"), ("Z", "Z")):
        return _halves_meet(corner_tag, side_tag)
    return None


def rule_center_into_side(direction, side_tag, center_tag):
    """tensors/_2d/sparse.py:36-58 -> (result tag or None, needs the site operator)."""
    k = (side_tag.kind, center_tag.kind)
    if k == ("I", "I"):
        return _I, False
    if k == ("C", "I"):
        return _C, False
    if k == ("I", "1"):
        return _C, True
    if k == ("2", "I"):
        return (side_tag.moveOut() if side_tag.direction != CENTER else None), False
    if k == ("I", "2"):
        d = center_tag.direction
        if d == L(direction):
            return center_tag.withNewDirectionAndPosition(LEFT, 0), True
        if d == R(direction):
            return center_tag.withNewDirectionAndPosition(RIGHT, 0), True
        if d == O(direction):
            return center_tag.withNewDirectionAndPosition(CENTER), True
        return None, True
    if k == ("2", "2"):
        ok = side_tag.id == center_tag.id and side_tag.direction == CENTER and direction == center_tag.direction
        return (_C if ok else None), True
    if k == ("Z", "I"):
        return side_tag, False
    return None, False


def rule_stage1(corner_tag, side_tag):
    """tensors/_2d/sparse.py:72-85."""
    k = (corner_tag.kind, side_tag.kind)
    std = _standard(*k)
    if std is not False:
        return std
    if k == ("2", "I"):
        return corner_tag if corner_tag.direction == LEFT else None
    if k == ("I", "2"):
        return side_tag if side_tag.direction in (RIGHT, CENTER) else None
    if k == ("Z", "I"):
        return corner_tag if corner_tag.direction == LEFT else None
    if k == ("I", "Z"):
        return side_tag if side_tag.direction == RIGHT else None
    if k in (("2", "2"), ("Z", "Z")):
        return _halves_meet(corner_tag, side_tag)
    return None


def rule_stage2(first_tag, second_tag):
    """tensors/_2d/sparse.py:86-99 (first = the stage-1 tensor contracted on its leg 0)."""
    k = (first_tag.kind, second_tag.kind)
    std = _standard(*k)
    if std is not False:
        return std
    if k == ("2", "I"):
        if first_tag.direction == RIGHT:
            return first_tag
        if first_tag.direction == CENTER:
            return first_tag.withNewDirectionAndPosition(CENTER, LEFT)
        return None
    if k == ("I", "2"):
        if second_tag.direction == LEFT:
            return second_tag
        if second_tag.direction == CENTER:
            return second_tag.withNewDirectionAndPosition(CENTER, RIGHT)
        return None
    if k == ("Z", "I"):
        return first_tag if first_tag.direction == RIGHT else None
    if k == ("I", "Z"):
        return second_tag if second_tag.direction == LEFT else None
    if k in (("2", "2"), ("Z", "Z")):
        return _halves_meet(second_tag, first_tag)
    return None


def stage3_term_allowed(x, y, z):
    """tensors/_2d/sparse.py:101-114: which (half 0, half 1, center operator) triples contribute."""
    k = (x.kind, y.kind, z.kind)
    if k in (("C", "I", "I"), ("I", "C", "I"), ("I", "I", "1")):
        return True
    if k == ("2", "I", "2"):
        return x.id == z.id and x.direction == CENTER and x.position == z.direction
    if k == ("I", "2", "2"):
        return y.id == z.id and y.direction == CENTER and y.position + 2 == z.direction
    if k == ("2", "2", "I"):
        return x.id == y.id and (x.direction, y.direction) in ((LEFT, RIGHT), (RIGHT, LEFT)) and \
            x.position == y.position
    if k == ("Z", "Z", "I"):
        return (x.direction, y.direction) in ((LEFT, RIGHT), (RIGHT, LEFT))
    return False


# -- planner / executor -------------------------------------------------------------------------------------------
def planSparseContraction(rule, tags_1, tags_2):
    """Flat product list [(tag_1, tag_2, result_tag, extra)] in the reference's double-loop order
    (sparse.py:223-240); ``rule`` returns a result tag, ``None``, or ``(result_tag, extra)``."""
    plan = []
    for tag_1 in tags_1:
        for tag_2 in tags_2:
            res = rule(tag_1, tag_2)
            extra = None
            if isinstance(res, tuple):
                res, extra = res
            if res is not None:
                plan.append((tag_1, tag_2, res, extra))
    return plan


def contractSparseTensors(rule, dense, tensor_1, tensor_2):
    """Run a plan: ``dense(data_1, data_2, accumulate_into, extra)`` is called once per product; products with
    the same result tag accumulate into one buffer inside the GEMM epilogue."""
    result = {}
    for tag_1, tag_2, tag, extra in planSparseContraction(rule, tensor_1, tensor_2):
        result[tag] = dense(tensor_1[tag_1], tensor_2[tag_2], result.get(tag), extra)
    return result


def mapOverSparseData(f, sparse):
    return {tag: f(data) for tag, data in sparse.items()}


def stripAllButIdentityFrom(sparse):
    return {Identity(): sparse[Identity()]}


def getInformationFromOperatorCenter(operator_center):
    for matrix in operator_center.values():
        return matrix.shape[0], matrix.dtype, type(matrix)
    raise ValueError("operator tensor has no term from which to extract the physical dimension")


# -- operator construction (reference sparse.py:288-339) -------------------------------------------------------------
def makeSparseOperator(Os=[], OO_UDs=[], OO_LRs=[]):
    """{tag: d x d matrix}: one-site terms, then (up, down) and (left, right) halves of two-site terms, then the
    identity.  ``TwoSiteOperator(id, direction, 0)`` holds the half whose partner sits at neighbour ``direction``."""
    operator = {}
    shape = None
    first = None

    def note(matrix):
        nonlocal shape, first
        if shape is None:
            shape, first = matrix.shape, matrix
        elif matrix.shape != shape:
            raise ValueError("incompatible site matrix shapes: {} and {}".format(shape, matrix.shape))

    for id, matrix in enumerate(Os):
        operator[OneSiteOperator(id)] = matrix
        note(matrix)
    for id, (up, down) in enumerate(OO_UDs):
        operator[TwoSiteOperator(id, 3, 0)] = up
        operator[TwoSiteOperator(id, 1, 0)] = down
        note(up)
        note(down)
    for id, (left, right) in enumerate(OO_LRs):
        operator[TwoSiteOperator(id, 0, 0)] = left
        operator[TwoSiteOperator(id, 2, 0)] = right
        note(left)
        note(right)
    if first is None:
        raise ValueError("No terms have been specified.")
    operator[Identity()] = first.newIdentity(shape[0])
    return operator


def makeMPO(I, Os=[], OOs=[]):
    """The 1D twin of makeSparseOperator (reference sparse.py:265-287): MPO tensor [size, size, d, d] with the identity
    on the two end channels, one-site terms from channel 0 to the last, and one channel per two-site term; returns
    (tensor, right boundary, right tags, left boundary, left tags) -- the order of the reference's return value.  Host ndarrays in, host ndarray out (it is a
    handful of d x d blocks; the 1D system uploads it)."""
    import numpy as np
    as_array = lambda x: x.toArray() if hasattr(x, "toArray") else np.asarray(x, dtype=np.complex128)
    I = as_array(I)
    size = 2 + len(OOs)
    last = size - 1
    tensor = np.zeros((size, size, len(I), len(I)), dtype=np.complex128)
    tensor[0, 0] = I
    tensor[last, last] = I
    for O in Os:
        tensor[0, last] += as_array(O)
    for id, OO in enumerate(OOs):
        tensor[0, id + 1] = as_array(OO[1])
        tensor[id + 1, last] = as_array(OO[0])
    middle = [TwoSiteOperator(id, 2) for id in range(len(OOs))]
    return (tensor, [1] + [0] * (size - 1), [Identity()] + middle + [Complete()],
            [0] * (size - 1) + [1], [Complete()] + middle + [Identity()])


def makeSimpleSparseOperator(O=None, OO_UD=None, OO_LR=None):
    return makeSparseOperator(
        [O] if O is not None else [],
        [OO_UD] if OO_UD is not None else [],
        [OO_LR] if OO_LR is not None else [],
    )


__all__ = [
    "Identity", "Complete", "OneSiteOperator", "TwoSiteOperator", "TwoSiteOperatorCompressed",
    "LEFT", "RIGHT", "CENTER",
    "contractSparseTensors", "planSparseContraction", "getInformationFromOperatorCenter", "makeSimpleSparseOperator",
    "makeMPO", "makeSparseOperator", "mapOverSparseData", "stripAllButIdentityFrom", "stage3_term_allowed",
]
