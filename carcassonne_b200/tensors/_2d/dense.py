"""Dense contraction recipes of the 2D corner/side environment on device tensors.

Same functions, argument order and output layouts as the reference's ``carcassonne/tensors/_2d/dense.py``; every
recipe here is one (batched) DMMA GEMM whose epilogue scatters the product straight into the layout the reference
obtains with a trailing ``join`` (a full transposing copy of up to 12 axes), via the offset tables of
``carc_zgemm_tab``.  ``accumulate_into`` lets the sparse layer fold the ``result[tag] += ...`` of
``contractSparseTensors`` (reference sparse.py:236) into the same epilogue (beta = 1).

Leg conventions (reference sketches/tensors/dense/{corner,side}.svg): corner [0 1 2 | 3 4 5], side
[0 1 2 | 3 4 5 | 6 7] with (state, state*, operator) triples on the left / right bonds and (state, state*) towards
the center; center [right, up, left, down, physical].
"""
from math import prod

import numpy as np

import ctypes as C

from ... import _lib as _la
from ..._lib import check, lib
from ...data import DeviceData, _empty, _ptr, _stream, gemm, gemm_scatter
from ...utils import DimensionMismatchError, L, Multiplier, O, R, UnexpectedTensorRankError

OP_N, OP_T, OP_C, OP_J = _la.OP_N, _la.OP_T, _la.OP_C, _la.OP_J


def _check_ranks(*pairs):
    for number, (tensor, rank) in enumerate(pairs):
        if tensor.ndim != rank:
            raise UnexpectedTensorRankError(number, rank, tensor.ndim)


def _check_bond(lt, li, ltensor, rt, ri, rtensor):
    if ltensor.shape[li] != rtensor.shape[ri]:
        raise DimensionMismatchError(lt, li, ltensor.shape[li], rt, ri, rtensor.shape[ri])


def _target(shape, accumulate_into):
    if accumulate_into is None:
        return _empty(shape), 0.0
    if accumulate_into.shape != tuple(shape):
        raise ValueError("cannot accumulate a result of shape {} into {}".format(tuple(shape), accumulate_into.shape))
    accumulate_into._touch()
    return accumulate_into._t, 1.0


def _shape(t):
    """int64[] of a tensor's shape for the one-call-per-recipe entry points of the C ABI."""
    shape = t if isinstance(t, (tuple, list)) else t.shape
    return (C.c_int64 * len(shape))(*shape)


def _row_major_strides(dims):
    strides, acc = [], 1
    for d in reversed(dims):
        strides.append(acc)
        acc *= d
    return strides[::-1]


# -- side -> corner -----------------------------------------------------------------------------------------
def absorbDenseSideIntoCornerFromLeft(corner, side, accumulate_into=None):
    """reference dense.py:11-15: sum corner(0,1,2) = side(3,4,5)  ->  [s0][s1][s2][c3 s6][c4 s7][c5]."""
    _check_ranks((corner, 6), (side, 8))
    for a in range(3):
        _check_bond(0, a, corner, 1, 3 + a, side)
    c, s = corner.shape, side.shape
    out_shape = (s[0], s[1], s[2], c[3] * s[6], c[4] * s[7], c[5])
    out, beta = _target(out_shape, accumulate_into)
    # per leading side index b = (s0 s1 s2):  C_b[(c3 c4 c5), (s6 s7)] = corner[K, (c3 c4 c5)]^T . side_b[K, (s6 s7)],
    # scattered into the joined layout by the GEMM epilogue (csrc/recipes.cu)
    check(lib.carc_absorb_side_into_corner(_ptr(corner._t), _shape(corner), _ptr(side._t), _shape(side), 1, _ptr(out),
                                           int(beta != 0.0), _stream()))
    return accumulate_into if accumulate_into is not None else DeviceData(out)


def absorbDenseSideIntoCornerFromRight(corner, side, accumulate_into=None):
    """reference dense.py:17-21: sum corner(3,4,5) = side(0,1,2)  ->  [c0 s6][c1 s7][c2][s3][s4][s5]."""
    _check_ranks((corner, 6), (side, 8))
    for a in range(3):
        _check_bond(0, 3 + a, corner, 1, a, side)
    c, s = corner.shape, side.shape
    out_shape = (c[0] * s[6], c[1] * s[7], c[2], s[3], s[4], s[5])
    out, beta = _target(out_shape, accumulate_into)
    check(lib.carc_absorb_side_into_corner(_ptr(corner._t), _shape(corner), _ptr(side._t), _shape(side), 0, _ptr(out),
                                           int(beta != 0.0), _stream()))
    return accumulate_into if accumulate_into is not None else DeviceData(out)


# -- center -> side -----------------------------------------------------------------------------------------
def _double_layer(direction, center, center_conj, operator):
    """E[(g h), (vL wL vR wR vO wO)] = sum_{s,z} center[.., s] O[z, s] conj[.., z] with g / h the center / conjugate
    legs facing side `direction`, L / R / O its left, right and opposite legs (reference dense.py:23-81 contracts
    the same three tensors pairwise; here they are pre-contracted once and shared by every sparse tag pair)."""
    _check_ranks((center, 5), (center_conj, 5))
    n, m = center.shape, center_conj.shape
    if n[4] != m[4]:
        raise DimensionMismatchError(1, 4, n[4], 2, 4, m[4])
    d = n[4]
    if operator is not None:
        if operator.ndim != 2:
            raise UnexpectedTensorRankError(3, 2, operator.ndim)
        if operator.shape != (d, d):
            raise DimensionMismatchError(3, 1, operator.shape[1], 1, 4, d)
    i, l, r, o = direction, L(direction), R(direction), O(direction)
    dims = (n[i], m[i], n[l], m[l], n[r], m[r], n[o], m[o])
    E = _empty((n[i] * m[i], prod(dims[2:])))
    # vo[.., z] = sum_s v[.., s] O[z, s], then a K = d product scattered into [(g h), (vL wL vR wR vO wO)]: one call
    check(lib.carc_double_layer_center(direction, _ptr(center._t), _shape(center), _ptr(center_conj._t),
                                       _shape(center_conj), _ptr(operator._t) if operator is not None else None, _ptr(E),
                                       _stream()))
    return E, dims


def _absorb_center(direction, side, E, dims, accumulate_into):
    _check_ranks((side, 8))
    s = side.shape
    if s[6] != dims[0]:
        raise DimensionMismatchError(0, 6, s[6], 1, direction, dims[0])
    if s[7] != dims[1]:
        raise DimensionMismatchError(0, 7, s[7], 2, direction, dims[1])
    nl, ml, nr, mr, no, mo = dims[2:]
    out_shape = (s[0] * nl, s[1] * ml, s[2], s[3] * nr, s[4] * mr, s[5], no, mo)
    out, beta = _target(out_shape, accumulate_into)
    check(lib.carc_absorb_center_into_side(_ptr(side._t), _shape(side), _ptr(E), _shape(dims), _ptr(out), int(beta != 0.0),
                                           _stream()))
    if accumulate_into is not None:
        return accumulate_into
    result = DeviceData(out)
    # side' = side (x) E over (g h): the state-bond compression that follows a contraction builds its Gram matrix
    # from these two small factors instead of the enlarged tensor (compression.factored_side_gram)
    result._factors = ("center_into_side", side, DeviceData(E), dims)
    return result


def absorbDenseCenterSSIntoSide(direction, side, center, center_conj, accumulate_into=None, _cache=None):
    """reference dense.py:23-49: side(6) = center(direction), side(7) = conj(direction), physical legs tied
    ->  [s0 v_L][s1 w_L][s2][s3 v_R][s4 w_R][s5][v_O][w_O]."""
    return absorbDenseCenterSOSIntoSide(direction, side, center, None, center_conj, accumulate_into, _cache)


def absorbDenseCenterSOSIntoSide(direction, side, center, operator, center_conj, accumulate_into=None, _cache=None):
    """reference dense.py:51-81: as above with the site operator O[z, s] between the physical legs (leg 0 to the
    conjugate, leg 1 to the state)."""
    key = id(operator)
    if _cache is not None and key in _cache:
        E, dims = _cache[key]
    else:
        E, dims = _double_layer(direction, center, center_conj, operator)
        if _cache is not None:
            _cache[key] = (E, dims)
    return _absorb_center(direction, side, E, dims, accumulate_into)


# -- environment halves -------------------------------------------------------------------------------------
def formNormalizationStage1(corner, side, accumulate_into=None):
    """reference dense.py:96-99: sum corner(3,4,5) = side(0,1,2)  ->  [(c0 c1 c2)][(s3 s4 s5)][s6][s7]."""
    _check_ranks((corner, 6), (side, 8))
    for a in range(3):
        _check_bond(0, 3 + a, corner, 1, a, side)
    c, s = corner.shape, side.shape
    out_shape = (c[0] * c[1] * c[2], s[3] * s[4] * s[5], s[6], s[7])
    out, beta = _target(out_shape, accumulate_into)
    check(lib.carc_form_stage1(_ptr(corner._t), _shape(corner), _ptr(side._t), _shape(side), _ptr(out), int(beta != 0.0),
                               _stream()))
    return accumulate_into if accumulate_into is not None else DeviceData(out)


def stage2SlabPlan(a, b, half, slab=None):
    """Launch plan of stage 2 as ONE batched GEMM  C_b[(A1 A2 A3), (B2 B3)] = A[K, (A1 A2 A3)]^T . B_b[K, (B2 B3)]
    (b = B0) with a scatter epilogue.  ``a`` / ``b``: shapes of the two stage-1 tensors; ``half`` = None (reference
    layout), 0 or 1 (stage-3 streaming layouts); ``slab`` = (rank, world) restricts the result to that rank's slab of
    the joined environment bond X -- contiguous in the SLOW factor of X, which is B0 for half 0 (a range of the GEMM
    batch) and A1 for half 1 (a range of the GEMM rows), the same index range on both halves because half 0's B0 and
    half 1's A1 are the two ends of the same ring bond ("each GPU builds its X-slab: slice x of
    the stage-1 tensors").  Pure host logic (no device call): the CPU suite checks it against NumPy slicing."""
    K = a[0]
    Nn = b[2] * b[3]
    plan = {"K": K, "N": Nn, "lda": a[1] * a[2] * a[3], "ldb": Nn, "strideB": K * Nn, "a_offset": 0, "b_offset": 0}
    if half is None:
        if slab is not None:
            raise ValueError("X slabs exist only in the stage-3 layouts (half = 0 or 1)")
        st = _row_major_strides((a[1], a[2], b[2], a[3], b[3]))
        plan.update(out_shape=(b[0], a[1], a[2], b[2], a[3], b[3]), M=a[1] * a[2] * a[3], batch=b[0],
                    rows=((a[1], st[0]), (a[2], st[1]), (a[3], st[3])), cols=((b[2], st[2]), (b[3], st[4])),
                    strideC=a[1] * a[2] * b[2] * a[3] * b[3], full_X=b[0] * a[1], x_range=(0, b[0] * a[1]))
        return plan
    rest = a[3] * b[3] * a[2] * b[2]
    st = _row_major_strides((a[3], b[3], a[2], b[2]))
    slow = b[0] if half == 0 else a[1]
    lo, hi = (0, slow) if slab is None else (slow * slab[0] // slab[1], slow * (slab[0] + 1) // slab[1])
    n_b0, n_a1 = (hi - lo, a[1]) if half == 0 else (b[0], hi - lo)
    a1_stride, stride_c = (rest, n_a1 * rest) if half == 0 else (n_b0 * rest, rest)
    inner = a[1] if half == 0 else b[0]
    plan.update(out_shape=(n_b0 * n_a1, a[3], b[3], a[2], b[2]), M=n_a1 * a[2] * a[3], batch=n_b0,
                rows=((n_a1, a1_stride), (a[2], st[2]), (a[3], st[0])), cols=((b[2], st[3]), (b[3], st[1])),
                strideC=stride_c, full_X=b[0] * a[1], x_range=(lo * inner, hi * inner))
    if half == 0:
        plan["b_offset"] = lo * K * Nn          # batches lo .. hi of B
    else:
        plan["a_offset"] = lo * a[2] * a[3]     # rows (A1 = lo .. hi) of the K x (A1 A2 A3) operand, lda unchanged
    return plan


def formNormalizationStage2(stage1_a, stage1_b, accumulate_into=None, half=None, slab=None):
    """reference dense.py:102-112: sum A(0) = B(1)  ->  [B0][A1][A2][B2][A3][B3].

    ``half`` = 0 / 1 writes instead the layout stage 3 consumes, [(B0 A1), A3, B3, A2, B2] /
    [(A1 B0), A3, B3, A2, B2] (the pre-joins of reference dense.py:130-131), so the X D^4-element transposing copy in
    front of every matvec disappears.  ``slab`` = (rank, world): only this rank's slab of X is built (multi-GPU)."""
    _check_ranks((stage1_a, 4), (stage1_b, 4))
    _check_bond(0, 0, stage1_a, 1, 1, stage1_b)
    p = stage2SlabPlan(stage1_a.shape, stage1_b.shape, half, slab)      # the plan carc_form_stage2 executes (shapes here)
    out, beta = _target(p["out_shape"], accumulate_into)
    rank, world = slab if slab is not None else (0, 1)
    check(lib.carc_form_stage2(_ptr(stage1_a._t), _shape(stage1_a), _ptr(stage1_b._t), _shape(stage1_b),
                               -1 if half is None else int(half), int(rank), int(world), _ptr(out), int(beta != 0.0),
                               _stream()))
    return accumulate_into if accumulate_into is not None else DeviceData(out)


def prejoinStage2(stage2, half):
    """reference dense.py:130-131 / 162-163 on a stage-2 tensor in the reference layout [x, y, Da, Db, Da*, Db*]."""
    _check_ranks((stage2, 6))
    return stage2.join((0, 1), 4, 5, 2, 3) if half == 0 else stage2.join((1, 0), 4, 5, 2, 3)


def unjoinStage2(joined, half, x, y):
    """Inverse of prejoinStage2 (x, y: the two environment bond extents of the reference layout)."""
    X, c, d, a, b = joined.shape
    if half == 0:
        return joined.split(x, y, c, d, a, b).join(0, 1, 4, 5, 2, 3)
    return joined.split(y, x, c, d, a, b).join(1, 0, 4, 5, 2, 3)


# -- stage 3 -------------------------------------------------------------------------------------------------
def stage3CostOfMultiply(A, B, d, with_operator, full_X=None):
    """cmac count the reference's CostTracker gives the generated contractors of dense.py:115-128 / 146-160
    (``full_X``: extent of the whole joined bond when A, B hold one rank's slab of it)."""
    X, c, dd, a, b = A.shape
    X = X if full_X is None else full_X
    _, g, h, e, f = B.shape
    cost = X * c * dd * (e * f * d) * (a * b) + (g * h) * (c * dd * d) * (e * f * X)
    if with_operator:
        cost += d * d * a * b * e * f
    return cost


def stage3CostOfFormMatrix(A, B, d, full_X=None):
    X, c, dd, a, b = A.shape
    X = X if full_X is None else full_X
    _, g, h, e, f = B.shape
    m = (a * b * c * dd) * (e * f * g * h)
    return m * X + m * d * d


def stage3FormMatrix(A, B, operator, accumulate_into=None):
    """reference dense.py:176-194: sum_X A[X,..] B[X,..] outer O  ->  [(D0* D1* D2* D3* s')][(D0 D1 D2 D3 s)].

    A = [X, D0*, D1*, D0, D1], B = [X, D2*, D3*, D2, D3] pre-joined, O [s', s] a host or device d x d matrix."""
    X, c, dd, a, b = A.shape
    Xb, g, h, e, f = B.shape
    if X != Xb:
        raise DimensionMismatchError(0, 0, X, 1, 0, Xb)
    op = np.asarray(operator.toArray() if hasattr(operator, "toArray") else operator, dtype=np.complex128)
    d = op.shape[0]
    P, Q, Rr, S = c * dd, a * b, g * h, e * f
    n_out, n_in = P * Rr * d, Q * S * d
    out, _ = _target((n_out, n_in), accumulate_into)
    if d > 4:
        raise NotImplementedError("site operators beyond d = 4 are not supported on device")
    opd = DeviceData.fromArray(op.reshape(d * d))
    # one library call: G[(P Q),(R S)] = sum_X A B (DMMA GEMM), then matrix[(P R s'),(Q S s)] += G * O[s', s]
    # (csrc/recipes.cu: carc_stage3_form_matrix); an empty X slab (more ranks than slow bond indices) contributes nothing
    check(lib.carc_stage3_form_matrix(_ptr(A._t), _ptr(B._t), X, P, Q, Rr, S, _ptr(opd._t), d, _ptr(out),
                                      int(accumulate_into is not None), _stream()))
    return accumulate_into if accumulate_into is not None else DeviceData(out)


def _stage3_multiplier(A, B, operator, d, sharding=None, full_X=None):
    """``sharding`` (a ``distributed.EnvironmentSharding``): A, B hold this rank's slab of the ``full_X`` joined bond;
    the operator then ends with the all-reduce over ranks and ``formMatrix`` sums the ranks' partial matrices."""
    from ...operator import Stage3Operator
    state_shape = (A.shape[3], A.shape[4], B.shape[3], B.shape[4], d)
    if A.shape[1:3] != A.shape[3:5] or B.shape[1:3] != B.shape[3:5]:
        raise ValueError("stage-2 halves must carry equal state / conjugate-state bonds")
    op = Stage3Operator(state_shape)
    if A.shape[0] > 0:
        op.add_term(A, B, operator)
    op.finalize()
    n = prod(state_shape)
    identity = np.eye(d, dtype=np.complex128)

    def form_matrix():
        matrix = stage3FormMatrix(A, B, identity if operator is None else operator)
        return sharding.sum_matrix_(matrix) if sharding is not None else matrix

    if sharding is not None:
        sharding.attach(op)
    multiplier = Multiplier(
        (n, n),
        op,
        stage3CostOfMultiply(A, B, d, operator is not None, full_X),
        form_matrix,
        stage3CostOfFormMatrix(A, B, d, full_X),
    )
    multiplier.device_operator = op
    multiplier.terms = [(A, B, operator)]
    return multiplier


def formNormalizationStage3(stage2_0, stage2_1, identity):
    """reference dense.py:115-144: the normalization Multiplier of two stage-2 halves (reference layout)."""
    return _stage3_multiplier(prejoinStage2(stage2_0, 0), prejoinStage2(stage2_1, 1), None, identity.shape[0])


def formDenseStage3(stage2_0, stage2_1, operator):
    """reference dense.py:146-203: as above with a site operator on the physical leg."""
    return _stage3_multiplier(prejoinStage2(stage2_0, 0), prejoinStage2(stage2_1, 1), operator, operator.shape[0])


def formNormalizationHalves(corners, sides):
    """-> (A, B, sharding, full_X): the two stage-2 halves in the stage-3 layout -- this rank's X slab of them in the
    multi-GPU mode (``distributed.shard_environment``)."""
    from ... import distributed as _dist
    sharding = _dist.environment_sharding()
    slab = sharding.slab if sharding is not None else None
    stage1 = [formNormalizationStage1(corners[i], sides[i]) for i in range(4)]
    full_X = stage1[1].shape[0] * stage1[0].shape[1]
    return (formNormalizationStage2(stage1[0], stage1[1], half=0, slab=slab),
            formNormalizationStage2(stage1[2], stage1[3], half=1, slab=slab), sharding, full_X)


def formNormalizationMultiplier(corners, sides, center_identity):
    """reference dense.py:82-94."""
    A, B, sharding, full_X = formNormalizationHalves(corners, sides)
    return _stage3_multiplier(A, B, None, center_identity.shape[0], sharding, full_X)


def formNormalizationSubmatrix(corners, sides):
    """reference dense.py:205-225: sum s2_0(0,1) = s2_1(1,0)  ->  [(D0* D1* D2* D3*)][(D0 D1 D2 D3)]."""
    A, B, sharding, _ = formNormalizationHalves(corners, sides)
    matrix = stage3FormMatrix(A, B, np.ones((1, 1), dtype=np.complex128))
    return sharding.sum_matrix_(matrix) if sharding is not None else matrix
