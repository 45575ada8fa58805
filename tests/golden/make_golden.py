#!/usr/bin/env python
"""Generate the golden input/output vectors in this directory by importing the UNMODIFIED reference
(gcross/Carcassonne) from /root/reference.  Run in the build container only:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the vectors it produces are committed as small .npz
fixtures next to this script.  Nothing here is imported by the product or the tests.
"""
import os
import random
import sys

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("CARCASSONNE_REFERENCE", "/root/reference"))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from carcassonne.compression import computeProductCompressor  # noqa: E402
from carcassonne.data import NDArrayData  # noqa: E402
from carcassonne.sparse import (Identity, makeSparseOperator)  # noqa: E402
from carcassonne.system import System  # noqa: E402
from carcassonne.tensors._2d import dense as rdense  # noqa: E402
from carcassonne.utils import relaxOver, computeCompressor, Multiplier  # noqa: E402
from carcassonne import policies as rpol  # noqa: E402

from oracle.tags import from_reference_tag  # noqa: E402

ND = NDArrayData


def seed(s):
    np.random.seed(s)
    random.seed(s)


def rnd(*shape):
    return ND.newRandom(*shape)


def tagkey(prefix, tag):
    return prefix + "|" + repr(from_reference_tag(tag))


def dump_sparse(out, prefix, sparse):
    for tag, data in sparse.items():
        out[tagkey(prefix, tag)] = data.toArray()
    out[prefix + "|order"] = np.array([repr(from_reference_tag(t)) for t in sparse])


# ------------------------------------------------------------------------------------------------------
def dense_recipes():
    out = {}
    seed(11)
    # absorb side into corner
    corner = rnd(2, 3, 2, 3, 2, 1)
    side_l = rnd(2, 2, 1, 2, 3, 2, 3, 2)
    out["afl_corner"], out["afl_side"] = corner.toArray(), side_l.toArray()
    out["afl_out"] = rdense.absorbDenseSideIntoCornerFromLeft(corner, side_l).toArray()
    side_r = rnd(3, 2, 1, 2, 2, 3, 2, 3)
    out["afr_corner"], out["afr_side"] = corner.toArray(), side_r.toArray()
    out["afr_out"] = rdense.absorbDenseSideIntoCornerFromRight(corner, side_r).toArray()
    # absorb center into side, all four directions, with and without operator
    for i in range(4):
        dims = [2, 3, 2, 3]
        g = dims[i]
        side = rnd(2, 3, 1, 3, 2, 2, g, g)
        center = rnd(*dims, 2)
        op = rnd(2, 2)
        out["ss%d_side" % i], out["ss%d_center" % i], out["ss%d_op" % i] = side.toArray(), center.toArray(), op.toArray()
        out["ss%d_out" % i] = rdense.absorbDenseCenterSSIntoSide(i, side, center).toArray()
        out["sos%d_out" % i] = rdense.absorbDenseCenterSOSIntoSide(i, side, center, op).toArray()
    # stages
    corner = rnd(2, 2, 1, 3, 3, 2)
    side = rnd(3, 3, 2, 2, 2, 1, 3, 3)
    s1 = rdense.formNormalizationStage1(corner, side)
    out["st1_corner"], out["st1_side"], out["st1_out"] = corner.toArray(), side.toArray(), s1.toArray()
    a = rnd(4, 5, 2, 3)
    b = rnd(6, 4, 3, 2)
    out["st2_a"], out["st2_b"] = a.toArray(), b.toArray()
    out["st2_out"] = rdense.formNormalizationStage2(a, b).toArray()
    s2_0 = rnd(3, 4, 2, 3, 2, 3)
    s2_1 = rnd(4, 3, 3, 2, 3, 2)
    op = rnd(2, 2)
    v = rnd(2, 3, 3, 2, 2)
    out["st3_s2_0"], out["st3_s2_1"], out["st3_op"], out["st3_v"] = (x.toArray() for x in (s2_0, s2_1, op, v))
    mn = rdense.formNormalizationStage3(s2_0, s2_1, ND.newIdentity(2))
    md = rdense.formDenseStage3(s2_0, s2_1, op)
    out["st3_norm_out"] = mn(v).toArray()
    out["st3_dense_out"] = md(v).toArray()
    out["st3_norm_cost"] = np.array([mn.cost_of_multiply, mn.cost_of_formMatrix])
    out["st3_dense_cost"] = np.array([md.cost_of_multiply, md.cost_of_formMatrix])
    out["st3_dense_matrix"] = md.formMatrix().toArray()
    out["st3_norm_matrix"] = mn.formMatrix().toArray()
    np.savez_compressed(os.path.join(HERE, "dense_recipes.npz"), **out)


def data_ops():
    out = {}
    seed(12)
    for n, (shape, axis) in enumerate([((3, 2, 4, 2, 2), 2), ((4, 3, 2), 0), ((2, 2, 2, 2, 3), 4), ((5, 1, 3), 1)]):
        t = rnd(*shape)
        iso, nrm, den = t.normalizeAxis(axis)
        out["na%d_in" % n] = t.toArray()
        out["na%d_axis" % n] = np.array(axis)
        out["na%d_iso" % n], out["na%d_nrm" % n], out["na%d_den" % n] = iso.toArray(), nrm.toArray(), den.toArray()
        if shape[axis] > 1:
            a, b = t.normalizeAxis(axis, True)
            out["na%d_sqrt_nrm" % n], out["na%d_sqrt_den" % n] = a.toArray(), b.toArray()
        m = rnd(shape[axis] + 1, shape[axis])
        out["am%d_m" % n] = m.toArray()
        out["am%d_out" % n] = t.absorbMatrixAt(axis, m).toArray()
    sample = rnd(5, 3)
    out["enl_sample"] = sample.toArray()
    out["enl_q"] = sample.qr(mode="economic")[0].toArray()
    m = rnd(6, 3)
    out["unitize_in"], out["unitize_out"] = m.toArray(), m.unitize().toArray()
    np.savez_compressed(os.path.join(HERE, "data_ops.npz"), **out)


def make_operator(name, J=0.5):
    X, Y, Z = ND.X, ND.Y, ND.Z
    if name == "tfim":
        return makeSparseOperator(Os=[-Z], OO_LRs=[(X, -J * X)], OO_UDs=[(X, -J * X)])
    if name == "heis":
        return makeSparseOperator(OO_LRs=[(X, X), (Y, Y), (Z, Z)], OO_UDs=[(X, X), (Y, Y), (Z, Z)])
    raise ValueError(name)


def random_system(chi, D, operator):
    """System.newRandom (_2d.py:35-69) with fixed dimensions instead of randint draws."""
    sides = [rnd(chi, chi, 1, chi, chi, 1, D, D) for _ in range(4)]
    for s in sides:
        s += s.join(1, 0, 2, 4, 3, 5, 7, 6).conj()
    corners = [rnd(chi, chi, 1, chi, chi, 1) for _ in range(4)]
    for c in corners:
        c += c.join(1, 0, 2, 4, 3, 5).conj()
    center = rnd(D, D, D, D, 2)
    return System(
        tuple({Identity(): c} for c in corners), tuple({Identity(): s} for s in sides), center, operator)


def dump_system(out, prefix, system):
    for i in range(4):
        dump_sparse(out, "%s.corner%d" % (prefix, i), system.corners[i])
        dump_sparse(out, "%s.side%d" % (prefix, i), system.sides[i])
    out[prefix + ".center"] = system.state_center_data.toArray()


def system_walk(name, chi, D, s, moves):
    out = {}
    seed(s)
    system = random_system(chi, D, make_operator(name))
    dump_system(out, "init", system)
    dump_sparse(out, "operator", system.operator_center_tensor)
    for step, direction in enumerate(moves):
        system.contractUnnormalizedTowards(direction)
    dump_system(out, "walked", system)
    out["moves"] = np.array(moves)
    H, N = system.formExpectationAndNormalizationMultipliers()
    v = rnd(*system.state_center_data.shape)
    out["v"] = v.toArray()
    out["Hv"] = H(v).toArray()
    out["Nv"] = N(v).toArray()
    out["costs"] = np.array([H.cost_of_multiply, H.cost_of_formMatrix, N.cost_of_multiply, N.cost_of_formMatrix])
    if H.shape[0] <= 200:
        out["Hmat"] = H.formMatrix().toArray()
        out["Nmat"] = N.formMatrix().toArray()
    e, n = system.computeExpectationAndNormalization()
    out["expectation"], out["normalization"] = np.array(e), np.array(n)
    # a normalised contraction and its effect on the center
    system.contractTowards(moves[0])
    out["after_ct.center"] = system.state_center_data.toArray()
    e, n = system.computeExpectationAndNormalization()
    out["after_ct.expectation"], out["after_ct.normalization"] = np.array(e), np.array(n)
    np.savez_compressed(os.path.join(HERE, "walk_%s_chi%d_D%d.npz" % (name, chi, D)), **out)


def relax_case():
    """relaxOver on a Hermitian-definite pencil obtained the way minimizeExpectation builds it."""
    out = {}
    seed(21)
    n = 24
    h = rnd(n, n).toArray()
    h = h + h.conj().T
    b = rnd(n, n).toArray()
    nm = b @ b.conj().T + n * np.eye(n)
    v0 = rnd(n).toArray()
    out["H"], out["N"], out["v0"] = h, nm, v0.copy()
    Hm = Multiplier.fromMatrix(ND(h))
    Nm = Multiplier.fromMatrix(ND(nm))
    # operator branch for both (utils.py:814-832): make the matrix look expensive
    Hm.cost_of_formMatrix = 10 ** 12
    Nm.cost_of_formMatrix = 10 ** 12
    res = relaxOver(ND(v0.copy()), Hm, Nm, maximum_number_of_multiplications=100).toArray()
    out["gmres_result"] = res
    out["gmres_rayleigh"] = np.array(np.vdot(res, h @ res) / np.vdot(res, nm @ res))
    Hm = Multiplier.fromMatrix(ND(h)); Hm.cost_of_multiply = 10 ** 12
    Nm = Multiplier.fromMatrix(ND(nm)); Nm.cost_of_multiply = 10 ** 12
    res = relaxOver(ND(v0.copy()), Hm, Nm, maximum_number_of_multiplications=100).toArray()
    out["lu_result"] = res
    out["lu_rayleigh"] = np.array(np.vdot(res, h @ res) / np.vdot(res, nm @ res))
    # no normalization multiplier, tight budget (3 multiplications -> exactly one restart)
    Hm = Multiplier.fromMatrix(ND(h)); Hm.cost_of_multiply = 10 ** 12
    res = relaxOver(ND(v0.copy()), Hm, None, maximum_number_of_multiplications=3).toArray()
    out["one_restart_result"] = res
    np.savez_compressed(os.path.join(HERE, "relax.npz"), **out)


def compressor_case():
    """computeProductCompressor on an exactly compressible product (tests/test_compression.py:11-28) and
    computeCompressor eigenvalues."""
    out = {}
    seed(31)
    l, r, op, new, old = 3, 4, 1, 2, 5
    Lc = rnd(l, new, new, op)
    Lc += Lc.transpose(0, 2, 1, 3).conj()
    Lt = np.zeros((l, old, old, op), dtype=np.complex128)
    Lt[:, :new, :new, :] = Lc.toArray()
    Rc = rnd(new, new, op, r)
    Rc += Rc.transpose(1, 0, 2, 3).conj()
    Rt = np.zeros((old, old, op, r), dtype=np.complex128)
    Rt[:new, :new, :, :] = Rc.toArray()
    out["L"], out["R"], out["new"] = Lt, Rt, np.array(new)
    state = np.random.get_state()
    c = computeProductCompressor(ND(Lt), ND(Rt), new)
    np.random.set_state(state)
    out["initial"] = ND.newRandom(old, new).toArray()      # the draw compression.py:35 made
    out["compressor"] = c.toArray()
    Lc2 = ND(Lt).absorbMatrixAt(1, c).absorbMatrixAt(2, c.conj())
    Rc2 = ND(Rt).absorbMatrixAt(0, c.conj()).absorbMatrixAt(1, c)
    out["compressed_product"] = Lc2.contractWith(Rc2, (1, 2, 3), (0, 1, 2)).toArray()
    out["exact_product"] = ND(Lt).contractWith(ND(Rt), (1, 2, 3), (0, 1, 2)).toArray()
    # one formMatrix evaluation of the ALS (compression.py:11-25), for the einsum restatement
    c0 = ND(out["initial"]).unitize()
    fm = computeProductCompressor.args[0]
    out["als_matrix"] = fm(ND(Lt), c0.conj(), c0.transpose().conj(), c0.transpose(), ND(Rt)).toArray()
    out["als_c0"] = c0.toArray()
    # computeCompressor (utils.py:268-303)
    m = rnd(7, 7).toArray()
    g = m.conj().T @ m
    out["gram"] = g
    for new_dim in (5, 2):
        comp, inv = computeCompressor(7, new_dim, Multiplier.fromMatrix(ND(g)) if False else
                                      Multiplier((7, 7), lambda v: g @ v, 49, lambda: g, 0), np.complex128, True)
        out["cc%d_comp" % new_dim], out["cc%d_inv" % new_dim] = comp, inv
    np.savez_compressed(os.path.join(HERE, "compressor.npz"), **out)


def bandwidth_case():
    out = {}
    seed(41)
    system = random_system(1, 2, make_operator("tfim"))
    for d in range(4):
        system.contractTowards(d)
    dump_system(out, "before", system)
    dump_sparse(out, "operator", system.operator_center_tensor)
    state = np.random.get_state()
    system.increaseBandwidth(0, by=1)
    np.random.set_state(state)
    out["sample"] = ND.newRandom(3, 2).toArray()
    dump_system(out, "after", system)
    system.minimizeExpectation()
    e, n = system.computeExpectationAndNormalization()
    out["after_min.expectation"], out["after_min.normalization"] = np.array(e), np.array(n)
    np.savez_compressed(os.path.join(HERE, "bandwidth.npz"), **out)


def run_case():
    """End-to-end runs the reference's own simulator tests perform (tests/test_simulator_2d_in_1d.py:36-47,
    tests/test_simulator_2d_in_15d.py:11-21) with their iteration counts and energies."""
    out = {}
    for direction in (0, 1):
        seed(51 + direction)
        kw = {"OO_LR" if direction == 0 else "OO_UD": [ND.X, -0.01 * ND.X]}
        system = System.newTrivialWithSimpleSparseOperator(O=-ND.Z, **kw)
        system.setPolicy("sweep convergence", rpol.RelativeStateDifferenceThresholdConvergencePolicy(1e-5))
        system.setPolicy("run convergence", rpol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-7))
        system.setPolicy("bandwidth increase", rpol.OneDirectionIncrementBandwidthIncreasePolicy(direction, 2))
        system.setPolicy("contraction", rpol.RepeatPatternContractionPolicy([0 + direction, 2 + direction]))
        system.runUntilConverged()
        out["tfim1d_dir%d_energy" % direction] = np.array(system.computeOneSiteExpectation())
        out["tfim1d_dir%d_counts" % direction] = np.array([system.number_of_sweeps, system.number_of_iterations])
        out["tfim1d_dir%d_shape" % direction] = np.array(system.state_center_data.shape)
    seed(61)
    system = System.newTrivialWithSimpleSparseOperator(O=ND.Z)
    system.setPolicy("state compression", rpol.ConstantStateCompressionPolicy(1))
    system.setPolicy("sweep convergence", rpol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-7))
    system.setPolicy("run convergence", rpol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-7))
    system.setPolicy("bandwidth increase", rpol.AllDirectionsIncrementBandwidthIncreasePolicy())
    system.setPolicy("contraction", rpol.RepeatPatternContractionPolicy(range(4)))
    system.runUntilConverged()
    out["zfield15d_energy"] = np.array(system.computeOneSiteExpectation())
    out["zfield15d_counts"] = np.array([system.number_of_sweeps, system.number_of_iterations])
    np.savez_compressed(os.path.join(HERE, "runs.npz"), **out)


if __name__ == "__main__":
    dense_recipes()
    data_ops()
    system_walk("tfim", 2, 2, 1, [0, 1, 2, 3])
    system_walk("heis", 1, 2, 2, [0, 1, 2, 3, 0, 2])
    system_walk("tfim", 2, 3, 3, [1, 3, 0])
    relax_case()
    compressor_case()
    bandwidth_case()
    run_case()
    print("golden vectors written to", HERE)
