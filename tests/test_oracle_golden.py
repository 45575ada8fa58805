"""Pin the CPU oracle (oracle/) to the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU-only; this is what makes the oracle trustworthy as the checker for
the CUDA path."""
import numpy as np
import pytest

from golden_io import load, relerr, sparse, system_parts
from oracle import dense, linalg, solver, tags
from oracle.system import System, stage3_multipliers

TOL = 1e-13


def test_absorb_side_into_corner():
    g = load("dense_recipes")
    assert relerr(dense.absorb_side_into_corner_from_left(g["afl_corner"], g["afl_side"]), g["afl_out"]) < TOL
    assert relerr(dense.absorb_side_into_corner_from_right(g["afr_corner"], g["afr_side"]), g["afr_out"]) < TOL


@pytest.mark.parametrize("i", range(4))
def test_absorb_center_into_side(i):
    g = load("dense_recipes")
    side, center, op = g["ss%d_side" % i], g["ss%d_center" % i], g["ss%d_op" % i]
    assert relerr(dense.absorb_center_ss_into_side(i, side, center), g["ss%d_out" % i]) < TOL
    assert relerr(dense.absorb_center_sos_into_side(i, side, center, op), g["sos%d_out" % i]) < TOL


def test_stages():
    g = load("dense_recipes")
    assert relerr(dense.stage1(g["st1_corner"], g["st1_side"]), g["st1_out"]) < TOL
    assert relerr(dense.stage2(g["st2_a"], g["st2_b"]), g["st2_out"]) < TOL
    s0, s1, op, v = g["st3_s2_0"], g["st3_s2_1"], g["st3_op"], g["st3_v"]
    assert relerr(dense.stage3_multiply(s0, s1, v), g["st3_norm_out"]) < TOL
    assert relerr(dense.stage3_multiply(s0, s1, v, op), g["st3_dense_out"]) < TOL
    assert relerr(dense.stage3_form_matrix(s0, s1, op), g["st3_dense_matrix"]) < TOL
    assert relerr(dense.stage3_form_matrix(s0, s1, np.eye(2)), g["st3_norm_matrix"]) < TOL
    d = 2
    assert dense.stage3_cost_of_multiply(s0.shape, s1.shape, d, False) == g["st3_norm_cost"][0]
    assert dense.stage3_cost_of_multiply(s0.shape, s1.shape, d, True) == g["st3_dense_cost"][0]
    assert dense.stage3_cost_of_form_matrix(s0.shape, s1.shape, d) == g["st3_dense_cost"][1]
    # multiplier == explicit matrix (reference tests/test_system.py:222-259)
    m = dense.stage3_form_matrix(s0, s1, op)
    assert relerr((m @ v.ravel()).reshape(v.shape), g["st3_dense_out"]) < TOL


@pytest.mark.parametrize("n", range(4))
def test_normalize_axis_and_absorb_matrix(n):
    g = load("data_ops")
    t, axis = g["na%d_in" % n], int(g["na%d_axis" % n])
    iso, nrm, den = linalg.normalize_axis(t, axis)
    assert relerr(iso, g["na%d_iso" % n]) < 1e-12
    assert relerr(nrm, g["na%d_nrm" % n]) < 1e-12
    assert relerr(den, g["na%d_den" % n]) < 1e-12
    if t.shape[axis] > 1:
        a, b = linalg.normalize_axis(t, axis, True)
        assert relerr(a, g["na%d_sqrt_nrm" % n]) < 1e-12
        assert relerr(b, g["na%d_sqrt_den" % n]) < 1e-12
    assert relerr(linalg.absorb_matrix_at(t, axis, g["am%d_m" % n]), g["am%d_out" % n]) < TOL


def test_enlargener_and_unitize():
    g = load("data_ops")
    q, qc = linalg.enlargener_from_random(g["enl_sample"])
    assert relerr(q, g["enl_q"]) < TOL
    assert relerr(qc, g["enl_q"].conj()) < TOL
    assert relerr(linalg.unitize(g["unitize_in"]), g["unitize_out"]) < 1e-12


WALKS = ["walk_tfim_chi2_D2", "walk_heis_chi1_D2", "walk_tfim_chi2_D3"]


@pytest.mark.parametrize("name", WALKS)
def test_system_walk(name):
    g = load(name)
    corners, sides, center = system_parts(g, "init")
    op = sparse(g, "operator")
    s = System(corners, sides, center, op)
    for direction in g["moves"]:
        s.contract_unnormalized_towards(int(direction))
    wc, ws, _ = system_parts(g, "walked")
    for i in range(4):
        # same tags, same insertion order, same data
        assert list(s.corners[i]) == list(wc[i])
        assert list(s.sides[i]) == list(ws[i])
        for t in wc[i]:
            assert relerr(s.corners[i][t], wc[i][t]) < TOL, ("corner", i, t)
        for t in ws[i]:
            assert relerr(s.sides[i][t], ws[i][t]) < TOL, ("side", i, t)
    H, N = s.multipliers()
    v = g["v"]
    assert relerr(H(v), g["Hv"]) < 1e-12
    assert relerr(N(v), g["Nv"]) < 1e-12
    assert [H.cost_of_multiply, H.cost_of_form_matrix, N.cost_of_multiply, N.cost_of_form_matrix] == list(g["costs"])
    if "Hmat" in g:
        assert relerr(H.form_matrix(), g["Hmat"]) < 1e-12
        assert relerr(N.form_matrix(), g["Nmat"]) < 1e-12
    e, n = s.expectation_and_normalization()
    assert abs(e - g["expectation"]) <= 1e-12 * abs(g["expectation"])
    assert abs(n - g["normalization"]) <= 1e-12 * abs(g["normalization"])
    s.contract_towards(int(g["moves"][0]))
    assert relerr(s.center, g["after_ct.center"]) < 1e-11
    e, n = s.expectation_and_normalization()
    assert abs(e - g["after_ct.expectation"]) <= 1e-10 * abs(g["after_ct.expectation"])
    assert abs(n - g["after_ct.normalization"]) <= 1e-10 * abs(g["after_ct.normalization"])


def _mult_from_matrix(m, cheap_matrix):
    n = m.shape[0]
    big = 10 ** 12
    return solver.Mult((n, n), lambda v: m @ v, n * n if cheap_matrix is None else (big if cheap_matrix else n * n),
                       lambda: m, 0 if cheap_matrix is None else (0 if cheap_matrix else big))


def test_relax_over():
    g = load("relax")
    h, nm, v0 = g["H"], g["N"], g["v0"]
    exact = np.linalg.eigvalsh(np.linalg.solve(np.linalg.cholesky(nm), np.linalg.solve(np.linalg.cholesky(nm), h).conj().T))[0]
    # GMRES branch for N^-1, operator branch for H
    res = solver.relax_over(v0, _mult_from_matrix(h, False), _mult_from_matrix(nm, False), 100)
    ray = np.vdot(res, h @ res) / np.vdot(res, nm @ res)
    assert abs(ray - g["gmres_rayleigh"]) < 1e-6          # GMRES rtol 1e-5 inside: only loosely comparable
    assert relerr(res, g["gmres_result"]) < 1e-4
    # LU + dense-matrix branches are deterministic: same iterates as the reference
    res = solver.relax_over(v0, _mult_from_matrix(h, True), _mult_from_matrix(nm, True), 100)
    ray = np.vdot(res, h @ res) / np.vdot(res, nm @ res)
    assert relerr(res, g["lu_result"]) < 1e-10
    assert abs(ray - g["lu_rayleigh"]) < 1e-12
    assert ray.real >= exact - 1e-9
    res = solver.relax_over(v0, _mult_from_matrix(h, True), None, 3)
    assert relerr(res, g["one_restart_result"]) < 1e-11


def test_product_compressor():
    g = load("compressor")
    Lt, Rt, new = g["L"], g["R"], int(g["new"])
    assert relerr(solver.product_compressor_matrix(Lt, g["als_c0"], Rt), g["als_matrix"]) < TOL
    c = solver.product_compressor(Lt, Rt, new, initial=g["initial"])
    am = linalg.absorb_matrix_at
    Lc = am(am(Lt, 1, c), 2, c.conj())
    Rc = am(am(Rt, 0, c.conj()), 1, c)
    prod = np.tensordot(Lc, Rc, axes=([1, 2, 3], [0, 1, 2]))
    # gauge-invariant comparisons: the compressed product and the projector c^H c
    assert relerr(prod, g["compressed_product"]) < 1e-8
    assert relerr(prod, g["exact_product"]) < 1e-8
    assert relerr(c.conj().T @ c, g["compressor"].conj().T @ g["compressor"]) < 1e-8


def test_compute_compressor():
    g = load("compressor")
    gram = g["gram"]
    m = solver.Mult((7, 7), lambda v: gram @ v, 49, lambda: gram, 0)
    for new in (5, 2):
        comp, inv = solver.compute_compressor(7, new, m, normalize=True)
        ref = g["cc%d_comp" % new]
        # eigenvalues (= squared row norms with normalize=True) are the gauge-invariant content
        assert np.allclose(np.sort(np.linalg.norm(comp, axis=1) ** 2), np.sort(np.linalg.norm(ref, axis=1) ** 2),
                           rtol=1e-10)
        assert relerr(comp.conj().T @ comp, ref.conj().T @ ref) < 1e-9
        assert relerr(inv.conj().T @ comp, g["cc%d_inv" % new].conj().T @ ref) < 1e-9


def test_increase_bandwidth():
    g = load("bandwidth")
    corners, sides, center = system_parts(g, "before")
    s = System(corners, sides, center, sparse(g, "operator"))
    s.increase_bandwidth(0, by=1, sample=g["sample"])
    ac, as_, acenter = system_parts(g, "after")
    assert s.center.shape == acenter.shape
    assert s.just_increased_bandwidth
    # The enlarged center is rank deficient, so LAPACK's null-space choice makes tensors gauge dependent;
    # the expectation after re-optimising is what must agree.
    s.minimize_expectation()
    e, n = s.expectation_and_normalization()
    assert abs(e - g["after_min.expectation"]) < 1e-5 * abs(g["after_min.expectation"])


def test_stage3_terms_tfim():
    """SURVEY.md section 10 worked example: 9 stage-3 terms after one contraction per direction."""
    g = load("walk_tfim_chi2_D2")
    corners, sides, center = system_parts(g, "walked")
    s = System(corners, sides, center, sparse(g, "operator"))
    H, _ = s.multipliers()
    assert len(H.terms) == 9
