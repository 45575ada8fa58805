"""The einsum restatements that tests/test_gpu_1d.py checks the device's 1D (MPS/MPO) recipes against, pinned to the
reference's own outputs (tests/golden/recipes_1d.npz, generator tests/golden/make_1d_golden.py), and the product's
makeMPO (host code) against the reference's.  CPU only: reference -> golden -> einsum (here) -> device (GPU suite)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "recipes_1d.npz")
X = np.array([[0, 1], [1, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
I2 = np.eye(2, dtype=complex)


def relerr(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_einsum_restatements_match_the_reference():
    g = np.load(GOLDEN)
    L, R, O, S, Sr = g["L"], g["R"], g["O"], g["S"], g["Sr"]
    assert relerr(np.einsum("ost,uoqp,asp,btq->uab", L, O, S, S.conj()), g["oss_left"]) < 1e-14
    assert relerr(np.einsum("ost,ouqp,sap,tbq->uab", R, O, Sr, Sr.conj()), g["oss_right"]) < 1e-14
    assert relerr(np.einsum("st,asp,btp->ab", g["L2"], S, S.conj()), g["ss_left"]) < 1e-14
    assert relerr(np.einsum("st,sap,tbp->ab", g["R2"], Sr, Sr.conj()), g["ss_right"]) < 1e-14
    ref = np.einsum("osa,utb,ouqp,stp->abq", g["Rm"], g["Lm"], O, g["Sc"])
    assert relerr(ref, g["mult_out"]) < 1e-14
    assert relerr((g["mult_matrix"] @ g["Sc"].ravel()).reshape(g["Sc"].shape), ref) < 1e-14


def test_make_mpo_matches_the_reference():
    from carcassonne_b200.sparse import Complete, Identity, TwoSiteOperator, makeMPO
    g = np.load(GOLDEN)
    tensor, right, right_tags, left, left_tags = makeMPO(I2, Os=[-Z], OOs=[(X, -0.7 * X)])
    assert np.array_equal(tensor, g["mpo1_tensor"])
    assert list(right) == list(g["mpo1_right"]) and list(left) == list(g["mpo1_left"])
    assert right_tags == [Identity(), TwoSiteOperator(0, 2), Complete()]
    assert left_tags == [Complete(), TwoSiteOperator(0, 2), Identity()]
    tensor, right, _, left, _ = makeMPO(I2, Os=[], OOs=[(X, X), (Z, -Z)])
    assert np.array_equal(tensor, g["mpo2_tensor"])
    assert list(right) == list(g["mpo2_right"]) and list(left) == list(g["mpo2_left"])
