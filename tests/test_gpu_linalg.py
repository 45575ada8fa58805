"""GPU parity tests of the device factorisations (Householder QR, Jacobi SVD, normalizeAxis, unitize,
newEnlargener) against the reference's golden vectors and SciPy.  Gauge-dependent outputs (U, Vh) are compared
through gauge-invariant quantities."""
import numpy as np
import pytest
import scipy.linalg as sla

from golden_io import load, relerr

pytestmark = pytest.mark.gpu


def crand(rng, *shape):
    return rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)


@pytest.fixture(scope="module")
def dd():
    from carcassonne_b200.data import DeviceData
    return DeviceData


@pytest.mark.parametrize("m,n", [(1, 1), (5, 3), (8, 8), (64, 7), (1024, 16), (4096, 64), (130, 80)])
def test_qr_matches_lapack(dd, m, n):
    rng = np.random.default_rng(m + n)
    a = crand(rng, m, n)
    q, r = dd.fromArray(a).qr(mode="economic")
    q, r = q.toArray(), r.toArray()
    qs, rs = sla.qr(a, mode="economic")
    assert relerr(q @ r, a) < 1e-13
    assert np.linalg.norm(q.conj().T @ q - np.eye(n)) < 1e-12
    # same Householder conventions as LAPACK -> the factors themselves agree
    assert relerr(q, qs) < 1e-10
    assert relerr(r, rs) < 1e-10


def test_enlargener_golden(dd):
    g = load("data_ops")
    q = dd.fromArray(g["enl_sample"]).qr(mode="economic")[0].toArray()
    assert relerr(q, g["enl_q"]) < 1e-12


@pytest.mark.parametrize("m,n", [(3, 3), (6, 3), (3, 6), (40, 40), (512, 16), (200, 80)])
def test_svd(dd, m, n):
    rng = np.random.default_rng(10 * m + n)
    a = crand(rng, m, n)
    u, s, vh = dd.fromArray(a).svd(full_matrices=False)
    u, s, vh = u.toArray(), s.toArray(), vh.toArray()
    k = min(m, n)
    assert u.shape == (m, k) and vh.shape == (k, n)
    ref = sla.svd(a, compute_uv=False)
    assert np.max(np.abs(s.real - ref)) < 1e-12 * ref[0]
    assert np.max(np.abs(s.imag)) == 0
    assert relerr((u * s) @ vh, a) < 1e-12
    assert np.linalg.norm(u.conj().T @ u - np.eye(k)) < 1e-11
    assert np.linalg.norm(vh @ vh.conj().T - np.eye(k)) < 1e-11


def test_unitize_golden(dd):
    g = load("data_ops")
    assert relerr(dd.fromArray(g["unitize_in"]).unitize().toArray(), g["unitize_out"]) < 1e-11


@pytest.mark.parametrize("n", range(4))
def test_normalize_axis_golden(dd, n):
    g = load("data_ops")
    t, axis = g["na%d_in" % n], int(g["na%d_axis" % n])
    iso, nrm, den = dd.fromArray(t).normalizeAxis(axis)
    assert relerr(iso.toArray(), g["na%d_iso" % n]) < 1e-11
    assert relerr(nrm.toArray(), g["na%d_nrm" % n]) < 1e-11
    assert relerr(den.toArray(), g["na%d_den" % n]) < 1e-11
    if t.shape[axis] > 1:
        a, b = dd.fromArray(t).normalizeAxis(axis, True)
        # the sqrt variants carry the gauge of Vh: compare the gauge-invariant products a^T conj(a), b^H b
        ra, rb = g["na%d_sqrt_nrm" % n], g["na%d_sqrt_den" % n]
        a, b = a.toArray(), b.toArray()
        assert relerr(a.T @ a.conj(), ra.T @ ra.conj()) < 1e-11
        assert relerr(b.conj().T @ b, rb.conj().T @ rb) < 1e-11


@pytest.mark.parametrize("shape,axis", [((8, 8, 8, 8, 2), 0), ((8, 8, 8, 8, 2), 2), ((4, 3, 5, 2, 2), 3), ((6, 6, 6, 6, 2), 1)])
def test_normalize_axis_properties(dd, shape, axis):
    """reference tests/test_data.py:98-132: the normalised tensor is an isometry on `axis`, and absorbing the
    denormaliser restores the input."""
    from oracle import linalg
    rng = np.random.default_rng(sum(shape) + axis)
    t = crand(rng, *shape)
    iso, nrm, den = dd.fromArray(t).normalizeAxis(axis)
    ref_iso, ref_nrm, ref_den = linalg.normalize_axis(t, axis)
    assert relerr(iso.toArray(), ref_iso) < 1e-10
    assert relerr(nrm.toArray(), ref_nrm) < 1e-10
    assert relerr(den.toArray(), ref_den) < 1e-10
    others = [i for i in range(len(shape)) if i != axis]
    gram = np.tensordot(iso.toArray().conj(), iso.toArray(), (others, others))
    assert np.linalg.norm(gram - np.eye(shape[axis])) < 1e-11
    back = iso.absorbMatrixAt(axis, den.transpose()).toArray()
    assert relerr(back, t) < 1e-11


def test_normalize_axis_rank_deficient(dd):
    """A rank-deficient axis (what increaseBandwidth produces): the projector onto the range must agree."""
    from oracle import linalg
    rng = np.random.default_rng(77)
    t = crand(rng, 3, 4, 2)
    pad = np.zeros((3, 4, 4), dtype=np.complex128)
    pad[:, :, :2] = t
    iso, nrm, den = dd.fromArray(pad).normalizeAxis(2)
    ref_iso, ref_nrm, ref_den = linalg.normalize_axis(pad, 2)
    assert relerr(den.toArray(), ref_den) < 1e-10
    assert relerr(iso.absorbMatrixAt(2, den.transpose()).toArray(), pad) < 1e-10
    m = iso.toArray().reshape(-1, 4)
    assert np.linalg.norm(m.conj().T @ m - np.eye(4)) < 1e-10
