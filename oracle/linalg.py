"""Small linear-algebra pieces of the hot path (normalizeAxis, absorbMatrixAt, unitize, enlargeners).

Oracle (test infrastructure) -- see ``oracle/__init__.py``.  The factorizations themselves live in
SciPy/LAPACK exactly as in the reference (scipy.linalg.svd -> zgesdd, scipy.linalg.qr -> zgeqrf).
"""
import numpy as np
import scipy.linalg as sla


def absorb_matrix_at(t, axis, m):
    """reference data/__init__.py:148-150: out[..., j, ...] = sum_k m[j, k] t[..., k, ...] on `axis`."""
    return np.ascontiguousarray(np.moveaxis(np.tensordot(m, t, axes=([1], [axis])), 0, axis))


def unitize(m):
    """reference utils.py:879-881: polar isometry U V^H of m = U S V^H."""
    u, _, vh = sla.svd(m, full_matrices=False)
    return u @ vh


def normalize_axis(t, axis, sqrt_svals=False, dont_recip_under=1e-14):
    """reference data/__init__.py:263-301 (NDArrayData.normalizeAxis).

    Returns (normalized tensor, normalizer, denormalizer), or (normalizer, denormalizer) when sqrt_svals.
    With M = join(other axes, axis) = U S V^H:  normalized = U V^H (axis restored),
    normalizer = conj(V S^-1 V^H), denormalizer = V S V^H (S^-1 skipped where S <= dont_recip_under).
    """
    if t.shape[axis] == 1:
        n = sla.norm(t)
        if sqrt_svals:
            n = np.sqrt(n)
            return np.array([[1 / n]]), np.array([[n]])
        return t / n, np.array([[1 / n]]), np.array([[n]])
    others = [i for i in range(t.ndim) if i != axis]
    m = np.ascontiguousarray(t.transpose(others + [axis])).reshape(-1, t.shape[axis])
    if m.shape[0] < m.shape[1]:
        raise ValueError("not enough degrees of freedom to normalize axis {}".format(axis))
    u, s, vh = sla.svd(m, full_matrices=False)
    si = s.copy()
    if dont_recip_under:
        nz = np.abs(si) > dont_recip_under
        si[nz] = 1.0 / si[nz]
    else:
        si = 1.0 / si
    if sqrt_svals:
        s = np.sqrt(s)
        si = np.sqrt(si)
        return (vh * si[:, None]).conj(), vh * s[:, None]
    iso = (u @ vh).reshape([t.shape[i] for i in others] + [t.shape[axis]])
    iso = np.ascontiguousarray(np.moveaxis(iso, -1, axis))
    normalizer = vh.T @ (vh * si[:, None]).conj()
    denormalizer = vh.T.conj() @ (vh * s[:, None])
    return iso, normalizer, denormalizer


def normalize_axis_and_denormalize(t, axis_to_norm, axis_to_denorm, other=None):
    """reference data/__init__.py:302-310."""
    if other is None:
        other = t
    iso, _, den = normalize_axis(t, axis_to_norm)
    return iso, absorb_matrix_at(other, axis_to_denorm, den)


def enlargener_from_random(sample):
    """reference data/__init__.py:43-50 (newEnlargener) given the random (new x old) draw: economic QR's Q
    and its conjugate."""
    q, _ = sla.qr(sample, mode="economic")
    return q, q.conj()


def random_complex(rng, *shape):
    """reference utils.py:795-797 (randomComplexSample): uniform [-1,1) + i[-1,1); real parts drawn first."""
    re = rng.random_sample(shape)
    im = rng.random_sample(shape)
    return re * 2 - 1 + im * 2j - 1j
