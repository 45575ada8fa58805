"""NumPy model of the device LU's schedule (csrc/solver.cu: lu_factor, csrc/lu_panel.cu) checked against
scipy.linalg.lu_factor on the CPU: 64-column panels factorised with partial pivoting whose interchanges reach the other
columns only afterwards (zlaswp), inside outer blocks of OB columns; a panel updates the rest of its outer block only, the
outer block's U rows on the right come from a block forward substitution, and the trailing matrix sees ONE product per outer
block.  The model pins the ORDER of operations the kernels implement (what is updated when, which columns an interchange
touches at which point); the kernels themselves are tested against SciPy on the GPU (tests/test_gpu_system.py)."""
import numpy as np
import pytest
import scipy.linalg as sla

NB, SUB = 64, 8


def panel(A, j0, nb, piv):
    """lu_panel_cluster_kernel: sub-panels of 8 columns; interchanges applied inside the sub-panel at once, to the panel
    columns on the right after each sub-panel, never to the columns on the left or outside the panel."""
    n = A.shape[0]
    pe = j0 + nb
    for s0 in range(j0, pe, SUB):
        c_sub = min(s0 + SUB, pe)
        swaps = []
        for j in range(s0, c_sub):
            col = A[j:, j]
            p = j + int(np.argmax(np.abs(col.real) + np.abs(col.imag)))       # LAPACK izamax: |re| + |im|, first on ties
            piv[j] = p
            swaps.append((j, p))
            if p != j:
                A[[j, p], s0:c_sub] = A[[p, j], s0:c_sub]
            A[j + 1:, j] = A[j + 1:, j] * (1.0 / A[j, j])
            A[j + 1:, j + 1:c_sub] -= np.outer(A[j + 1:, j], A[j, j + 1:c_sub])
        if c_sub < pe:
            for j, p in swaps:                                                 # CTA 0: interchanges on the right part
                if p != j:
                    A[[j, p], c_sub:pe] = A[[p, j], c_sub:pe]
            L11 = np.tril(A[s0:c_sub, s0:c_sub], -1) + np.eye(c_sub - s0)
            A[s0:c_sub, c_sub:pe] = sla.solve_triangular(L11, A[s0:c_sub, c_sub:pe], lower=True, unit_diagonal=True)
            A[c_sub:, c_sub:pe] -= A[c_sub:, s0:c_sub] @ A[s0:c_sub, c_sub:pe]


def laswp(A, j0, nb, piv):
    """lu_laswp_kernel: all interchanges on the columns outside the panel, those of LATER sub-panels on a panel column."""
    n = A.shape[0]
    pe = j0 + nb
    for c in range(n):
        ks = SUB * ((c - j0) // SUB + 1) if j0 <= c < pe else 0
        for k in range(ks, nb):
            p = piv[j0 + k]
            if p != j0 + k:
                A[[j0 + k, p], c] = A[[p, j0 + k], c]


def trsm(A, j0, nb, c0, c1):
    L = np.tril(A[j0:j0 + nb, j0:j0 + nb], -1) + np.eye(nb)
    A[j0:j0 + nb, c0:c1] = sla.solve_triangular(L, A[j0:j0 + nb, c0:c1], lower=True, unit_diagonal=True)


def lu_schedule(A, OB):
    n = A.shape[0]
    piv = np.arange(n)
    for o0 in range(0, n, OB):
        o1 = min(n, o0 + OB)
        for j0 in range(o0, o1, NB):
            nb = min(NB, o1 - j0)
            pe = j0 + nb
            panel(A, j0, nb, piv)
            laswp(A, j0, nb, piv)
            if pe < o1:
                trsm(A, j0, nb, pe, o1)
                A[pe:, pe:o1] -= A[pe:, j0:pe] @ A[j0:pe, pe:o1]
        if o1 < n:
            for j0 in range(o0, o1, NB):
                pe = min(j0 + NB, o1)
                trsm(A, j0, pe - j0, o1, n)
                A[pe:o1, o1:] -= A[pe:o1, j0:pe] @ A[j0:pe, o1:]
            A[o1:, o1:] -= A[o1:, o0:o1] @ A[o0:o1, o1:]
    return A, piv


@pytest.mark.parametrize("n, OB", [(5, 256), (64, 256), (65, 128), (130, 128), (200, 64), (333, 256), (333, 128), (520, 512)])
def test_blocked_schedule_matches_scipy(n, OB):
    rng = np.random.default_rng(n + OB)
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    lu, piv = lu_schedule(a.copy(), OB)
    ref_lu, ref_piv = sla.lu_factor(a)
    assert np.array_equal(piv, ref_piv)
    assert np.abs(lu - ref_lu).max() <= 1e-10 * np.abs(ref_lu).max()


def test_interchanges_as_one_gather():
    """lu_perm_build_kernel: replaying "swap(b[j], b[piv[j]])" on an index array gives the gather permutation that
    lu_solve_fast folds into the first substitution's right-hand-side read (csrc/solver.cu)."""
    rng = np.random.default_rng(3)
    for n in (1, 7, 64, 300):
        piv = np.array([rng.integers(j, n) for j in range(n)])
        b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        expect = b.copy()
        for j in range(n):
            expect[[j, piv[j]]] = expect[[piv[j], j]]
        perm = np.arange(n)
        for j in range(n):
            perm[[j, piv[j]]] = perm[[piv[j], j]]
        assert np.array_equal(b[perm], expect)
