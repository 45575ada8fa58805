"""CPU oracle for the Carcassonne center-site optimisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``carcassonne_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or
the CPU baseline -- never as the thing measured or shipped.

It is a NumPy/SciPy restatement of the reference's algorithm (gcross/Carcassonne,
pure Python) written as explicit ``einsum`` index strings instead of the
reference's generated pairwise ``tensordot`` + ``join`` code.  Every function
cites the reference ``file:line`` it follows (paths relative to the reference
checkout).  The arithmetic the reference delegates to third-party libraries is
delegated to the same libraries here: NumPy (``tensordot``/``einsum`` ->
OpenBLAS zgemm) and ``scipy.linalg.{svd, eig, eigh, lu_factor, lu_solve, qr}``,
``scipy.sparse.linalg.{gmres, eigsh}``; the reference pins no versions (it has
no setup.py / requirements); this oracle was pinned with numpy 2.3.5 /
scipy 1.18.1.

Parity status: PINNED.  ``tests/golden/*.npz`` hold input/output vectors
produced by importing the unmodified reference from ``/root/reference`` in the
build container (generator: ``tests/golden/make_golden.py``); the CPU test
suite (``tests/test_oracle_golden.py``) checks every oracle function against
them.  Iterative pieces whose results depend on random draws or on LAPACK's
choice of null-space vectors (``relaxOver``, ``computeProductCompressor``,
rank-deficient ``normalizeAxis``) are pinned on gauge-invariant quantities
(Rayleigh quotients, compressed products, projectors), as SURVEY.md section 8c
prescribes.
"""
from . import dense, tags, linalg, solver, system  # noqa: F401
