"""carcassonne_b200 -- B200 (sm_100a) implementation of Carcassonne's center-site optimisation hot path.

Same Python API as the reference's ``carcassonne.system`` / ``carcassonne.tensors`` / ``carcassonne.policies``
for that path, with every tensor resident in HBM and every operation running in hand-written CUDA behind the
C ABI of ``libcarc_b200.so`` (``include/carc_b200.h``).  Importing the package loads the shared library and
fails loudly when it has not been built -- there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (loads libcarc_b200.so or raises ImportError)

__version__ = "0.1.0"
