"""Lane-level CPU model of the folded stage-3 tiling (scripts/stage3f_model.py): the fragment index maps the kernel
relies on -- first product in {2c+e} K order with a trailing single step, C fragment reused as the A fragment of the
second product against B[8 rt + r, 4 j + c], both spins of one S column in one lane -- reproduce the contraction."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


@pytest.mark.parametrize("shape", [(4, 4, 4, 4), (9, 9, 9, 9), (25, 25, 25, 25), (36, 36, 36, 36), (6, 6, 9, 4),
                                   (16, 16, 4, 4), (1, 1, 1, 1), (5, 13, 3, 7), (8, 4, 8, 12)])
def test_fragment_maps(shape):
    import stage3f_model
    rng = np.random.default_rng(sum(shape))
    assert stage3f_model.model(*shape, rng) < 1e-13
