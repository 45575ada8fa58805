#!/usr/bin/env python
"""Reference-generated fixture for a genuinely LOSSY state-bond compression ("lossy compression is
unpinned against the reference").  Imports the UNMODIFIED reference from /root/reference (build container only) and
writes tests/golden/compressor_lossy.npz.

For each case: random Hermitian-symmetrised L [l, old, old, 1] and R [old, old, 1, r] (the layouts
compressCornerStateTowards hands over, system/_2d.py:199-228), the reference's computeProductCompressor (compression.py:
26-45: random start, 4 ALS rounds of GMRES on the normal equations, polar projection) run after a seed, the draw it made,
its compressor and the relative error of the compressed product -- the only quality measure the reference's construction
defines.  Cases where the reference's own `assert info == 0` fires (GMRES did not reach rtol 1e-5) are skipped and the
next seed is tried; what was skipped is recorded.
"""
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("CARCASSONNE_REFERENCE", "/root/reference"))
HERE = os.path.dirname(os.path.abspath(__file__))

from carcassonne.compression import computeProductCompressor  # noqa: E402
from carcassonne.data import NDArrayData as ND  # noqa: E402


def product_error(Lt, Rt, c):
    L2 = ND(Lt).absorbMatrixAt(1, c).absorbMatrixAt(2, c.conj())
    R2 = ND(Rt).absorbMatrixAt(0, c.conj()).absorbMatrixAt(1, c)
    exact = ND(Lt).contractWith(ND(Rt), (1, 2, 3), (0, 1, 2)).toArray()
    got = L2.contractWith(R2, (1, 2, 3), (0, 1, 2)).toArray()
    return float(np.linalg.norm(got - exact) / np.linalg.norm(exact))


def main():
    out, skipped, found = {}, [], 0
    # (l, old, new, r, noise): noise = None -> fully random tensors (barely compressible); otherwise a product that
    # lives on a `new`-dimensional subspace plus `noise` x random (the regime a converging run is in)
    shapes = [(6, 6, 3, 20, None), (5, 8, 4, 12, None), (9, 6, 2, 9, None), (6, 6, 3, 20, 0.05), (7, 8, 3, 10, 0.2),
              (4, 9, 4, 16, 0.01)]
    for (l, old, new, r, noise) in shapes:
        for seed in range(40):
            np.random.seed(1000 + seed)
            Lt = ND.newRandom(l, old, old, 1)
            Lt += Lt.transpose(0, 2, 1, 3).conj()
            Rt = ND.newRandom(old, old, 1, r)
            Rt += Rt.transpose(1, 0, 2, 3).conj()
            Lt, Rt = Lt.toArray(), Rt.toArray()
            if noise is not None:
                mask = np.zeros((old, old))
                mask[:new, :new] = 1.0
                Lt = Lt * (mask + noise * (1 - mask))[None, :, :, None]
                Rt = Rt * (mask + noise * (1 - mask))[:, :, None, None]
            state = np.random.get_state()
            try:
                c = computeProductCompressor(ND(Lt), ND(Rt), new)
            except AssertionError:
                skipped.append((l, old, new, r, seed))
                continue
            np.random.set_state(state)
            initial = ND.newRandom(old, new).toArray()      # the draw compression.py:35 made
            k = "case%d." % found
            out[k + "L"], out[k + "R"], out[k + "new"] = Lt, Rt, np.array(new)
            out[k + "initial"], out[k + "compressor"] = initial, c.toArray()
            out[k + "product_error"] = np.array(product_error(Lt, Rt, c))
            print("shape", (l, old, new, r, noise), "seed", seed, "reference product error", out[k + "product_error"])
            found += 1
            break
    out["cases"] = np.array(found)
    out["skipped"] = np.array(skipped if skipped else np.zeros((0, 5)), dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "compressor_lossy.npz"), **out)
    print("cases:", found, "skipped (reference GMRES assert):", len(skipped))


if __name__ == "__main__":
    main()
