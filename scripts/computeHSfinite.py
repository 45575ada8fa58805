#!/usr/bin/env python
"""Exact diagonalisation of the Haldane-Shastry ring sum_{m,n} (X X + Y Y + Z Z)_{m,m+n} / (2 sin^2(n pi / N)) (the
reference's scripts/computeHSfinite.py, which prints E / N^3 under the label E/N), and the simulator's nearest-neighbour
Heisenberg chain -- the model whose sparse operator the library ships (six two-site terms) -- next to 1 - 4 ln 2.

    python scripts/computeHSfinite.py N [--no-device]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import _drivers as drv  # noqa: E402


def main():
    ap = drv.add_device_arguments(drv.parser(__doc__))
    ap.add_argument("N", type=int)
    args = ap.parse_args()
    energy = drv.haldane_shastry_ring(args.N)
    print("E/N = {:.15f}".format(energy / args.N ** 3))
    if drv.want_device(args):
        e, seconds, bond = drv.run_heisenberg_chain(seed=args.seed)
        exact = 1.0 - 4.0 * np.log(2.0)
        print("device: nearest-neighbour Heisenberg chain E/site = {:.8f}  (Bethe ansatz {:.8f}, {:+.1e})  "
              "bond dimension {}  {:.2f} s".format(e, exact, e - exact, bond, seconds))


if __name__ == "__main__":
    main()
