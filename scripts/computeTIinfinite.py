#!/usr/bin/env python
"""Infinite transverse-Ising chain: exact energy per site, and the simulator's on the device next to it.

    python scripts/computeTIinfinite.py J [--no-device]

Prints the two lines of the reference's scripts/computeTIinfinite.py -- `E(m)/(pi/2)  1+lam` and the energy per site of
H = -sum Z - (J/2) sum X X (its lam = J/2 convention) -- then runs that chain through carcassonne_b200's 2D system
driven along one axis."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import _drivers as drv  # noqa: E402


def main():
    ap = drv.add_device_arguments(drv.parser(__doc__))
    ap.add_argument("J", type=float)
    args = ap.parse_args()
    from scipy.special import ellipe
    lam = args.J / 2.0
    print(ellipe(4.0 * lam / (1.0 + lam) ** 2) / (np.pi / 2.0), 1.0 + lam)
    exact = drv.tfim_infinite_chain_energy(args.J)
    print("{:.15f}".format(exact))
    if drv.want_device(args):
        energy, seconds, bond, sweeps = drv.run_tfim_chain(args.J / 2.0, seed=args.seed)
        print("device: E/site = {:.12f}  (exact {:+.2e})  bond dimension {}  {} sweeps  {:.2f} s".format(
            energy, energy - exact, bond, sweeps, seconds))


if __name__ == "__main__":
    main()
