#!/usr/bin/env python
"""Exact diagonalisation of the transverse-Ising ring H = -sum Z_i - J sum X_i X_{i+1} (the reference's
scripts/computeTIfinite.py), and the simulator's infinite-chain energy per site next to E/N.

    python scripts/computeTIfinite.py N J [--no-device]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _drivers as drv  # noqa: E402


def main():
    ap = drv.add_device_arguments(drv.parser(__doc__))
    ap.add_argument("N", type=int)
    ap.add_argument("J", type=float)
    args = ap.parse_args()
    energy, exp_z, exp_xx = drv.tfim_ring(args.N, args.J)
    print("<Z> =", exp_z)
    print("<XX> =", exp_xx)
    print("E = {:.15f}".format(energy))
    if drv.want_device(args):
        e, seconds, bond, sweeps = drv.run_tfim_chain(args.J, seed=args.seed)
        print("device: E/site (infinite chain) = {:.12f}  ring E/N = {:.12f}  infinite-chain exact = {:.12f}  "
              "bond dimension {}  {:.2f} s".format(e, energy / args.N, drv.tfim_infinite_chain_energy(2.0 * args.J), bond,
                                                  seconds))


if __name__ == "__main__":
    main()
