#!/usr/bin/env python
"""Free-fermion estimate of the transverse-Ising ring energy (the reference's scripts/estimateTIfinite.py).

    python scripts/estimateTIfinite.py N J [--no-device]

Prints -sum_m sqrt(1 + lam^2 + 2 lam cos(2 pi m / N)), lam = J/2, over the N momenta; with a GPU also the simulator's
infinite-chain energy per site times N for the same Hamiltonian, H = -sum Z - (J/2) sum X X."""
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
    estimate = drv.tfim_ring_energy_estimate(args.N, args.J)
    print(estimate)
    if drv.want_device(args):
        energy, seconds, bond, sweeps = drv.run_tfim_chain(args.J / 2.0, seed=args.seed)
        print("device: N x E/site(infinite chain) = {:.12f}  (ring estimate {:+.2e})  bond dimension {}  {:.2f} s".format(
            args.N * energy, args.N * energy - estimate, bond, seconds))


if __name__ == "__main__":
    main()
