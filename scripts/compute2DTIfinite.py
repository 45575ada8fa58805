#!/usr/bin/env python
"""Exact diagonalisation of the transverse-Ising model on the N x N helical lattice of the reference's
scripts/compute2DTIfinite.py (site i couples to i+1 and i+N modulo N^2; the three expectation lines sum over the
first N sites, as the reference's loop does).  With --device (off by default for this driver: full-2D runs are
experimental, see _drivers.run_tfim_plane) the simulator's infinite square-lattice energy per site is printed next to
E/N^2.

    python scripts/compute2DTIfinite.py N J [--device] [--chi 2] [--max-bandwidth 2]
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
    ap.add_argument("--chi", type=int, default=2, help="boundary bond kept by the state-compression policy")
    ap.add_argument("--max-bandwidth", type=int, default=2, help="largest state bond dimension of the device run")
    args = ap.parse_args()
    energy, exp_z, exp_h, exp_v = drv.tfim_helical_lattice(args.N, args.J)
    print("<Z> =", exp_z)
    print("<XX>_H =", exp_h)
    print("<XX>_V =", exp_v)
    print("E = {:.15f}".format(energy))
    if args.device:
        energies, seconds, bond, sweeps, note = drv.run_tfim_plane(args.J, args.chi, max_bandwidth=args.max_bandwidth,
                                                                   seed=args.seed)
        print("device: E/site (infinite square lattice, chi = {}) by state bond: [{}]  lattice E/N^2 = {:.10f}  "
              "{} sweeps  {:.2f} s  ({})".format(args.chi, ", ".join("%.10f" % e for e in energies), energy / args.N ** 2,
                                               sweeps, seconds, note))


if __name__ == "__main__":
    main()
