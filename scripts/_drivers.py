"""Shared pieces of the compute*.py / estimate*.py drivers: exact reference values (closed forms, bit-operation exact
diagonalisation) and the device runs of the corresponding infinite-lattice models through carcassonne_b200.

The reference's scripts/ directory (scripts/computeTIfinite.py, computeTIinfinite.py, compute2DTIfinite.py,
computeHSfinite.py, estimateTIfinite.py) holds stand-alone calculators of the exact numbers its author compares the
simulator against; the drivers here keep their names, command lines and printed lines, compute the same numbers
with their own implementation, and -- what the reference leaves to the reader -- run the simulator itself on the
device next to them (`--device`, the default when a GPU is present; `--no-device` prints the exact part only)."""
import argparse
import random
import time

import numpy as np

X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)


# ---- exact values -----------------------------------------------------------------------------------------------
def tfim_infinite_chain_energy(J, Gamma=1.0):
    """Ground-state energy per site of H = -Gamma sum Z - J sum X X on the infinite chain (free-fermion solution):
    -(2 Gamma / pi) (1 + lam) E(m), lam = J / (2 Gamma), m = 4 lam / (1 + lam)^2, E the complete elliptic integral of
    the second kind in the parameter convention of scipy.special.ellipe."""
    from scipy.special import ellipe
    lam = J / (2.0 * Gamma)
    m = 4.0 * lam / (1.0 + lam) ** 2
    return -Gamma * 2.0 / np.pi * (1.0 + lam) * ellipe(m)


def tfim_ring_energy_estimate(N, J):
    """Free-fermion sum over the N momenta of the ring (the estimate estimateTIfinite.py prints)."""
    lam = J / 2.0
    m = np.arange(N) - N // 2 if N % 2 == 0 else np.arange(N) - (N - 1) // 2
    return -float(np.sum(np.sqrt(1.0 + lam * lam + 2.0 * lam * np.cos(2.0 * np.pi * m / N))))


class SpinHalfLattice:
    """States of n spins as bit strings (site 0 = most significant bit, as numpy's reshape((2,) * n) orders them).
    Z_i is diagonal, X_i X_j flips two bits: every Hamiltonian term is an index permutation with a sign."""

    def __init__(self, n):
        self.n = n
        self.index = np.arange(1 << n, dtype=np.int64)

    def mask(self, i):
        return 1 << (self.n - 1 - i)

    def z(self, i):
        return 1.0 - 2.0 * ((self.index & self.mask(i)) != 0)

    def apply_xx(self, i, j, v):
        return v[self.index ^ (self.mask(i) | self.mask(j))]

    def apply_yy(self, i, j, v):
        # Y_i Y_j = -(z_i z_j) X_i X_j acting on the flipped state's source
        return -(self.z(i) * self.z(j)) * self.apply_xx(i, j, v)

    def ground_state(self, matvec, dtype=float):
        from scipy.sparse.linalg import LinearOperator, eigsh
        op = LinearOperator(shape=(1 << self.n,) * 2, matvec=matvec, dtype=dtype)
        evals, evecs = eigsh(op, k=1, which="SA")
        return float(evals[0].real), evecs[:, 0]


def tfim_ring(N, J):
    """-sum Z_i - J sum X_i X_{i+1} on an N-site ring: (E, <sum -Z>, <sum -J X X>)."""
    lat = SpinHalfLattice(N)
    zsum = sum(lat.z(i) for i in range(N))

    def matvec(v):
        out = -zsum * v
        for i in range(N):
            out = out - J * lat.apply_xx(i, (i + 1) % N, v)
        return out

    e, v = lat.ground_state(matvec)
    exp_z = float(np.vdot(v, -zsum * v).real)
    exp_xx = float(sum(np.vdot(v, -J * lat.apply_xx(i, (i + 1) % N, v)).real for i in range(N)))
    return e, exp_z, exp_xx


def tfim_helical_lattice(N, J):
    """The N x N lattice of the reference's compute2DTIfinite.py: site i couples to (i+1) mod N^2 and (i+N) mod N^2
    (a helical, not toroidal, boundary -- SURVEY.md section 9).  Returns E and the three expectation sums over the
    FIRST N sites, which is what the reference's loop (range(N)) prints."""
    n = N * N
    lat = SpinHalfLattice(n)
    zsum = sum(lat.z(i) for i in range(n))

    def matvec(v):
        out = -zsum * v
        for i in range(n):
            out = out - J * lat.apply_xx(i, (i + 1) % n, v) - J * lat.apply_xx(i, (i + N) % n, v)
        return out

    e, v = lat.ground_state(matvec)
    exp_z = float(sum(np.vdot(v, -lat.z(i) * v).real for i in range(N)))
    exp_h = float(sum(np.vdot(v, -J * lat.apply_xx(i, (i + 1) % n, v)).real for i in range(N)))
    exp_v = float(sum(np.vdot(v, -J * lat.apply_xx(i, (i + N) % n, v)).real for i in range(N)))
    return e, exp_z, exp_h, exp_v


def haldane_shastry_ring(N):
    """sum_m sum_{n=1}^{N-1} (X X + Y Y + Z Z)_{m, m+n} / (2 sin^2(n pi / N)): ground-state energy."""
    lat = SpinHalfLattice(N)
    c = [0.0] + [0.5 / np.sin(n * np.pi / N) ** 2 for n in range(1, N)]
    diag = np.zeros(1 << N)
    for m in range(N):
        for n in range(1, N):
            diag += c[n] * lat.z(m) * lat.z((m + n) % N)

    def matvec(v):
        out = diag * v
        for m in range(N):
            for n in range(1, N):
                j = (m + n) % N
                flipped = lat.apply_xx(m, j, v)
                out = out + c[n] * (1.0 - lat.z(m) * lat.z(j)) * flipped    # X X + Y Y: 2 on anti-parallel pairs
        return out

    e, _ = lat.ground_state(matvec)
    return e


# ---- device runs ------------------------------------------------------------------------------------------------
def device_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def add_device_arguments(ap):
    ap.add_argument("--device", dest="device", action="store_true", default=None,
                    help="also run the simulator on the GPU (default when one is present)")
    ap.add_argument("--no-device", dest="device", action="store_false")
    ap.add_argument("--seed", type=int, default=0)
    return ap


def want_device(args):
    return device_available() if args.device is None else args.device


def _seed(seed):
    np.random.seed(seed)
    random.seed(seed)


def _bond_capped(policy_class, max_bond):
    """The run-convergence policy with a ceiling on the state bond: near the critical coupling the energy keeps creeping
    with every bandwidth increase and an uncapped run does not end in any useful time."""
    class Capped(policy_class):
        def converged(self):
            done = policy_class.converged(self)
            if max_bond and max(self.system.state_center_data.shape[:4]) >= max_bond and self.last is not None:
                return True
            return done
    return Capped


def run_tfim_chain(J, seed=0, sweep_tol=1e-5, run_tol=1e-7, increment=2, counts=None, max_bond=None,
                   max_iterations_per_sweep=None):
    """Infinite transverse-Ising chain through the 2D system driven along one axis (reference
    tests/test_simulator_2d_in_1d.py:36-47 with the coupling as a parameter).  Returns (energy per site, seconds,
    final bond dimension, sweeps)."""
    from carcassonne_b200 import policies as pol
    from carcassonne_b200.data import DeviceData, _init_constants
    from carcassonne_b200.system import System
    _init_constants()
    _seed(seed)
    system = System.newTrivialWithSimpleSparseOperator(O=-DeviceData.Z, OO_LR=[DeviceData.X, -J * DeviceData.X])
    sweep = pol.RelativeStateDifferenceThresholdConvergencePolicy(sweep_tol)
    if max_iterations_per_sweep:
        sweep = pol.BoundedConvergencePolicy(sweep, max_iterations_per_sweep)
    system.setPolicy("sweep convergence", sweep)
    system.setPolicy("run convergence",
                     _bond_capped(pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy, max_bond)(run_tol))
    system.setPolicy("bandwidth increase", pol.OneDirectionIncrementBandwidthIncreasePolicy(0, increment))
    system.setPolicy("contraction", pol.RepeatPatternContractionPolicy([0, 2]))
    t0 = time.perf_counter()
    system.runUntilConverged()
    energy = complex(system.computeOneSiteExpectation())
    if counts is not None:
        counts.update(sweeps=system.number_of_sweeps, iterations=system.number_of_iterations)
    return energy.real, time.perf_counter() - t0, system.state_center_data.shape[0], system.number_of_sweeps


def run_tfim_plane(J, chi, max_bandwidth=2, seed=0, tol=1e-6, max_iterations_per_sweep=40):
    """Infinite square-lattice transverse-Ising model: all four directions absorbed in turn, boundary bond compressed
    back to chi after every absorption, state bond grown until the one-site energy settles (or max_bandwidth).

    EXPERIMENTAL, as in the reference: its run loop never renormalises the environment, and full-2D runs lose their norm
    after a few dozen absorptions (SURVEY.md section 9 records the same for the reference itself), after which every
    convergence test compares NaNs and the loop never returns.  Each sweep here is therefore bounded
    (policies.BoundedConvergencePolicy) and a non-finite energy ends the run with the energies obtained so far.
    Returns (energies by state bond, seconds, final state bond, sweeps, note)."""
    from carcassonne_b200 import policies as pol
    from carcassonne_b200.data import DeviceData, _init_constants
    from carcassonne_b200.system import System
    _init_constants()
    _seed(seed)
    system = System.newTrivialWithSimpleSparseOperator(O=-DeviceData.Z, OO_LR=[DeviceData.X, -J * DeviceData.X],
                                                       OO_UD=[DeviceData.X, -J * DeviceData.X])
    bounded = pol.BoundedConvergencePolicy(pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(tol),
                                           max_iterations_per_sweep)
    system.setPolicy("state compression", pol.ConstantStateCompressionPolicy(chi))
    system.setPolicy("sweep convergence", bounded)
    system.setPolicy("bandwidth increase", pol.AllDirectionsIncrementBandwidthIncreasePolicy())
    system.setPolicy("contraction", pol.RepeatPatternContractionPolicy(range(4)))
    t0 = time.perf_counter()
    energies, note = [], "converged"
    try:
        while True:
            system.sweepUntilConverged()
            energy = complex(system.computeOneSiteExpectation()).real
            if not np.isfinite(energy):
                raise FloatingPointError("one-site energy is not finite")
            energies.append(energy)
            if len(energies) > 1 and abs(energies[-1] - energies[-2]) <= tol * abs(energies[-1]):
                break
            if system.state_center_data.shape[0] >= max_bandwidth:
                note = "stopped at the largest state bond asked for"
                break
            system._applyPolicy("bandwidth increase")
    except FloatingPointError as error:
        note = "stopped: %s (the environment's norm left the floating-point range)" % error
    except Exception as error:      # a solver that refuses non-finite input, a failed relaxation policy, ...
        note = "stopped: %s: %s" % (type(error).__name__, error)
    return energies, time.perf_counter() - t0, system.state_center_data.shape[0], system.number_of_sweeps, note


def run_heisenberg_chain(seed=0, sweep_tol=1e-5, run_tol=1e-3):
    """Infinite nearest-neighbour Heisenberg chain through the 1D (MPS/MPO) system on device (reference
    tests/test_simulator_1d.py:146-170).  Returns (energy per site of sum sigma.sigma, seconds, bond dimension)."""
    from carcassonne_b200 import policies as pol
    from carcassonne_b200.data import _init_constants
    from carcassonne_b200.system._1d import System as System1D
    _init_constants()
    _seed(seed)
    I2 = np.eye(2, dtype=complex)
    tensor = np.zeros((5, 5, 2, 2), dtype=complex)
    tensor[0, 0] = I2
    tensor[0, 1], tensor[0, 2], tensor[0, 3] = X, Y, Z
    tensor[1, 4], tensor[2, 4], tensor[3, 4] = -X, -Y, Z      # the reference's sub-lattice-rotated form
    tensor[4, 4] = I2
    system = System1D([1, 0, 0, 0, 0], [0, 0, 0, 0, 1], tensor, np.ones((1, 1, 2)))
    system.setPolicy("sweep convergence", pol.RelativeEstimatedOneSiteExpectationDifferenceThresholdConvergencePolicy(sweep_tol))
    system.setPolicy("run convergence", pol.RelativeEstimatedOneSiteExpectationDifferenceThresholdConvergencePolicy(run_tol))
    system.setPolicy("bandwidth increase", pol.OneDirectionIncrementBandwidthIncreasePolicy(0, 2))
    system.setPolicy("contraction", pol.RepeatPatternContractionPolicy([0, 1]))
    t0 = time.perf_counter()
    system.runUntilConverged()
    energy = complex(system.computeEstimatedOneSiteExpectation()).real
    return energy, time.perf_counter() - t0, system.state_center_data.shape[0]


def parser(description):
    return argparse.ArgumentParser(description=description)
