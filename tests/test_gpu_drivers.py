"""Device halves of the compute*.py drivers (scripts/_drivers.py): the simulator's energies next to the exact numbers
the reference's scripts/ compute."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_transverse_ising_chain_against_the_closed_form():
    import _drivers as drv
    for coupling in (0.01, 0.25):
        energy, seconds, bond, sweeps = drv.run_tfim_chain(coupling, seed=0)
        exact = drv.tfim_infinite_chain_energy(2.0 * coupling)     # the closed form's J is twice the coupling
        assert abs(energy - exact) < 2e-6, (coupling, energy, exact)
        assert bond >= 1 and sweeps >= 1


def test_heisenberg_chain_against_bethe_ansatz():
    import _drivers as drv
    energy, seconds, bond = drv.run_heisenberg_chain(seed=0)
    assert abs(energy / 4 - (0.25 - np.log(2.0))) < 2e-3     # the reference's own tolerance (test_simulator_1d.py:170)


def test_driver_command_line_with_device():
    import subprocess
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "computeTIinfinite.py"), "0.02", "--device"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = res.stdout.strip().splitlines()
    assert lines[-1].startswith("device: E/site = ")
    assert abs(float(lines[-1].split()[3]) - float(lines[1])) < 2e-6
