"""Per-minimisation trace (sweep, iteration, center shape, one-site energy, host RNG position) of selected end-to-end runs
on the device, to line up against the same trace of the unmodified reference when an iteration count or an energy differs."""
import os
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carcassonne_b200 import policies as pol  # noqa: E402
from carcassonne_b200.data import DeviceData as dd, _init_constants  # noqa: E402
from carcassonne_b200.system import System  # noqa: E402

_init_constants()
One = pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy


def trace(system, label):
    orig = system.minimizeExpectation

    def traced(*a, **k):
        st = {}
        orig(statistics=st)
        print(label, "sweep", system.number_of_sweeps, "iter", system.number_of_iterations, "shape",
              system.state_center_data.shape, "E1", complex(system.computeOneSiteExpectation()), "rng",
              np.random.get_state()[2], "mults", st.get("multiplications"), st.get("normalization"), flush=True)
    system.minimizeExpectation = traced


for d in (0, 1):
    np.random.seed(320 + d)
    random.seed(320 + d)
    kw = {"OO_LR" if d == 0 else "OO_UD": [dd.X, -0.01 * dd.X]}
    system = System.newTrivialWithSimpleSparseOperator(O=-dd.Z, **kw)
    system.setPolicy("state compression", pol.ConstantStateCompressionPolicy(1))
    system.setPolicy("sweep convergence", One(1e-7))
    system.setPolicy("run convergence", One(1e-7))
    system.setPolicy("bandwidth increase", pol.OneDirectionIncrementBandwidthIncreasePolicy(d))
    system.setPolicy("contraction", pol.RepeatPatternContractionPolicy(range(4)))
    trace(system, "15d.tfim.dir%d" % d)
    system.runUntilConverged()
    print("final", system.computeOneSiteExpectation(), system.number_of_sweeps, system.number_of_iterations, flush=True)
