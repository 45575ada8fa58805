import faulthandler, os, sys, random
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, ".")
faulthandler.dump_traceback_later(60, exit=True)
import numpy as np
import logging
logging.basicConfig(level=logging.DEBUG, stream=sys.stdout)
from carcassonne_b200.data import DeviceData as dd, _init_constants
_init_constants()
from carcassonne_b200 import policies as pol
from carcassonne_b200.system import System
direction = int(sys.argv[1]) if len(sys.argv) > 1 else 0
np.random.seed(51 + direction); random.seed(51 + direction)
kw = {"OO_LR" if direction == 0 else "OO_UD": [dd.X, -0.01 * dd.X]}
system = System.newTrivialWithSimpleSparseOperator(O=-dd.Z, **kw)
system.setPolicy("sweep convergence", pol.RelativeStateDifferenceThresholdConvergencePolicy(1e-5))
system.setPolicy("run convergence", pol.RelativeOneSiteExpectationDifferenceThresholdConvergencePolicy(1e-7))
system.setPolicy("bandwidth increase", pol.OneDirectionIncrementBandwidthIncreasePolicy(direction, 2))
system.setPolicy("contraction", pol.RepeatPatternContractionPolicy([0 + direction, 2 + direction]))
orig = system.minimizeExpectation
def traced():
    st = {}
    print("minimize: center", system.state_center_data.shape, "sides", [s[list(s)[0]].shape for s in system.sides], flush=True)
    orig(statistics=st)
    print("  ->", st, flush=True)
system.minimizeExpectation = traced
system.runUntilConverged()
print("energy", system.computeOneSiteExpectation(), system.number_of_sweeps, system.number_of_iterations, system.state_center_data.shape)
