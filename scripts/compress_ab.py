"""Time the eight corner compressions of a D = 3, chi = 6 sweep iteration (CARC_B200_LIB selects the library build)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carcassonne_b200 import synthetic
ts = []
for rep in range(6):
    system = synthetic.device_system(6, 3, seed=0)
    np.random.seed(0)
    tot = 0.0
    for direction in (0, 1, 2, 3):
        system.minimizeExpectation()
        system.contractTowards(direction)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for corner_id in range(4):
            for d2 in range(2):
                system.compressCornerStateTowards(corner_id, d2, 6)
        torch.cuda.synchronize(); tot += time.perf_counter() - t0
    ts.append(tot / 4)
print(os.environ.get("CARC_B200_LIB", "in-tree"), "compress s per iteration:", [round(t, 5) for t in ts])
