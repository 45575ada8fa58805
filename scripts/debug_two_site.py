import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import test_gpu_two_site as t
from carcassonne_b200.data import DeviceData as dd, _init_constants
_init_constants()
for fn, args in [(t.test_two_steps_chain, (dd, 2, 1)), (t.test_many_steps_uniform, (dd, 1, "LR")), (t.test_one_step_product_expectation, (dd, 4, 0))]:
    try:
        fn(*args); print(fn.__name__, args[1:], "ok")
    except Exception as e:
        print(fn.__name__, args[1:], "FAILED:", repr(e)[:300])
        traceback.print_exc(limit=6)
