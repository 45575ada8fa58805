"""Where do the sporadic 0.2 - 1.5 s stalls of a host-driven sweep come from?  Runs the sweep_bench iteration loop several
times at one size, times every phase call, and prints the calls that took more than 4x their phase's median together with
what changed in the allocators during the call."""
import sys, os, time, json, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from carcassonne_b200 import synthetic

D, chi, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 6
records = []

# every library call timed on the host: calls slower than 50 ms are reported by name (which C entry does a stall sit in?)
from carcassonne_b200 import _lib
slow_calls = []

def _timed(name, fn):
    def wrapper(*args):
        t0 = time.perf_counter()
        rc = fn(*args)
        dt = time.perf_counter() - t0
        if dt > 0.05:
            slow_calls.append((name, round(dt, 3)))
        return rc
    return wrapper

for _name in _lib.SIGNATURES:
    setattr(_lib.lib, _name, _timed(_name, getattr(_lib.lib, _name)))
_sync = torch.cuda.synchronize

def snap():
    st = torch.cuda.memory_stats()
    free, total = torch.cuda.mem_get_info()
    return {"segments": st.get("num_device_alloc", 0), "frees": st.get("num_device_free", 0), "retries": st.get("num_alloc_retries", 0),
            "reserved_mb": st.get("reserved_bytes.all.current", 0) >> 20, "free_mb": free >> 20, "gc": gc.get_count()}

def timed(key, rep, it, fn):
    torch.cuda.synchronize()
    a = snap(); t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    records.append((key, rep, it, dt, a, snap()))

for rep in range(reps):
    system = synthetic.device_system(chi, D, seed=0)
    np.random.seed(0)
    for it, direction in enumerate((0, 1, 2, 3)):
        timed("minimize", rep, it, lambda: system.minimizeExpectation())
        timed("contract", rep, it, lambda: system.contractTowards(direction))
        def compress():
            for corner_id in range(4):
                for d2 in range(2):
                    system.compressCornerStateTowards(corner_id, d2, chi)
        timed("compress", rep, it, compress)
by = {}
for key, rep, it, dt, a, b in records:
    if rep > 0: by.setdefault((key, it), []).append(dt)
med = {k: float(np.median(v)) for k, v in by.items()}
print("medians (s):", {"%s[%d]" % k: round(v, 4) for k, v in med.items()})
for key, rep, it, dt, a, b in records:
    m = med.get((key, it), dt)
    if rep > 0 and dt > 4 * m:
        print("STALL rep %d %s[%d]: %.3f s (median %.4f)  before %s  after %s" % (rep, key, it, dt, m, a, b))
print("library calls slower than 50 ms:", slow_calls)
print("total per rep:", [round(sum(r[3] for r in records if r[1] == rep), 3) for rep in range(reps)])
