#!/usr/bin/env python
"""Seconds per sweep iteration vs bond dimension: one iteration = minimizeExpectation + contractTowards(direction)
+ ConstantStateCompressionPolicy(chi) (8 corner compressions), on a synthetic double-layer TFIM environment; four
iterations (one per direction) are timed per (D, chi).  The CPU column runs the oracle's restatement of the same
calls on the host cores for the sizes where it finishes in seconds.

    python scripts/sweep_bench.py [--sizes 2x4,3x6,4x8] [--cpu-max-D 4] [--phases]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def device_iterations(chi, D, phases, seed=0, directions=(0, 1, 2, 3)):
    import torch
    from carcassonne_b200 import synthetic
    system = synthetic.device_system(chi, D, seed=seed)
    np.random.seed(seed)
    out = {"minimize": 0.0, "contract": 0.0, "compress": 0.0}
    stats_all = []

    def timed(key, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        out[key] += time.perf_counter() - t0

    t_start = time.perf_counter()
    for direction in directions:
        st = {}
        timed("minimize", lambda: system.minimizeExpectation(statistics=st))
        stats_all.append(st)
        timed("contract", lambda: system.contractTowards(direction))

        def compress():
            for corner_id in range(4):
                for d2 in range(2):
                    system.compressCornerStateTowards(corner_id, d2, chi)
        timed("compress", compress)
    torch.cuda.synchronize()
    total = time.perf_counter() - t_start
    energy = system.computeExpectation()
    out["total"] = total
    out["per_iteration"] = total / len(directions)
    out["mults"] = [s.get("multiplications") for s in stats_all]
    out["normalization"] = [s.get("normalization") for s in stats_all]
    out["expectation"] = [float(np.real(energy)), float(np.imag(energy))]
    out["tags_per_side"] = [len(s) for s in system.sides]
    return out


def cpu_iterations(chi, D, seed=0, directions=(0, 1, 2, 3)):
    from carcassonne_b200 import synthetic
    from oracle import tags
    from oracle.system import System
    corners, sides, center = synthetic.double_layer_environment(chi, D, 2, 2, seed)
    Os, UDs, LRs = synthetic.tfim_operator_arrays(1.0)
    op = tags.make_sparse_operator(Os, UDs, LRs)
    system = System([{tags.I: c} for c in corners], [{tags.I: s} for s in sides], center, op)
    np.random.seed(seed)
    out = {"minimize": 0.0, "contract": 0.0, "compress": 0.0}
    t_start = time.perf_counter()
    for direction in directions:
        t0 = time.perf_counter()
        system.minimize_expectation()
        t1 = time.perf_counter()
        system.contract_towards(direction)
        t2 = time.perf_counter()
        for corner_id in range(4):
            for d2 in range(2):
                system.compress_corner_state_towards(corner_id, d2, chi)
        t3 = time.perf_counter()
        out["minimize"] += t1 - t0
        out["contract"] += t2 - t1
        out["compress"] += t3 - t2
    total = time.perf_counter() - t_start
    e = system.expectation()
    out["total"] = total
    out["per_iteration"] = total / len(directions)
    out["expectation"] = [float(np.real(e)), float(np.imag(e))]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2x4,3x6,4x8")
    ap.add_argument("--cpu-max-D", type=int, default=4)
    ap.add_argument("--no-gpu", action="store_true")
    args = ap.parse_args()
    rows = []
    for item in args.sizes.split(","):
        D, chi = (int(x) for x in item.split("x"))
        row = {"D": D, "chi": chi}
        if not args.no_gpu:
            device_iterations(min(chi, 2), min(D, 2), False)      # warm-up (kernels, tables, allocator)
            row["gpu"] = device_iterations(chi, D, False)
        if D <= args.cpu_max_D:
            row["cpu"] = cpu_iterations(chi, D)
            if "gpu" in row:
                row["speedup"] = row["cpu"]["per_iteration"] / row["gpu"]["per_iteration"]
        print(json.dumps(row), flush=True)
        rows.append(row)


if __name__ == "__main__":
    main()
