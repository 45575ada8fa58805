"""Achieved HBM bandwidth of the memory-bound kernels on the path (permute = `join`/`transpose`/`conj`, axpby, reductions)
against the measured copy bandwidth in MEASURED_PEAKS.json.  Shapes are the joins of the reference's recipes at
benchmark scale (tensors/_2d/dense.py: the 8-axis corner/side joins, the stage-3 pre-joins, a conj, a plain transpose).

    python scripts/hbm_kernels.py [--out profiles/file.md]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from carcassonne_b200.data import DeviceData  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    peak = 6453.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def rnd(*shape):
        t = torch.empty(shape, dtype=torch.complex128, device="cuda")
        torch.view_as_real(t).normal_()
        return DeviceData(t)

    chi, D = 16, 8
    cases = [
        ("stage-3 pre-join half 0: [x,y,D,D,D,D].join((0,1),4,5,2,3)   (dense.py:130)", (chi * chi // 4, chi * chi, D, D, D, D),
         lambda t: t.join((0, 1), 4, 5, 2, 3), 1),
        ("stage-3 pre-join half 1: join((1,0),4,5,2,3)                  (dense.py:131)", (chi * chi // 4, chi * chi, D, D, D, D),
         lambda t: t.join((1, 0), 4, 5, 2, 3), 1),
        ("8-axis join of an absorbed corner, both bond pairs swapped       (dense.py:11-21)",
         (chi * D, chi * D, 1, chi, D, chi, D, 1), lambda t: t.join(1, 0, 2, (5, 6), (3, 4), 7), 1),
        ("Hermitian symmetrisation: join(1,0,2,4,3,5,7,6).conj()        (system/_2d.py:42-47)",
         (chi * 2, chi * 2, 1, chi * 2, chi * 2, 1, D, D), lambda t: t.join(1, 0, 2, 4, 3, 5, 7, 6).conj(), 2),
        ("plain 2-D transpose 16384 x 8192", (16384, 8192), lambda t: t.transpose(1, 0), 1),
        ("conj (elementwise)", (chi ** 4 // 2, D * D, D * D), lambda t: t.conj(), 1),
    ]
    rows = ["| kernel / case | elements | ms | GB/s (read + write) | of %.0f GB/s measured copy |" % peak, "|---|---|---|---|---|"]
    for name, shape, fn, passes in cases:
        t = rnd(*shape)
        n = t.size()
        ms = timed(lambda: fn(t))
        gbs = passes * 32.0 * n / ms / 1e6
        rows.append("| permute: %s | %d | %.3f | %.0f | %.2f |" % (name, n, ms, gbs, gbs / peak))
        print(rows[-1], flush=True)
        del t
        torch.cuda.empty_cache()
    n = 1 << 28
    x, y = rnd(n), rnd(n)
    ms = timed(lambda: y.__iadd__(x))
    gbs = 48.0 * n / ms / 1e6
    rows.append("| axpby (`+=`, reads x and y, writes y) | %d | %.3f | %.0f | %.2f |" % (n, ms, gbs, gbs / peak))
    print(rows[-1], flush=True)
    ms = timed(lambda: x.norm())
    gbs = 16.0 * n / ms / 1e6
    rows.append("| sumsq (`norm`, incl. the scalar read-back) | %d | %.3f | %.0f | %.2f |" % (n, ms, gbs, gbs / peak))
    print(rows[-1], flush=True)
    ms = timed(lambda: x.contractWithAlongAll(y))
    gbs = 32.0 * n / ms / 1e6
    rows.append("| dot (`contractWithAlongAll`, incl. the read-back) | %d | %.3f | %.0f | %.2f |" % (n, ms, gbs, gbs / peak))
    print(rows[-1], flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
