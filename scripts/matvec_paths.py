"""Center-site expectation matvec vs bond dimension, one row per device path (1 = fused kernel, 3 = fused kernel with the
folded tiling, 0 = the library's automatic choice), TFIM term structure (T = 9), tensors larger than L2.

    python scripts/matvec_paths.py [--sizes 3:9,4:16,5:16,6:16,7:16,8:16] [--steps 5] [--out file.md]

Every path is checked against the unfused DMMA-GEMM path on an X slab before it is timed."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from carcassonne_b200.data import DeviceData  # noqa: E402
from carcassonne_b200.operator import Stage3Operator  # noqa: E402


def build(D, chi, terms_table, lo=None, hi=None, tensors=None):
    na, nb, terms = terms_table
    X = chi ** 4
    if tensors is None:
        gen = torch.Generator(device="cuda")
        gen.manual_seed(7)
        scale = 1.0 / (D * D * np.sqrt(X))

        def rnd(*shape):
            t = torch.empty(shape, dtype=torch.complex128, device="cuda")
            torch.view_as_real(t).normal_(generator=gen)
            return t.mul_(scale)

        tensors = ([rnd(X, D, D, D, D) for _ in range(na)], [rnd(X, D, D, D, D) for _ in range(nb)])
    A, B = tensors
    op = Stage3Operator((D, D, D, D, 2))
    for a, b, o in terms:
        op.add_term(DeviceData(A[a] if lo is None else A[a][lo:hi]), DeviceData(B[b] if lo is None else B[b][lo:hi]), o)
    return op.finalize(), tensors


def time_op(op, v, out, steps):
    for _ in range(3):
        op.apply_raw(v, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        op.apply_raw(v, out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="3:9,4:16,5:16,6:16,7:16,8:16")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--paths", default="1,3,2,0", help="device paths to time (2 = unfused DMMA GEMMs)")
    args = ap.parse_args()
    args.paths = [int(x) for x in args.paths.split(",")]
    table = bench.term_table_device("tfim")
    rows = ["| D | chi | path | kernel | ms / matvec | reference-equivalent GFLOP/s | executed TFLOP/s | vs unfused slab |",
            "|---|---|---|---|---|---|---|---|"]
    for size in args.sizes.split(","):
        D, chi = (int(x) for x in size.split(":"))
        X = chi ** 4
        op, tensors = build(D, chi, table)
        v = torch.empty((D, D, D, D, 2), dtype=torch.complex128, device="cuda")
        torch.view_as_real(v).normal_()
        out = torch.empty_like(v)
        lo, hi = X // 2, min(X, X // 2 + 300)
        ref_op, _ = build(D, chi, table, lo, hi, tensors)
        ref_op.set_path(2)
        ref = torch.empty_like(v)
        ref_op.apply_raw(v, ref)
        flops = 8.0 * bench.cost_of_multiply(table[2], X, D, 2)
        for path in args.paths:
            op.set_path(path)
            slab_op, _ = build(D, chi, table, lo, hi, tensors)
            slab_op.set_path(path)
            got = torch.empty_like(v)
            try:
                slab_op.apply_raw(v, got)
            except Exception:              # shape outside this kernel's envelope
                slab_op.close()
                continue
            err = float((got - ref).norm() / ref.norm())
            ms = time_op(op, v, out, args.steps)
            rows.append("| %d | %d | %d | %d | %.3f | %.0f | %.2f | %.1e |" % (
                D, chi, path, op.path, ms, flops / ms / 1e6, op.executed_flops / ms / 1e9, err))
            print(rows[-1], flush=True)
            slab_op.close()
        ref_op.close()
        op.close()
        del tensors, op, ref_op
        torch.cuda.empty_cache()
    text = "\n".join(rows) + "\n"
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)


if __name__ == "__main__":
    main()
