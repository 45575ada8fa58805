"""Per-warp mbarrier wait shares of the folded stage-3 kernel (library built with `make EXTRA=-DS3F_PROFILE`)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench  # noqa: E402
import matvec_paths  # noqa: E402
from carcassonne_b200 import _lib  # noqa: E402

for size in (sys.argv[1] if len(sys.argv) > 1 else "6:12,5:12,7:12").split(","):
    D, chi = (int(x) for x in size.split(":"))
    table = bench.term_table_device("tfim")
    op, tensors = matvec_paths.build(D, chi, table)
    op.set_path(3)
    v = torch.empty((D, D, D, D, 2), dtype=torch.complex128, device="cuda")
    torch.view_as_real(v).normal_()
    out = torch.empty_like(v)
    ms = matvec_paths.time_op(op, v, out, 3)
    buf = np.zeros((148, 12, 4), dtype=np.uint64)
    _lib.check(_lib.lib.carc_stage3f_profile_read(buf.ctypes.data))
    b = buf.astype(np.float64)
    print("D=%d chi=%d: %.3f ms, %.2f TFLOP/s executed" % (D, chi, ms, op.executed_flops / ms / 1e9))
    tot = b[:, :, 3]
    act = tot > 0
    print("  warp: share of its time in [first-term A wait, later-term A wait, B wait], mean over CTAs; total Mcycles")
    for w in range(12):
        if act[:, w].any():
            sh = [(b[:, w, k][act[:, w]] / tot[:, w][act[:, w]]).mean() for k in range(3)]
            print("  warp %2d (sub-partition %d): %5.1f%% %5.1f%% %5.1f%%   total %.2f" % (
                w, w % 4, 100 * sh[0], 100 * sh[1], 100 * sh[2], tot[:, w][act[:, w]].mean() / 1e6))
    cta_tot = tot.max(axis=1)
    print("  CTA duration Mcycles: min %.2f median %.2f max %.2f" % (cta_tot[cta_tot > 0].min() / 1e6,
                                                                  np.median(cta_tot[cta_tot > 0]) / 1e6, cta_tot.max() / 1e6))
    op.close()
    del op, tensors
    torch.cuda.empty_cache()
