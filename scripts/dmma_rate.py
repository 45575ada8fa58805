import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.cuda.init()
from carcassonne_b200 import _lib
tf = C.c_double()
for warps in (4, 8, 12, 16):
    _lib.check(_lib.lib.carc_dmma_rate(4000, warps, 100, C.byref(tf), None))
    print('warps/SM %d, 16 accumulators, 4x4 distinct operands: %.1f TFLOP/s' % (warps, tf.value))
for warps in (8,):
    row = []
    for chains in (2, 4, 8, 16):
        _lib.check(_lib.lib.carc_dmma_rate(4000, warps, chains, C.byref(tf), None))
        row.append("%5.1f" % tf.value)
    print("warps/SM %2d: chains 2/4/8/16 -> %s TFLOP/s" % (warps, " ".join(row)))
# DMMA next to DFMA: do the two FP64 instruction streams overlap?
tm, tf2 = C.c_double(), C.c_double()
for warps in (8, 16):
    for nd, nf in ((16, 0), (0, 64), (16, 16), (16, 32), (16, 64), (16, 128)):
        _lib.check(_lib.lib.carc_fp64_mix_rate(4000, warps, nd, nf, C.byref(tm), C.byref(tf2), None))
        print("warps/SM %2d: %2d DMMA + %3d DFMA per trip -> DMMA %5.1f + DFMA %5.1f = %5.1f TFLOP/s"
              % (warps, nd, nf, tm.value, tf2.value, tm.value + tf2.value))
