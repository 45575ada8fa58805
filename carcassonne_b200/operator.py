"""Device-resident center-site operator: the sum of stage-3 terms behind ``carc_operator_*``.

Host-side handle for what the reference builds in ``formExpectationStage3`` / ``formNormalizationStage3`` /
``formDenseStage3`` (reference tensors/_2d/sparse.py:100-161, tensors/_2d/dense.py:115-203): a list of
(stage2_0[tag_x], stage2_1[tag_y], O[tag_z]) triples applied to the center tensor and summed.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check
from .data import DeviceData, _empty, _ptr, _stream


def prejoin_halves(stage2_0, stage2_1):
    """reference dense.py:130-131 / 162-163: A = s2_0.join((0,1),4,5,2,3), B = s2_1.join((1,0),4,5,2,3)."""
    return stage2_0.join((0, 1), 4, 5, 2, 3), stage2_1.join((1, 0), 4, 5, 2, 3)


class Stage3Operator:
    """out[D0*,D1*,D2*,D3*,d] = sum_t B_t . (A_t . (O_t v)),  v of shape [D0,D1,D2,D3,d]."""

    def __init__(self, state_shape):
        d0, d1, d2, d3, d = (int(s) for s in state_shape)
        self.state_shape = (d0, d1, d2, d3, d)
        self.P = self.Q = d0 * d1
        self.R = self.S = d2 * d3
        self.d = d
        self._keep = []          # the term tensors must outlive the handle
        self._handle = C.c_void_p()
        check(lib.carc_operator_create(C.byref(self._handle), self.P, self.Q, self.R, self.S, d))
        self._finalized = False

    def add_term(self, A, B, site_operator=None):
        """A = [X, D0*, D1*, D0, D1], B = [X, D2*, D3*, D2, D3] (pre-joined), site_operator d x d or None."""
        d0, d1, d2, d3, d = self.state_shape
        if A.ndim != 5 or B.ndim != 5:
            raise ValueError("stage-3 halves must be pre-joined to rank 5")
        if A.shape[1:] != (d0, d1, d0, d1) or B.shape[1:] != (d2, d3, d2, d3) or A.shape[0] != B.shape[0]:
            raise ValueError("stage-3 halves {} / {} do not fit state shape {}".format(A.shape, B.shape,
                                                                                      self.state_shape))
        op = None
        if site_operator is not None:
            host = np.ascontiguousarray(
                site_operator.toArray() if hasattr(site_operator, "toArray") else site_operator, dtype=np.complex128)
            if host.shape != (d, d):
                raise ValueError("site operator must be {0} x {0}".format(d))
            op = host.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))
            self._keep.append(host)
        self._keep.extend([A, B])
        check(lib.carc_operator_add_term(self._handle, _ptr(A._t), _ptr(B._t), A.shape[0], op))
        return self

    def finalize(self):
        check(lib.carc_operator_finalize(self._handle))
        self._finalized = True
        return self

    def set_path(self, path):
        """0 auto, 1 fused kernel only, 2 unfused DMMA GEMM path, 3 fused kernel with the folded tiling only."""
        check(lib.carc_operator_set_path(self._handle, int(path)))
        return self

    @property
    def path(self):
        """The device path `__call__` runs: 1 fused kernel, 3 fused kernel with the folded tiling, 2 unfused GEMMs."""
        return lib.carc_operator_path(self._handle)

    @property
    def num_terms(self):
        return lib.carc_operator_num_terms(self._handle)

    @property
    def cost_of_multiply(self):
        return int(lib.carc_operator_cost_of_multiply(self._handle))

    @property
    def num_groups(self):
        return lib.carc_operator_num_groups(self._handle)

    @property
    def executed_flops(self):
        """FP64 flops the fused kernel issues per apply (<= 8 * cost_of_multiply thanks to shared-B grouping)."""
        return float(lib.carc_operator_executed_flops(self._handle))

    def apply_raw(self, v_t, out_t):
        """torch buffers in, torch buffer out; asynchronous on the current stream."""
        if not self._finalized:
            self.finalize()
        check(lib.carc_operator_apply(self._handle, _ptr(v_t), _ptr(out_t), _stream()))

    def __call__(self, v):
        if v.shape != self.state_shape:
            raise ValueError("state of shape {} does not fit operator for {}".format(v.shape, self.state_shape))
        out = _empty(self.state_shape)
        self.apply_raw(v._t, out)
        return DeviceData(out)

    def close(self):
        if self._handle:
            lib.carc_operator_destroy(self._handle)
            self._handle = C.c_void_p()
            self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
