"""One process per GPU on one node: X-slab sharding of the center-site operator and the NVLink peer-memory
communicator.  ``torch.distributed`` is used for rendezvous and for exchanging the CUDA IPC
handles only; the data path is ``carc_comm_*`` (csrc/comm.cu) -- or NCCL when ``reduce="nccl"`` is asked for, which
is kept as the comparison baseline.

The planning helpers (``slab_bounds``, ``shard_terms``) are pure host logic and run on CPU (gloo tests).
"""
import ctypes as C

import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def slab_bounds(X, rank, world):
    """Contiguous slab [lo, hi) of the joined environment bond owned by ``rank``; slabs tile [0, X) exactly and
    differ in length by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank {} outside world of {}".format(rank, world))
    return X * rank // world, X * (rank + 1) // world


def shard_terms(terms, rank, world):
    """[(A, B, O)] with A = [X, ...], B = [X, ...] -> the same list restricted to this rank's X slab.  Works on
    anything that slices along axis 0 (torch tensors, ndarrays); empty slabs are dropped."""
    out = []
    for A, B, O in terms:
        lo, hi = slab_bounds(A.shape[0], rank, world)
        if hi > lo:
            out.append((A[lo:hi], B[lo:hi], O))
    return out


def allreduce_sum_(tensor):
    """In-place sum over ranks through torch.distributed (NCCL on GPU tensors, gloo on CPU tensors)."""
    if world() > 1:
        dist.all_reduce(tensor)
    return tensor


class PeerComm:
    """carc_comm handle: exchange buffers and flags mapped into every peer through CUDA IPC."""

    def __init__(self, max_elems):
        from ._lib import lib, check
        self.rank, self.world = rank(), world()
        self._handle = C.c_void_p()
        check(lib.carc_comm_create(C.byref(self._handle), self.rank, self.world, int(max_elems)))
        if self.world > 1:
            mine = (C.c_ubyte * 128)()
            check(lib.carc_comm_local_handles(self._handle, mine))
            local = torch.tensor(list(mine), dtype=torch.uint8, device="cuda")
            gathered = [torch.empty_like(local) for _ in range(self.world)]
            dist.all_gather(gathered, local)
            blob = bytes(torch.cat(gathered).cpu().tolist())
            check(lib.carc_comm_connect(self._handle, blob))
            dist.barrier()

    def allreduce_(self, t):
        from ._lib import lib, check
        if t.dtype != torch.complex128 or not t.is_cuda or not t.is_contiguous():
            raise TypeError("PeerComm.allreduce_ needs a contiguous complex128 CUDA tensor")
        check(lib.carc_comm_allreduce(self._handle, C.c_void_p(t.data_ptr()), t.numel(),
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return t

    def timed_out(self):
        from ._lib import lib, check
        flag = C.c_int(0)
        check(lib.carc_comm_status(self._handle, C.byref(flag)))
        return bool(flag.value)

    def close(self):
        from ._lib import lib
        if self._handle:
            if self.world > 1 and dist.is_initialized():
                torch.cuda.synchronize()
                dist.barrier()           # nobody unmaps while a peer may still be reading
            lib.carc_comm_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            if self._handle and self.world == 1:
                self.close()
        except Exception:
            pass


def shard_operator(operator, comm):
    """Attach the communicator to a ``Stage3Operator`` whose terms hold this rank's X slabs: every apply then
    returns the full vector on every rank."""
    from ._lib import lib, check
    check(lib.carc_operator_set_comm(operator._handle, comm._handle if comm is not None else None))
    operator._comm = comm
    return operator
