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


class ExchangeTimeout(RuntimeError):
    """A bounded device-side wait of the peer all-reduce gave up: a rank fell out of step (CARC_ERR_EXCHANGE)."""


class EnvironmentSharding:
    """System-level multi-GPU mode: while active, every ``System`` of this process builds only
    its rank's X slab of the stage-2 halves (``formExpectationStage2``), its expectation / normalization operators
    finish each matvec with the all-reduce of the output vector, and the dense matrices of the
    ``isCheaperToFormMatrix`` branches are summed over ranks -- so ``System.minimizeExpectation`` and everything built
    on the multipliers run on all GPUs, while absorption / compression (small next to the environment) are
    replicated.  Every rank must execute the same sequence of calls on the same (seeded) system."""

    def __init__(self):
        self.rank, self.world = rank(), world()
        self._comm = None
        self._comm_elems = 0
        self._retired = []      # outgrown communicators stay mapped while operators built on them may still exist

    @property
    def slab(self):
        return (self.rank, self.world)

    def comm_for(self, n):
        """The peer communicator, grown (collectively: every rank asks for the same sizes in the same order) to hold
        vectors of ``n`` elements."""
        if self.world == 1:
            return None
        if self._comm is None or n > self._comm_elems:
            if self._comm is not None:
                self._retired.append(self._comm)
            self._comm_elems = max(int(n), 1 << 14)
            self._comm = PeerComm(self._comm_elems)
        return self._comm

    def attach(self, operator):
        """Make ``operator`` (a finalized ``Stage3Operator`` over this rank's slabs) return the sum over ranks."""
        if self.world == 1:
            return operator
        # the exchange is part of carc_operator_apply itself (the device solver calls it without coming back to Python),
        # so it is always the NVLink peer-memory all-reduce fused into the partial-sum pass; NCCL is only used for the
        # one-off sums of dense matrices
        shard_operator(operator, self.comm_for(operator.P * operator.R * operator.d))
        return operator

    def sum_matrix_(self, matrix):
        """In-place sum over ranks of a dense matrix formed from this rank's slabs (``Multiplier.formMatrix``)."""
        if self.world > 1:
            dist.all_reduce(matrix._t)
            matrix._touch()
        return matrix

    def raise_if_timed_out(self):
        if self._comm is not None and self._comm.timed_out():
            raise ExchangeTimeout("peer all-reduce timed out on rank {}".format(self.rank))

    def close(self):
        for comm in self._retired + ([self._comm] if self._comm is not None else []):
            comm.close()
        self._retired, self._comm = [], None


_environment_sharding = None


def shard_environment():
    """Turn the system-level multi-GPU mode on (collective; needs an initialised process group)."""
    global _environment_sharding
    unshard_environment()
    from .tensors._2d.sparse import environment_cache
    environment_cache.clear()
    _environment_sharding = EnvironmentSharding() if world() > 1 else None
    return _environment_sharding


def unshard_environment():
    global _environment_sharding
    if _environment_sharding is not None:
        _environment_sharding.close()
        _environment_sharding = None
        from .tensors._2d.sparse import environment_cache
        environment_cache.clear()


def environment_sharding():
    return _environment_sharding


def balance_terms(costs, world):
    """Term sharding (BASELINE config 4): whole (A_t, B_t) pairs per rank.  Longest-processing-time greedy: terms in
    decreasing cost go to the least loaded rank -> list of term indices per rank.  Pure host logic."""
    order = sorted(range(len(costs)), key=lambda t: (-costs[t], t))
    load, owner = [0] * world, [[] for _ in range(world)]
    for t in order:
        r = min(range(world), key=lambda q: (load[q], q))
        owner[r].append(t)
        load[r] += costs[t]
    return [sorted(o) for o in owner]


def shard_operator(operator, comm):
    """Attach the communicator to a ``Stage3Operator`` whose terms hold this rank's X slabs: every apply then
    returns the full vector on every rank."""
    from ._lib import lib, check
    check(lib.carc_operator_set_comm(operator._handle, comm._handle if comm is not None else None))
    operator._comm = comm
    return operator
